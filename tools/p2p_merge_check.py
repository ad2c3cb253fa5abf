"""Multi-GPU check of the fused peer-memory merge (run under torchrun, one rank per GPU):
each rank renders its sample range into a library-owned film, films are exchanged as CUDA IPC handles, every rank runs
ONE film_gather_merge kernel over its pixel slice (reading all peers over NVLink, writing into rank 0's merged film and
images), and rank 0 compares with the NCCL merge and prints timings of both."""
import importlib, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np
import torch, torch.distributed as dist
import common
cuda = importlib.import_module("daily-ray-trace_b200.cuda")
film_mod = importlib.import_module("daily-ray-trace_b200.film")

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = h = int(sys.argv[1]) if len(sys.argv) > 1 else 512
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, spp, 4)
n = scene.num_wavelengths
ctx = cuda.Context(local)
ctx.upload_scene(scene, camera, tables)
prm = common.structs.RenderParams(w, h, rank * spp, (rank + 1) * spp, 4, cfg.pixel_scheme, 5)
dev = torch.device("cuda", local)

# NCCL path (torch-allocated planes)
nccl_film = film_mod.FilmPlanes(w, h, n, dev)
ctx.render_device(prm, nccl_film.as_drt_film())
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
film_mod.merge_distributed_(nccl_film)
torch.cuda.synchronize(); dist.barrier()
t_nccl = time.perf_counter() - t0

# fused peer-memory path (library-allocated planes, CUDA IPC)
mine = ctx.film_alloc(w, h)
merged = ctx.film_alloc(w, h) if rank == 0 else None
imgs = [torch.zeros(w * h, dtype=torch.int32, device=dev) for _ in range(3)] if rank == 0 else None
ctx.render_device(prm, mine)
torch.cuda.synchronize()
handles = [None] * world
dist.all_gather_object(handles, ctx.film_ipc_export(mine))
root = [ctx.film_ipc_export(merged) if rank == 0 else None]
dist.broadcast_object_list(root, src=0)
# IPC handles of torch tensors cannot be exported portably: the images live in rank 0's film-sized scratch instead
films = [mine if r == rank else ctx.film_ipc_open(handles[r]) for r in range(world)]
dst = merged if rank == 0 else ctx.film_ipc_open(root[0])
npix = w * h
p0, p1 = rank * npix // world, (rank + 1) * npix // world
dist.barrier()
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    ctx.film_merge_many(dst, films, w, h, p0, p1, bgra=[t.data_ptr() for t in imgs] if rank == 0 else None)
    torch.cuda.synchronize(); dist.barrier()
    t_p2p = time.perf_counter() - t0
if rank == 0:
    def plane(ptr, count):
        out = torch.empty(count, dtype=torch.float32, device=dev)
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
        cudart.cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(count * 4), 3)
        return out.cpu().numpy()
    ok = True
    for name, count in (("sum", npix * n), ("mean", npix * n), ("m2", npix * n), ("filter", npix)):
        got = plane(getattr(merged, name), count)
        ref = getattr(nccl_film, name).cpu().numpy().reshape(-1)
        floor = 1e-5 * np.abs(ref).max()
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), floor)
        print(f"{name}: max rel diff fused-vs-NCCL {rel.max():.2e}")
        ok &= rel.max() < 3e-4
    print(f"world {world}, {w}x{h}: NCCL merge {1e3 * t_nccl:.2f} ms, fused peer-memory merge+3 images {1e3 * t_p2p:.2f} ms, match={ok}")
dist.barrier()
dist.destroy_process_group()

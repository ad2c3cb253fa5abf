"""One render launch of a named workload, for ncu (python tools/profile_render.py scene W H spp [f64])."""
import importlib, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import common, oracledriver
import torch
cuda = importlib.import_module("daily-ray-trace_b200.cuda")
scene, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctx = cuda.Context(0)
cfg, tables, sc, cam = common.load(scene, w, h, spp, 4)
ctx.upload_scene(sc, cam, tables)
ctx.set_geometry_precision(1 if len(sys.argv) > 5 and sys.argv[5] == "f64" else 0)
n = sc.num_wavelengths
planes = [torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h, device="cuda"), torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h*n, device="cuda")]
film = cuda.film_from_tensors(*planes)
prm = oracledriver.params(w, h, 0, spp, 4, 2, 1)
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ctx.render_device(prm, film); b.record(); torch.cuda.synchronize()
    print(f"{scene} {w}x{h}x{spp}: {a.elapsed_time(b):.3f} ms, {w*h*spp/a.elapsed_time(b)/1e3:.1f} Mpaths/s")

#!/usr/bin/env python
"""bench.py -- camera paths/s of the render hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            one rank per GPU (torchrun for N > 1)
    python bench.py --impl reference ...                           the reference's own CPU code on the host cores

One "step" is one complete render of the workload: every pixel, `spp` samples per pixel per GPU, film accumulated, and
(N > 1) the per-GPU films combined over NCCL.  Default workload = BASELINE.json configs[1]: init_cornell.scn,
1024x1024, 1024 spp, max_cast_depth 4, pixel_random.  Scaling is WEAK: every GPU renders `spp` samples of every pixel
(global sample indices [rank*spp, (rank+1)*spp)), so N GPUs deliver an N*spp image.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

PKG = "daily-ray-trace_b200"
import atexit
import shutil
import tempfile
SCRATCH = os.environ.get("DRT_BENCH_SCRATCH") or tempfile.mkdtemp(prefix="drt_bench_")
os.environ["DRT_BENCH_SCRATCH"] = SCRATCH          # pool workers (spawned) share the parent's directory
if os.environ.get("DRT_BENCH_SCRATCH_OWNER") is None:
    os.environ["DRT_BENCH_SCRATCH_OWNER"] = str(os.getpid())
    atexit.register(lambda: shutil.rmtree(SCRATCH, ignore_errors=True))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="init_cornell")
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--geometry", default="f32", choices=["f32", "f64"])
    ap.add_argument("--merge", default="scatter", choices=["scatter", "p2p", "nccl"],
                    help="multi-GPU film combination: render kernel scatters finished pixels to their owner over NVLink + local merge "
                         "(default), gather-merge kernel over peer memory, or NCCL collectives")
    return ap.parse_args()


def workload_name(a):
    return f"{a.scene}.scn {a.width}x{a.height} @{a.spp}spp/GPU depth{a.depth} pixel_random, N=69 wavelengths"


# ----------------------------------------------------------------------------------------------- CPU arms

def _cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_W = {}


def _cpu_worker_init(kind, scene, width, height, depth, seed):
    """Runs in every pool process: load the scene once into the reference library (or the oracle port)."""
    import common
    host = importlib.import_module(PKG + ".host")
    assets = os.path.join(REPO, "assets")
    if kind == "reference":
        import refdriver
        parsed = host.parse_scene_text(open(common.scene_path(scene)).read())
        root = os.path.join(SCRATCH, f"ref_root_{os.getpid()}")
        refdriver.make_root(root, assets, host.scene_to_text(parsed), "bench_scene.scn")
        cfg = host.make_config_text(scene="scenes\\bench_scene.scn", width=width, height=height, spp=1, depth=depth)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)          # the reference's CSV loader and scene dump are chatty (read_scene.c:803-844)
        try:
            _W["ref"] = refdriver.Ref(root, cfg, seed=seed)
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)
    else:
        import oracledriver
        cfg, tables, sc, cam = common.load(scene, width, height, 1, depth)
        _W["oracle"] = (oracledriver, sc, cam, cfg)
    _W["kind"], _W["dims"], _W["depth"], _W["seed"] = kind, (width, height), depth, seed


def _cpu_worker_render(job):
    """Renders rows [y0, y1) of sample index s through the reference's sample_scene; returns the number of paths."""
    y0, y1, s = job
    w, h = _W["dims"]
    if _W["kind"] == "reference":
        _W["ref"].render_tile(0, y0, w, y1, s, s + 1)
    else:
        od, sc, cam, cfg = _W["oracle"]
        prm = od.params(w, h, s, s + 1, _W["depth"], cfg.pixel_scheme, _W["seed"])
        od.render_tile(sc, cam, prm, 0, y0, w, y1)
    return (y1 - y0) * w


def _cpu_worker_tile(job):
    """Per-path spectra [(pixels), samples, N] of a pixel rectangle through the reference's sample_scene (same per-path streams)."""
    x0, y0, x1, y1, s0, s1 = job
    w, h = _W["dims"]
    if _W["kind"] == "reference":
        return _W["ref"].render_tile(x0, y0, x1, y1, s0, s1, want_paths=True)[3]
    od, sc, cam, cfg = _W["oracle"]
    prm = od.params(w, h, s0, s1, _W["depth"], cfg.pixel_scheme, _W["seed"])
    return od.render_tile(sc, cam, prm, x0, y0, x1, y1, want_paths=True)[3]


class CpuArm:
    """The reference's CPU implementation of the path on all host cores: one single-threaded process per core (the
    reference has global RNG/scratch state), rows of the image split between them."""

    def __init__(self, a):
        import multiprocessing as mp
        import refdriver
        self.kind = "reference" if refdriver.available() else "port"
        self.cores = _cpu_cores()
        self.a = a
        os.makedirs(SCRATCH, exist_ok=True)
        self.pool = mp.get_context("spawn").Pool(self.cores, _cpu_worker_init,
                                                (self.kind, a.scene, a.width, a.height, a.depth, a.seed))

    def step(self, sample_index, rows=None):
        """One bounded sample of the workload: `rows` image rows (default all) at ONE sample per pixel."""
        h = self.a.height if rows is None else min(rows, self.a.height)
        per = max(1, h // (self.cores * 4))
        jobs = [(y, min(y + per, h), sample_index) for y in range(0, h, per)]
        t0 = time.perf_counter()
        paths = sum(self.pool.map(_cpu_worker_render, jobs, chunksize=1))
        return paths, time.perf_counter() - t0

    def tile_paths(self, x0, y0, x1, y1, s0, s1):
        import numpy as np
        rows = self.pool.map(_cpu_worker_tile, [(x0, y, x1, y + 1, s0, s1) for y in range(y0, y1)], chunksize=1)
        return np.concatenate(rows, axis=0)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(a)
    for i in range(a.warmup):
        arm.step(i)
    paths, secs = 0, 0.0
    for i in range(a.steps):
        p, t = arm.step(a.warmup + i)
        paths += p
        secs += t
    arm.close()
    value = paths / secs
    sample = f"{a.width}x{a.height} at 1 sample per pixel per step ({paths // a.steps} paths/step), rows split over {arm.cores} processes"
    line = {
        "impl": "reference", "metric": "camera_paths_per_sec", "value": value, "unit": "paths/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * secs / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": f"shipped scene assets/scenes/{a.scene}.scn",
        "config": {"workload": workload_name(a), "note": "CPU arm renders a bounded sample of the same workload: cost per sample pass is independent of spp"},
        "cpu_baseline": {"value": value, "unit": "paths/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz, mx, reasons = [], None, set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                mhz.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mhz.sort()
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(mhz)}


# ----------------------------------------------------------------------------------------------- roofline arithmetic

def algorithmic_flops_per_path(scene, st, n):
    """flops/path = R_c*F_c + R_s*F_s + B*19N + 8N + 40, with per-surface costs 33 (plane) / 24 (sphere)
    and the ray, shadow-ray and shaded-bounce counts measured by the kernel itself."""
    per_ray = 0
    for i in range(scene.num_surfaces):
        t = scene.surfaces[i].type
        per_ray += 33 if t == 3 else 24 if t == 2 else 0
    rc, rs, b = st.closest_rays / st.paths, st.shadow_rays / st.paths, st.shaded_bounces / st.paths
    return rc * (per_ray + 32) + rs * (per_ray + 33) + b * 19 * n + 8 * n + 40, rc, rs, b


# ----------------------------------------------------------------------------------------------- GPU arm

def run_b200_arm(a):
    import torch
    import torch.distributed as dist
    import common
    cuda = importlib.import_module(PKG + ".cuda")
    film_mod = importlib.import_module(PKG + ".film")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 and world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks (WORLD_SIZE={world})")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cfg, tables, scene, camera = common.load(a.scene, a.width, a.height, a.spp, a.depth)
    n = scene.num_wavelengths
    ctx = cuda.Context(local)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F64 if a.geometry == "f64" else cuda.GEOMETRY_F32)
    dev = torch.device("cuda", local)
    peers, merge_kind = None, "none"
    if world > 1 and a.merge in ("scatter", "p2p"):
        try:
            peers = film_mod.PeerFilmGroup(ctx, a.width, a.height, scatter=a.merge == "scatter")
            merge_kind = ("scatter: the render kernel stores each finished pixel into its owner's staging film over NVLink (CUDA IPC), "
                          "then one local merge+images kernel per rank") if a.merge == "scatter" else \
                         "p2p: one fused gather-merge+images kernel per rank over CUDA-IPC peer memory"
        except Exception as exc:
            print(f"[rank {rank}] peer-memory merge unavailable ({exc}); using NCCL", file=sys.stderr)
            peers = None
    if world > 1 and peers is None:
        merge_kind = "nccl: 2 all_reduce + 1 reduce"

    class _DevArray:   # view of library-owned device memory for torch copies (plumbing only)
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}

    npix_all = a.width * a.height
    if peers is not None:
        drt_film = peers.mine
        film = None
        root_planes = None
        if rank == 0:
            m = peers.merged
            root_planes = {"sum": torch.as_tensor(_DevArray(m.sum, (npix_all, n)), device=dev),
                           "filter": torch.as_tensor(_DevArray(m.filter, (npix_all,)), device=dev),
                           "mean": torch.as_tensor(_DevArray(m.mean, (npix_all, n)), device=dev),
                           "m2": torch.as_tensor(_DevArray(m.m2, (npix_all, n)), device=dev)}
    else:
        film = film_mod.FilmPlanes(a.width, a.height, n, dev)
        drt_film = film.as_drt_film()
        root_planes = {"sum": film.sum, "filter": film.filter, "mean": film.mean, "m2": film.m2}

    def combine():
        if peers is not None:
            return peers.merge(stream.cuda_stream, stream.synchronize)
        return film_mod.merge_distributed_(film)
    stream = torch.cuda.current_stream()
    prm = common.structs.RenderParams(a.width, a.height, rank * a.spp, (rank + 1) * a.spp, a.depth, cfg.pixel_scheme, a.seed)
    paths_per_step = a.width * a.height * a.spp * world

    def render():
        if peers is not None:
            peers.render(prm, stream.cuda_stream)
        else:
            ctx.render_device(prm, drt_film, accumulate=False, stream=stream.cuda_stream)

    def step():
        render()
        return combine()

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    peak_tf = ctx.measure_fp32_peak(False) if rank == 0 else 0.0
    for _ in range(max(a.warmup, 3)):
        step()
    fence()
    sampler = ClockSampler(local) if rank == 0 else None
    k_start = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    k_stop = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(a.steps):
        k_start[i].record(stream)
        render()
        k_stop[i].record(stream)
        combine()
        launches += 1 + (1 if peers is not None else 0)
    e1.record(stream)
    fence()
    t1 = time.perf_counter()
    merge_ms = None
    if peers is not None:      # one more (untimed) step with the exchange's phases timed on the host
        peers.timing = []
        step()
        fence()
        merge_ms = {k: 1e3 * v for k, v in zip(("wait_own_render", "barrier_before", "merge_kernel", "barrier_after"), peers.timing[-1])}
        peers.timing = None
    ms_total = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    kernel_ms = sum(s.elapsed_time(e) for s, e in zip(k_start, k_stop)) / a.steps
    clocks = sampler.stop(t0, t1) if sampler else None
    st = ctx.stats()

    # ---- end to end through host buffers: scene upload (H2D) + render (+ NCCL merge) + film read-back (D2H), every step
    npix = a.width * a.height
    pinned = {k: torch.empty(shape, dtype=torch.float32).pin_memory()
              for k, shape in (("sum", (npix, n)), ("filter", (npix,)), ("mean", (npix, n)), ("m2", (npix, n)))} if rank == 0 else None
    host_film = cuda.Film(*(pinned[k].data_ptr() for k in ("sum", "filter", "mean", "m2"))) if rank == 0 else None

    def e2e_step():
        ctx.upload_scene(scene, camera, tables)
        if world == 1:
            ctx.render_host_into(prm, host_film)           # the C-ABI host-buffer call of include/drt_cuda.h
        else:
            render()
            combine()
            if rank == 0:
                for k in ("sum", "filter", "mean", "m2"):
                    pinned[k].copy_(root_planes[k], non_blocking=True)
            torch.cuda.synchronize()

    e2e_step()
    fence()
    te0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    fence()
    e2e_s = torch.tensor([time.perf_counter() - te0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    upload_bytes = ctx.scene_upload_bytes()
    d2h_bytes = (3 * npix * n + npix) * 4

    if rank == 0:
        kinfo = ctx.render_kernel_info(prm)
        flops_path, rc, rs, b = algorithmic_flops_per_path(scene, st, n)
        paths_launch = a.width * a.height * a.spp
        achieved_tf = flops_path * paths_launch / (kernel_ms * 1e-3) / 1e12
        film_bytes = (3 * npix * n + npix) * 4
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        value = paths_per_step * a.steps / (ms_total * 1e-3)
        traffic = None
        try:
            tj = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json")))
            if (a.width, a.height) == (1024, 1024):
                traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_written_per_launch"]
        except Exception:
            traffic = None
        line = {
            "metric": "camera_paths_per_sec", "value": value, "unit": "paths/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "impl": "b200",
            "data": f"shipped scene assets/scenes/{a.scene}.scn through the legacy-compat parser; per-path Philox4x32-10 streams, seed {a.seed}",
            "config": {"workload": workload_name(a), "geometry": a.geometry, "film_merge": merge_kind,
                       "l2": "inputs (scene + spectra, < 64 KB) live in shared memory; each step writes 4 fresh film planes "
                             f"({film_bytes / 1e6:.0f} MB > 126 MB L2), nothing is re-read between steps",
                       "samples_per_pixel_total": a.spp * world},
            "rays_per_sec": value * (rc + rs),
            "rays_per_path": {"closest": rc, "shadow": rs, "shaded_bounces": b},
            "clocks": clocks,
            "e2e": {"value": paths_per_step * a.steps / e2e_s, "unit": "paths/s", "h2d_bytes_per_step": upload_bytes,
                    "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches,
            "film_exchange_ms": merge_ms,
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture (profiles/ncu_traffic.json); null for other image sizes",
                         "kernel": kinfo[0], "warps_per_cta": kinfo[1], "ctas_per_sm": kinfo[2], "kernel_ms": kernel_ms,
                         "algorithmic_flops_per_path": flops_path,
                         "peak_source": "measured in this run by drt_cuda_measure_fp32_peak (FFMA, 2 flops); MEASURED_PEAKS.json has no FP32 entry",
                         "hbm": {"achieved": film_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": film_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                                 "note": "algorithmic HBM bytes per launch = one write of the four film planes"}},
        }
        if world == 1 and not a.no_cpu_baseline:
            try:
                arm = CpuArm(a)
                arm.step(0, rows=max(64, a.height // 8))
                paths, secs = 0, 0.0
                passes = 0
                while secs < 10.0 and passes < 256:   # a bounded sample: about 10 s of work on all host cores
                    p, t = arm.step(1 + passes)
                    paths += p
                    secs += t
                    passes += 1
                # image RMSE against the reference on the same per-path random streams: a central tile, 8 samples per pixel
                try:
                    import numpy as np
                    tw = min(64, a.width); th = min(64, a.height)
                    x0, y0 = (a.width - tw) // 2, (a.height - th) // 2
                    prm8 = common.structs.RenderParams(a.width, a.height, 0, 8, a.depth, cfg.pixel_scheme, a.seed)
                    ref_paths = arm.tile_paths(x0, y0, x0 + tw, y0 + th, 0, 8)
                    gpu_paths = ctx.sample_paths(prm8, x0, y0, x0 + tw, y0 + th)
                    ref_mean, gpu_mean = ref_paths.mean(axis=1), gpu_paths.astype(np.float64).mean(axis=1)
                    rmse = float(np.sqrt(np.mean((gpu_mean - ref_mean) ** 2)))
                    perr = common.path_errors(gpu_paths, ref_paths)
                    line["image_rmse"] = {"value": rmse, "relative": rmse / float(np.abs(ref_mean).mean()),
                                          "unit": "spectral radiance, RMSE over pixels and wavelengths of the 8-sample mean",
                                          "tile": f"{tw}x{th} at the image centre, samples 0-7, {ref_paths.shape[0] * 8} paths",
                                          "rng": "matched: both sides draw the same per-path Philox streams",
                                          "paths_within_1e-3": float((perr <= 1e-3).mean()), "reference": arm.kind}
                except Exception as exc:
                    line["image_rmse"] = {"value": None, "error": repr(exc)}
                arm.close()
                line["cpu_baseline"] = {"value": paths / secs, "unit": "paths/s", "cores": arm.cores, "kind": arm.kind,
                                        "sample": f"{passes} full-frame passes of {a.width}x{a.height} at 1 sample per pixel ({paths} paths, {secs:.1f} s), "
                                                  f"one single-threaded process per core"}
            except Exception as exc:   # the baseline is reported, never a gate
                line["cpu_baseline"] = {"value": None, "unit": "paths/s", "cores": _cpu_cores(), "kind": "unavailable", "sample": repr(exc)}
        print(json.dumps(line), flush=True)
    if peers is not None:
        peers.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()

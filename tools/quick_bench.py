"""Render-kernel throughput of named scenes on one GPU: python tools/quick_bench.py [--spp S] [--size W] [--depth D] [--f64] scene ...
(no scene = the three shipped Cornell boxes; prints the FP32 probe first)."""
import argparse, importlib, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import common, oracledriver, torch
cuda = importlib.import_module("daily-ray-trace_b200.cuda")
ap = argparse.ArgumentParser()
ap.add_argument("scenes", nargs="*", default=["init_cornell", "cornell_plane_light", "cornell_large_box"])
ap.add_argument("--spp", type=int, default=64); ap.add_argument("--size", type=int, default=1024); ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--f64", action="store_true"); ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
ctx = cuda.Context(0)
print("fp32 peak TF", round(ctx.measure_fp32_peak(False), 1), "packed", round(ctx.measure_fp32_peak(True), 1))
for scene in a.scenes:
    w = h = a.size
    cfg, tables, sc, cam = common.load(scene, w, h, a.spp, a.depth)
    ctx.upload_scene(sc, cam, tables); n = sc.num_wavelengths
    ctx.set_geometry_precision(1 if a.f64 else 0)
    planes = [torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h, device="cuda"), torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h*n, device="cuda")]
    film = cuda.film_from_tensors(*planes); prm = oracledriver.params(w, h, 0, a.spp, a.depth, 2, 1)
    ctx.render_device(prm, film); torch.cuda.synchronize()
    best = 1e30
    for _ in range(a.reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ctx.render_device(prm, film); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    st = ctx.stats(); k = ctx.render_kernel_info(prm)
    print(f"{scene} {w}x{h}x{a.spp} d{a.depth}: {best:.2f} ms {w*h*a.spp/best/1e3:.0f} Mpaths/s rays/path {(st.closest_rays+st.shadow_rays)/st.paths:.2f} "
          f"bounces/path {st.shaded_bounces/st.paths:.2f} {k[0]} {k[1]}w x {k[2]}cta", flush=True)

import importlib, sys, time, ctypes as C
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
sys.path.insert(0, os.path.join(HERE, ".."))
import common, oracledriver
import torch
cuda = importlib.import_module("daily-ray-trace_b200.cuda")
ctx = cuda.Context(0)
print("fp32 peak TF", ctx.measure_fp32_peak(False), "packed", ctx.measure_fp32_peak(True))
for scene, w, h, spp in [("cornell_plane_light", 1024, 1024, 64), ("init_cornell", 1024, 1024, 64), ("cornell_large_box", 1024, 1024, 64)]:
    cfg, tables, sc, cam = common.load(scene, w, h, spp, 4)
    ctx.upload_scene(sc, cam, tables)
    n = sc.num_wavelengths
    planes = [torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h, device="cuda"), torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h*n, device="cuda")]
    film = cuda.film_from_tensors(*planes)
    for geo in (0, 1):
        ctx.set_geometry_precision(geo)
        prm = oracledriver.params(w, h, 0, spp, 4, 2, 1)
        ctx.render_device(prm, film); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctx.render_device(prm, film); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        st = ctx.stats()
        print(f"{scene} geo={'f64' if geo else 'f32'} {w}x{h}x{spp}: {ms:.2f} ms  {w*h*spp/ms/1e3:.1f} Mpaths/s  rays/path {(st.closest_rays+st.shadow_rays)/st.paths:.2f}", flush=True)

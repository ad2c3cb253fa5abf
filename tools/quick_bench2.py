import importlib, sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import common, oracledriver, torch
cuda = importlib.import_module("daily-ray-trace_b200.cuda")
ctx = cuda.Context(0)
for scene in sys.argv[1:]:
    w = h = 1024; spp = 64
    cfg, tables, sc, cam = common.load(scene, w, h, spp, 4)
    ctx.upload_scene(sc, cam, tables); n = sc.num_wavelengths
    planes = [torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h, device="cuda"), torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h*n, device="cuda")]
    film = cuda.film_from_tensors(*planes); prm = oracledriver.params(w, h, 0, spp, 4, 2, 1)
    ctx.render_device(prm, film); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ctx.render_device(prm, film); b.record(); torch.cuda.synchronize()
    st = ctx.stats()
    print(f"{scene}: {a.elapsed_time(b):.2f} ms {w*h*spp/a.elapsed_time(b)/1e3:.0f} Mpaths/s rays/path {(st.closest_rays+st.shadow_rays)/st.paths:.2f} bounces/path {st.shaded_bounces/st.paths:.2f}", flush=True)

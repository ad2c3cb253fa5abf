"""Error anatomy of the stress scene: f32 vs f64 geometry, by depth."""
import importlib, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import common, oracledriver
cuda = importlib.import_module("daily-ray-trace_b200.cuda")
ctx = cuda.Context(0)
w, h, spp = 64, 48, 5
for depth in (1, 2, 3, 4, 6):
    cfg, tables, sc, cam = common.load("stress_all", w, h, spp, depth)
    ctx.upload_scene(sc, cam, tables)
    prm = oracledriver.params(w, h, 0, spp, depth, 2, 0xC0FFEE)
    _, _, _, ref, cnt = oracledriver.render_tile(sc, cam, prm, 0, 0, w, h, want_paths=True)
    for geo in (0, 1):
        ctx.set_geometry_precision(geo)
        gpu = ctx.sample_paths(prm, 0, 0, w, h)
        err = common.path_errors(gpu, ref)
        q = np.quantile(err, [0.5, 0.9, 0.99, 0.999])
        print(f"depth {depth} geo={'f64' if geo else 'f32'}: <=1e-3 {100*(err<=1e-3).mean():.3f}%  <=1e-4 {100*(err<=1e-4).mean():.3f}%  q50/90/99/99.9 = " + " ".join(f"{x:.1e}" for x in q) + f" max {err.max():.1e}")

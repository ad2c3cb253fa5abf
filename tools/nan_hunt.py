"""Find NaN/inf film pixels in a large render and trace them to the sample and to the oracle's value for it."""
import importlib, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
import common, oracledriver
cuda = importlib.import_module("daily-ray-trace_b200.cuda")
film_mod = importlib.import_module("daily-ray-trace_b200.film")
scene_name = sys.argv[1] if len(sys.argv) > 1 else "cornell_plane_light"
w = h = 1024; spp = 64; depth = 4
ctx = cuda.Context(0)
cfg, tables, scene, camera = common.load(scene_name, w, h, spp, depth)
ctx.upload_scene(scene, camera, tables)
n = scene.num_wavelengths
for seed, s0 in ((5, 0), (6, 0)):
    for geo in (0,):
        ctx.set_geometry_precision(geo)
        film = film_mod.FilmPlanes(w, h, n, torch.device("cuda", 0))
        prm = common.structs.RenderParams(w, h, s0, s0 + spp, depth, cfg.pixel_scheme, seed)
        ctx.render_device(prm, film.as_drt_film()); torch.cuda.synchronize()
        bad = (~torch.isfinite(film.sum)).any(dim=1).nonzero().flatten().cpu().numpy()
        print(f"seed {seed} samples [{s0},{s0+spp}) geo {'f64' if geo else 'f32'}: {len(bad)} non-finite pixels", bad[:8])
        for p in bad[:3]:
            x, y = int(p % w), int(p // w)
            paths = ctx.sample_paths(prm, x, y, x + 1, y + 1)[0]
            which = np.where(~np.isfinite(paths).all(axis=1))[0]
            for s in which[:2]:
                o = oracledriver.sample(scene, camera, oracledriver.params(w, h, s0, s0 + spp, depth, cfg.pixel_scheme, seed), x, y, s0 + int(s))
                badl = np.where(~np.isfinite(paths[s]))[0]
                good = np.isfinite(paths[s])
                rel = np.abs(paths[s][good] - o[good]) / np.maximum(np.abs(o[good]), 1e-9)
                recs = ctx.debug_records(prm, x, y, x + 1, y + 1)[0][int(s)]
                nbv = int(recs[0]); ew = 2; bw = 2 + 1 * (ew + 1) + ew
                print("      record: nb", nbv, "vig", recs[1:2].view(np.float32))
                for b in range(nbv):
                    r = recs[2 + b * bw: 2 + (b + 1) * bw]
                    print(f"        bounce {b}: hdr 0x{int(r[0]):08x} on_dot {r[1:2].view(np.float32)[0]:.7g} words {r[2:].view(np.float32)}")
                print(f"   pixel ({x},{y}) sample {s0+int(s)}: non-finite wavelengths {badl.tolist()} values {paths[s][badl][:4]}; finite part max rel err vs oracle {rel.max():.2e}; oracle at bad {o[badl][:4]}")
        del film

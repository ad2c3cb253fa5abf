"""Render time against samples per pixel, and per single sample index (python tools/spp_sweep.py scene [W]).
A time that does not go to zero with spp, or one slow sample index, points at a few pathological paths (a rejection
sampler that spins) rather than at throughput."""
import importlib, sys, os, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import common, oracledriver, torch
cuda = importlib.import_module("daily-ray-trace_b200.cuda")
scene = sys.argv[1]; w = h = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ctx = cuda.Context(0)
cfg, tables, sc, cam = common.load(scene, w, h, 64, 4)
ctx.upload_scene(sc, cam, tables); n = sc.num_wavelengths
planes = [torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h, device="cuda"), torch.zeros(w*h*n, device="cuda"), torch.zeros(w*h*n, device="cuda")]
film = cuda.film_from_tensors(*planes)
def run(s0, s1):
    prm = oracledriver.params(w, h, s0, s1, 4, 2, 1)
    ctx.render_device(prm, film); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ctx.render_device(prm, film); b.record(); torch.cuda.synchronize()
    st = ctx.stats()
    return a.elapsed_time(b), st.rng_draws / st.paths
for spp in (1, 2, 4, 8, 16, 32, 64, 128, 256, 1024):
    t, d = run(0, spp)
    print(f"{scene} spp {spp:5d}: {t:8.2f} ms  {w*h*spp/t/1e3:8.0f} Mpaths/s  draws/path {d:.3f}", flush=True)
for s in range(0, 8):
    t, d = run(s, s + 1)
    print(f"{scene} sample {s}: {t:8.2f} ms draws/path {d:.3f}", flush=True)

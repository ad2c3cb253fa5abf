"""Generates tests/golden/converged/*.npz: film means written by the VERBATIM reference binary (oracle/_ref/raytrace_ref =
unmodified win32_main.c, one sequential glibc rand() stream) for small full frames at a few hundred samples per pixel.

Run here, where /root/reference exists:   python tests/golden/make_converged.py
The fixtures are the "image" side of the parity contract (BASELINE.json: image RMSE vs reference): a render of the same
scene with INDEPENDENT random streams (the per-path Philox streams of the CUDA path or of the oracle) must agree with
them within the Monte Carlo noise bound.  The reference normalises its variance file per pixel (daily_ray_trace.c:766-769),
so the noise estimate comes from the renderer under test."""
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import common      # noqa: E402
import refdriver   # noqa: E402

host = importlib.import_module("daily-ray-trace_b200.host")

# name, W, H, spp, depth
CASES = [("init_cornell", 40, 40, 1024, 4), ("cornell_plane_light", 40, 40, 1024, 4), ("cornell_large_box", 32, 32, 1024, 4)]


def main():
    assert os.path.exists(refdriver.REF_BIN), "build oracle/_ref first (make -C oracle ref)"
    for name, w, h, spp, depth in CASES:
        parsed = host.parse_scene_text(open(common.scene_path(name)).read())
        upgraded = host.scene_to_text(parsed)
        with tempfile.TemporaryDirectory() as root:
            refdriver.make_root(root, common.ASSETS, upgraded, "upgraded.scn")
            cfg = host.make_config_text(scene="scenes\\upgraded.scn", width=w, height=h, spp=spp, depth=depth)
            open(os.path.join(root, "config.cfg"), "w").write(cfg)
            subprocess.run([refdriver.REF_BIN], cwd=root, check=True, stdout=subprocess.DEVNULL)
            raw = open(os.path.join(root, "output", "average.spd"), "rb").read()
            n = (len(raw) - 40) // (8 * w * h)
            mean = np.frombuffer(raw[40:], dtype=np.float64).reshape(w * h, n)
            np.savez_compressed(os.path.join(HERE, "converged", f"{name}.npz"), meta=np.array([w, h, spp, depth], dtype=np.int64),
                                mean=mean.astype(np.float32))
            print(name, mean.shape, "mean radiance", float(mean.mean()))


if __name__ == "__main__":
    main()

"""GPU tier: the Linux main (host/drt_main.c, the replacement of win32_main.c) end to end: run it in a directory laid out like
the reference's (config.cfg, scenes/, spectra/, output/) and compare its three .spd files with the library path."""
import importlib
import os
import subprocess

import numpy as np
import pytest

import common
import oracledriver
import refdriver

cuda = importlib.import_module("daily-ray-trace_b200.cuda")
pytestmark = pytest.mark.gpu

BIN = os.path.join(common.REPO, "daily-ray-trace_b200", "drt_raytrace")


@pytest.mark.parametrize("scene", ["cornell_plane_light", "init_cornell"])
def test_linux_main_writes_reference_format(host, tmp_path, scene):
    if not os.path.exists(BIN):
        pytest.fail("drt_raytrace is not built (make -C daily-ray-trace_b200)")
    root = refdriver.make_root(str(tmp_path), common.ASSETS)
    w, h, spp = 40, 24, 48
    cfg_text = host.make_config_text(scene=f"scenes\\{scene}.scn", width=w, height=h, spp=spp)
    open(os.path.join(root, "config.cfg"), "w").write(cfg_text)
    out = subprocess.run([BIN, "--seed", "11"], cwd=root, check=True, capture_output=True, text=True).stdout
    assert "Render complete." in out and "Total render time" in out and "Converted." in out
    n = 69
    files = {k: open(os.path.join(root, "output", k), "rb").read() for k in
             ("output.spd", "average.spd", "variance.spd", "output.bmp", "average.bmp", "variance.bmp")}
    assert len(files["output.spd"]) == 40 + w * h * (n + 1) * 8 and len(files["average.spd"]) == 40 + w * h * n * 8
    assert len(files["output.bmp"]) == 54 + w * h * 4 and files["output.bmp"][:2] == b"BM"
    hdr = np.frombuffer(files["output.spd"][:20], dtype=np.uint32)
    assert list(hdr) == [0xedfeefbe, w, h, n, 1]

    # the same render through the library from Python
    cfg, tables, sc, cam = common.load(scene, w, h, spp, 4)
    ctx = cuda.Context(0)
    ctx.upload_scene(sc, cam, tables)
    film = ctx.render_host(oracledriver.params(w, h, 0, spp, 4, cfg.pixel_scheme, 11))
    ctx.close()
    body = np.frombuffer(files["output.spd"][40:], dtype=np.float64).reshape(w * h, n + 1)
    assert np.array_equal(body[:, n], np.full(w * h, float(spp)))
    assert np.array_equal(body[:, :n].astype(np.float32), film["sum"])
    mean = np.frombuffer(files["average.spd"][40:], dtype=np.float64).reshape(w * h, n)
    assert np.array_equal(mean.astype(np.float32), film["mean"])
    var = np.frombuffer(files["variance.spd"][40:], dtype=np.float64).reshape(w * h, n)
    peak = film["m2"].astype(np.float64).max(axis=1, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        expect = film["m2"].astype(np.float64) / peak
    same = (var == expect) | (np.isnan(var) & np.isnan(expect))
    assert same.all()


@pytest.mark.parametrize("gpus", [2, 3])
def test_linux_main_on_several_gpus(host, tmp_path, gpus):
    """`drt_raytrace --gpus G` (drt_cuda_render_host_multi: samples split over the devices of one process, scattered exchange without
    IPC) writes the same films as the single-device run: identical sample sets, partial films merged in another order."""
    import torch
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    root = refdriver.make_root(str(tmp_path), common.ASSETS)
    w, h, spp, n = 37, 23, 96, 69
    open(os.path.join(root, "config.cfg"), "w").write(host.make_config_text(scene="scenes\\cornell_plane_light.scn", width=w, height=h, spp=spp))

    def run(args):
        out = subprocess.run([BIN, "--seed", "5"] + args, cwd=root, check=True, capture_output=True, text=True).stdout
        assert "Render complete." in out
        raw = {k: open(os.path.join(root, "output", k + ".spd"), "rb").read() for k in ("output", "average", "variance")}
        return (np.frombuffer(raw["output"][40:], dtype=np.float64).reshape(w * h, n + 1).copy(),
                np.frombuffer(raw["average"][40:], dtype=np.float64).reshape(w * h, n).copy(), out)

    one_sum, one_mean, _ = run([])
    many_sum, many_mean, log = run(["--gpus", str(gpus)])
    assert f"Camera paths: {w * h * spp} " in log
    assert np.array_equal(many_sum[:, n], one_sum[:, n])
    for a, b in ((many_sum[:, :n], one_sum[:, :n]), (many_mean, one_mean)):
        floor = 1e-5 * np.abs(b).max()
        assert (np.abs(a - b) / np.maximum(np.abs(b), floor)).max() < 2e-4


def test_more_devices_than_samples(host, tmp_path):
    """`--gpus 2` with ONE sample per pixel: the samples cannot be split, the surplus device idles (drt_cuda_render_host_multi falls back
    to the devices that have work) and the films equal the single-device run bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = refdriver.make_root(str(tmp_path), common.ASSETS)
    w, h, n = 24, 16, 69
    open(os.path.join(root, "config.cfg"), "w").write(host.make_config_text(scene="scenes\\cornell_plane_light.scn", width=w, height=h, spp=1))

    def run(args):
        subprocess.run([BIN, "--seed", "5"] + args, cwd=root, check=True, capture_output=True, text=True)
        return {k: open(os.path.join(root, "output", k), "rb").read() for k in ("output.spd", "average.spd", "output.bmp")}

    assert run([]) == run(["--gpus", "2"])


@pytest.mark.parametrize("spp,depth", [(0, 4), (3, 0)])
def test_main_with_no_samples_or_no_depth(host, tmp_path, spp, depth):
    """num_pixel_samples 0 / max_cast_depth 0 are valid configurations of the reference (its loops simply do not run): black films,
    filter = sample count, not an error."""
    root = refdriver.make_root(str(tmp_path), common.ASSETS)
    w, h, n = 16, 12, 69
    open(os.path.join(root, "config.cfg"), "w").write(host.make_config_text(scene="scenes\\cornell_plane_light.scn", width=w, height=h, spp=spp, depth=depth))
    out = subprocess.run([BIN], cwd=root, check=True, capture_output=True, text=True).stdout
    assert "Render complete." in out
    body = np.frombuffer(open(os.path.join(root, "output", "output.spd"), "rb").read()[40:], dtype=np.float64).reshape(w * h, n + 1)
    assert not body[:, :n].any() and (body[:, n] == spp).all()

"""Image-level parity (BASELINE.json: "image RMSE vs reference"): a render with INDEPENDENT random streams must agree with the
film mean written by the verbatim reference binary (tests/golden/converged/*.npz, made by tests/golden/make_converged.py
from the unmodified win32_main.c with its one sequential glibc rand() stream) within the Monte Carlo noise bound.

Tolerance: RMSE(mean - reference mean) <= 1.5 * sigma, sigma^2 = mean over pixels and wavelengths of
var * (1/K + 1/K_ref) with var = M2 / (K - 1) from the renderer under test (the reference normalises its variance file per
pixel, daily_ray_trace.c:766-769).  An unbiased renderer sits at RMSE / sigma = 1 (0.77 .. 1.05 measured for the oracle); the
summed radiance of the whole frame must agree within 4 sigma of its own noise.
CPU tier: the oracle restatement's film path.  GPU tier: the CUDA path through the host-buffer C-ABI call."""
import importlib
import os

import numpy as np
import pytest

import common
import oracledriver

CASES = ["init_cornell", "cornell_plane_light", "cornell_large_box"]


def _check(name, k, mean, m2, ref_mean, k_ref):
    var = m2.astype(np.float64) / (k - 1)
    sig2 = var * (1.0 / k + 1.0 / k_ref)
    d = mean.astype(np.float64) - ref_mean
    rmse, sigma = float(np.sqrt(np.mean(d * d))), float(np.sqrt(np.mean(sig2)))
    # the wavelengths of a pixel come from the same paths (fully correlated), pixels are independent
    total_err, total_sigma = float(abs(d.sum())), float(np.sqrt((np.sqrt(sig2).sum(axis=1) ** 2).sum()))
    print(f"\n{name} K={k} vs K_ref={k_ref}: RMSE {rmse:.4g}, noise bound sigma {sigma:.4g}, ratio {rmse / sigma:.3f}, "
          f"relative RMSE {rmse / ref_mean.mean():.4f}; frame total off by {total_err / total_sigma:.2f} sigma")
    assert np.isfinite(mean).all() and np.isfinite(m2).all()
    assert rmse <= 1.5 * sigma
    assert total_err <= 4.0 * total_sigma


def _fixture(name):
    fx = np.load(os.path.join(common.GOLDEN, "converged", f"{name}.npz"))
    w, h, k_ref, depth = (int(v) for v in fx["meta"])
    return w, h, k_ref, depth, fx["mean"].astype(np.float64)


@pytest.mark.parametrize("name", CASES)
def test_oracle_image_within_noise_of_reference(name):
    w, h, k_ref, depth, ref_mean = _fixture(name)
    k = 256
    cfg, tables, scene, camera = common.load(name, w, h, k, depth)
    prm = oracledriver.params(w, h, 0, k, depth, cfg.pixel_scheme, 4242)
    _, avg, m2, _, _ = oracledriver.render_tile(scene, camera, prm, 0, 0, w, h)
    _check(name, k, avg, m2, ref_mean, k_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [256, 4096])
@pytest.mark.parametrize("name", CASES)
def test_cuda_image_within_noise_of_reference(name, k):
    cuda = importlib.import_module("daily-ray-trace_b200.cuda")
    w, h, k_ref, depth, ref_mean = _fixture(name)
    cfg, tables, scene, camera = common.load(name, w, h, k, depth)
    ctx = cuda.Context(0)
    try:
        ctx.upload_scene(scene, camera, tables)
        film = ctx.render_host(oracledriver.params(w, h, 0, k, depth, cfg.pixel_scheme, 777))
    finally:
        ctx.close()
    assert np.all(film["filter"] == k)
    _check(name, k, film["mean"], film["m2"], ref_mean, k_ref)

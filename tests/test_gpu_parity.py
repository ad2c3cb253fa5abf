"""GPU parity (the -m gpu tier): the CUDA path through the C ABI against the CPU oracle, RNG-matched per path.

Tolerances: per (pixel, sample) the max-over-wavelength relative error must be <= 1e-3 (absolute floor
1e-6 x the largest radiance) for >= 99.5 % of paths with f32 geometry; the rest are discrete-branch flips (a ray that
lands the other side of an edge in f32) and are counted.  With f64 geometry the flips disappear, which is the evidence
that they are branch flips and not arithmetic error."""
import importlib

import numpy as np
import pytest

import common
import oracledriver

cuda = importlib.import_module("daily-ray-trace_b200.cuda")

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3
CASES = [
    # scene, W, H, tile (x0,y0,x1,y1), spp, depth, scheme
    ("cornell_plane_light", 64, 64, (0, 0, 64, 64), 4, 4, "pixel_random"),
    ("init_cornell", 64, 48, (0, 0, 64, 48), 3, 4, "pixel_random"),
    ("cornell_large_box", 48, 48, (8, 8, 40, 40), 33, 4, "pixel_random"),
    ("cornell_downward", 48, 48, (0, 0, 48, 48), 2, 4, "pixel_random"),
    ("first_scene", 32, 32, (0, 0, 32, 32), 2, 4, "pixel_center"),
    ("example_scene", 32, 32, (0, 0, 32, 32), 1, 3, "pixel_random"),
    ("stress_all", 64, 48, (0, 0, 64, 48), 5, 6, "pixel_random"),
    ("rotated_room", 64, 48, (0, 0, 64, 48), 4, 5, "pixel_random"),     # all-plastic kernel on planes in general position
    ("sky_cornell", 64, 48, (0, 0, 64, 48), 3, 4, "pixel_random"),      # pinhole + emissive escape material (Q19): no pixel may be culled
    ("classed_all", 64, 48, (0, 0, 64, 48), 6, 6, "pixel_random"),      # classed kernel: plastics, mirror, fs_conductor, glass, transmit, ct_conductor
]


@pytest.fixture(scope="module")
def ctx():
    c = cuda.Context(0)
    yield c
    c.close()


def _run(ctx, scene_name, w, h, tile, spp, depth, scheme, geometry, seed=0xC0FFEE):
    cfg, tables, scene, camera = common.load(scene_name, w, h, spp, depth, scheme)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(geometry)
    prm = oracledriver.params(w, h, 0, spp, depth, cfg.pixel_scheme, seed)
    x0, y0, x1, y1 = tile
    gpu = ctx.sample_paths(prm, x0, y0, x1, y1)
    st = ctx.stats()
    o_sum, o_avg, o_m2, o_paths, cnt = oracledriver.render_tile(scene, camera, prm, x0, y0, x1, y1, want_paths=True)
    return gpu, o_paths, st, cnt


@pytest.mark.parametrize("scene,w,h,tile,spp,depth,scheme", CASES)
def test_per_path_radiance_f32(ctx, scene, w, h, tile, spp, depth, scheme):
    gpu, ref, st, cnt = _run(ctx, scene, w, h, tile, spp, depth, scheme, cuda.GEOMETRY_F32)
    err = common.path_errors(gpu, ref)
    ok = (err <= REL_TOL).mean()
    print(f"\n{scene}: {err.size} paths, within {REL_TOL:g}: {100 * ok:.3f} %, median err {np.median(err):.2e}, "
          f"p99 {np.quantile(err, 0.99):.2e}, mismatching {int((err > REL_TOL).sum())}")
    assert ok >= 0.995
    assert st.paths == cnt.paths == err.size
    # work counters agree with the oracle up to the few flipped paths
    for a, b in ((st.closest_rays, cnt.closest_rays), (st.shadow_rays, cnt.shadow_rays), (st.rng_draws, cnt.rng_draws)):
        assert abs(a - b) <= 0.01 * b + 8


@pytest.mark.parametrize("scene,w,h,tile,spp,depth,scheme", CASES)
def test_per_path_radiance_f64_geometry(ctx, scene, w, h, tile, spp, depth, scheme):
    gpu, ref, st, cnt = _run(ctx, scene, w, h, tile, spp, depth, scheme, cuda.GEOMETRY_F64)
    err = common.path_errors(gpu, ref)
    ok = (err <= 1e-4).mean()
    print(f"\n{scene} [f64 geometry]: within 1e-4: {100 * ok:.4f} %, max err {err.max():.2e}")
    assert ok >= 0.9995
    assert (st.closest_rays, st.shadow_rays, st.shaded_bounces) == (cnt.closest_rays, cnt.shadow_rays, cnt.shaded_bounces) or ok < 1.0


def test_film_matches_oracle(ctx):
    """Film planes (sum, filter, Welford mean and M2) through the host-buffer entry point."""
    w, h, spp, depth = 48, 40, 37, 4
    cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, spp, depth)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F64)
    prm = oracledriver.params(w, h, 0, spp, depth, cfg.pixel_scheme, 99)
    film = ctx.render_host(prm)
    o_sum, o_avg, o_m2, _, cnt = oracledriver.render_tile(scene, camera, prm, 0, 0, w, h)
    n = scene.num_wavelengths
    assert np.array_equal(film["filter"], o_sum[:, n].astype(np.float32))
    scale = np.abs(o_sum[:, :n]).max()
    bad_px = np.zeros(w * h, bool)
    for name, ref in (("sum", o_sum[:, :n]), ("mean", o_avg), ("m2", o_m2)):
        floor = 1e-5 * np.abs(ref).max()
        rel = np.abs(film[name] - ref) / np.maximum(np.abs(ref), floor)
        bad_px |= rel.max(axis=1) > 2e-3
        print(f"\nfilm {name}: max rel {rel.max():.2e}, pixels over 2e-3: {(rel.max(axis=1) > 2e-3).sum()}")
    assert bad_px.mean() <= 0.002
    assert scale > 0


@pytest.mark.parametrize("scene,w,h,spp", [("init_cornell", 64, 48, 33), ("init_cornell", 40, 56, 5), ("cornell_downward", 48, 48, 32),
                                           ("first_scene", 40, 30, 34), ("rotated_room", 56, 40, 32), ("rotated_room", 33, 41, 7),
                                           ("sky_cornell", 64, 48, 33), ("sky_cornell", 40, 56, 5)])
def test_full_frame_with_unseen_pixels(ctx, scene, w, h, spp):
    """Whole frames of scenes that fill only part of the image: pixels outside the screen-space bound of the scene are counted, not
    traced (RenderLaunch::hit_*), in both task shapes.  The film and every work counter must equal the oracle's, which traces them."""
    depth = 4
    cfg, tables, sc, camera = common.load(scene, w, h, spp, depth)
    ctx.upload_scene(sc, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F32)
    prm = oracledriver.params(w, h, 0, spp, depth, cfg.pixel_scheme, 5)
    film = ctx.render_host(prm)
    st = ctx.stats()
    o_sum, o_avg, o_m2, _, cnt = oracledriver.render_tile(sc, camera, prm, 0, 0, w, h)
    n = sc.num_wavelengths
    assert np.array_equal(film["filter"], o_sum[:, n].astype(np.float32))
    lit_ref, lit_gpu = np.abs(o_sum[:, :n]).max(axis=1) > 0, np.abs(film["sum"]).max(axis=1) > 0
    print(f"\n{scene} {w}x{h}x{spp}: {100 * (1 - lit_ref.mean()):.1f} % of the pixels see nothing")
    if scene == "sky_cornell":      # the sky lights every pixel, also those that see no surface (31 % of this frame)
        assert lit_ref.all()
    assert np.array_equal(lit_ref, lit_gpu)                      # no pixel that receives light was skipped, none was invented
    bad = np.zeros(w * h, bool)
    for name, ref in (("sum", o_sum[:, :n]), ("mean", o_avg), ("m2", o_m2)):
        floor = 1e-5 * np.abs(ref).max()
        bad |= (np.abs(film[name] - ref) / np.maximum(np.abs(ref), floor)).max(axis=1) > 2e-3
    # one path that takes another branch in f32 (an edge hit; at most 0.5 % of the paths, see the per-path tests) moves its whole pixel
    assert bad.mean() <= 0.02
    assert st.paths == cnt.paths == w * h * spp
    assert abs(int(st.closest_rays) - int(cnt.closest_rays)) <= 4 * max(int(bad.sum()), 1) * spp
    assert int(st.rng_draws) >= 2 * w * h * spp


def test_sample_ranges_compose(ctx):
    """Rendering [0,a) then accumulating [a,b) equals rendering [0,b): same per-path streams; the partial films are merged in a
    different order (pairwise update of count/mean/M2), so planes agree to rounding, the sample count exactly."""
    w, h, depth = 32, 32, 4
    cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, 48, depth)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F32)
    import torch
    n = scene.num_wavelengths

    def planes():
        return [torch.zeros(w * h * n, device="cuda"), torch.zeros(w * h, device="cuda"),
                torch.zeros(w * h * n, device="cuda"), torch.zeros(w * h * n, device="cuda")]

    a, b = planes(), planes()
    ctx.render_device(oracledriver.params(w, h, 0, 48, depth, cfg.pixel_scheme, 5), cuda.film_from_tensors(*a))
    ctx.render_device(oracledriver.params(w, h, 0, 20, depth, cfg.pixel_scheme, 5), cuda.film_from_tensors(*b))
    ctx.render_device(oracledriver.params(w, h, 20, 48, depth, cfg.pixel_scheme, 5), cuda.film_from_tensors(*b), accumulate=True)
    torch.cuda.synchronize()
    assert torch.equal(a[1], b[1])
    for x, y in zip((a[0], a[2], a[3]), (b[0], b[2], b[3])):
        scale = float(x.abs().max())
        assert float((x - y).abs().max()) <= 2e-5 * scale

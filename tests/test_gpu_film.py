"""GPU tier: film epilogues (K7 SPD->XYZ->RGB->u8, Chan merge), ragged shapes, the 4096^2 configuration, accumulate mode."""
import ctypes as C
import importlib

import numpy as np
import pytest

import common
import oracledriver

cuda = importlib.import_module("daily-ray-trace_b200.cuda")
film_mod = importlib.import_module("daily-ray-trace_b200.film")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cuda.Context(0)
    yield c
    c.close()


def _rel(a, b, floor_frac=1e-5):
    floor = floor_frac * max(float(np.abs(b).max()), 1e-30)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


@pytest.mark.parametrize("w,h,spp", [(1, 1, 1), (7, 3, 5), (33, 17, 31), (16, 16, 32), (9, 5, 67), (13, 11, 100)])
def test_ragged_shapes_match_oracle(ctx, w, h, spp):
    """Image sizes and sample counts that are not multiples of the warp batch (32): pixels per task, partial batches."""
    cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, spp, 4)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F64)
    prm = oracledriver.params(w, h, 0, spp, 4, cfg.pixel_scheme, 1234)
    film = ctx.render_host(prm)
    gpu_paths = ctx.sample_paths(prm, 0, 0, w, h)
    o_sum, o_avg, o_m2, o_paths, cnt = oracledriver.render_tile(scene, camera, prm, 0, 0, w, h, want_paths=True)
    n = scene.num_wavelengths
    assert ctx.stats().paths == w * h * spp
    assert np.array_equal(film["filter"], np.full(w * h, spp, np.float32))
    assert (common.path_errors(gpu_paths, o_paths) <= 1e-4).mean() >= 0.995
    for name, ref in (("sum", o_sum[:, :n]), ("mean", o_avg), ("m2", o_m2)):
        assert (_rel(film[name], ref).max(axis=1) <= 2e-3).mean() >= 0.99, name


def test_film_to_rgb_matches_host_conversion(ctx, host):
    import torch
    w, h, spp = 40, 24, 16
    cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, spp, 4)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F32)
    n = scene.num_wavelengths
    film = film_mod.FilmPlanes(w, h, n, torch.device("cuda", 0))
    prm = oracledriver.params(w, h, 0, spp, 4, cfg.pixel_scheme, 5)
    ctx.render_device(prm, film.as_drt_film())
    rgb = torch.zeros(w * h, 3, device="cuda")
    bgra = torch.zeros(w * h, dtype=torch.int32, device="cuda")
    L = host.lib()
    dp = C.POINTER(C.c_double)
    planes = {0: (film.sum / film.filter.unsqueeze(1)), 1: film.mean, 2: film.m2 / film.m2.max(dim=1, keepdim=True).values}
    for which in (0, 1, 2):
        ctx.film_to_rgb(film.as_drt_film(), w, h, which, rgb.data_ptr(), bgra.data_ptr())
        torch.cuda.synchronize()
        src = planes[which].double().cpu().numpy()
        got_rgb, got_q = rgb.cpu().numpy(), bgra.cpu().numpy().view(np.uint32)
        worst = 0
        for p in range(0, w * h, 7):
            spd = np.ascontiguousarray(src[p])
            ref = np.zeros(3)
            L.drt_spectrum_to_rgb(C.byref(tables), spd.ctypes.data_as(dp), ref.ctypes.data_as(dp))
            if np.isnan(ref).any():
                assert got_q[p] == 0          # 0/0 variance pixel -> NaN -> byte 0 (Q17)
                continue
            assert np.allclose(got_rgb[p], ref, rtol=2e-4, atol=2e-6), (which, p)
            q = L.drt_rgb_to_bgra8(ref.ctypes.data_as(dp))
            diff = max(abs(int((got_q[p] >> s) & 255) - int((q >> s) & 255)) for s in (0, 8, 16))
            worst = max(worst, diff)
        assert worst <= 1            # truncating quantiser: f32 vs f64 may differ by one code at a boundary


def test_film_merge_kernel_matches_single_render(ctx):
    """Two sample ranges rendered separately and merged on the device (the pairwise step of multi-GPU sharding)."""
    import torch
    w, h, depth = 32, 24, 4
    cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, 64, depth)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F32)
    n = scene.num_wavelengths
    dev = torch.device("cuda", 0)
    whole, a, b = (film_mod.FilmPlanes(w, h, n, dev) for _ in range(3))
    ctx.render_device(oracledriver.params(w, h, 0, 64, depth, cfg.pixel_scheme, 3), whole.as_drt_film())
    ctx.render_device(oracledriver.params(w, h, 0, 24, depth, cfg.pixel_scheme, 3), a.as_drt_film())
    ctx.render_device(oracledriver.params(w, h, 24, 64, depth, cfg.pixel_scheme, 3), b.as_drt_film())
    torch.cuda.synchronize()
    a2 = film_mod.FilmPlanes(w, h, n, dev)
    for name in ("sum", "filter", "mean", "m2"):
        getattr(a2, name).copy_(getattr(a, name))
    ctx.film_merge(a.as_drt_film(), b.as_drt_film(), w, h)          # CUDA kernel
    film_mod.merge_pair_(a2, b)                                      # the same arithmetic in torch
    torch.cuda.synchronize()
    assert torch.equal(a.filter, whole.filter)
    for name in ("sum", "mean", "m2"):
        got, ref, tw = getattr(a, name).cpu().numpy(), getattr(whole, name).cpu().numpy(), getattr(a2, name).cpu().numpy()
        assert _rel(got, ref).max() < 2e-4, name
        assert _rel(got, tw).max() < 2e-5, name


def test_4096_square_large_box_tiles(ctx):
    """BASELINE configs[4] geometry: cornell_large_box at 4096x4096 (13.9 GB of film, > 4 GiB offsets), checked on tiles
    against the oracle as the reference itself cannot allocate this frame (u32 sizes, Q22)."""
    import torch
    w = h = 4096
    spp, depth = 2, 4
    free, _ = torch.cuda.mem_get_info()
    if free < 16 * 2 ** 30:
        pytest.skip("needs 16 GB of free device memory")
    cfg, tables, scene, camera = common.load("cornell_large_box", w, h, spp, depth)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F64)
    n = scene.num_wavelengths
    film = film_mod.FilmPlanes(w, h, n, torch.device("cuda", 0))
    prm = oracledriver.params(w, h, 0, spp, depth, cfg.pixel_scheme, 77)
    ctx.render_device(prm, film.as_drt_film())
    torch.cuda.synchronize()
    st = ctx.stats()
    assert st.paths == w * h * spp
    assert float(film.filter.min()) == spp == float(film.filter.max())
    for (x0, y0) in ((0, 0), (2040, 2040), (4080, 4080), (4088, 8)):
        x1, y1 = x0 + 8, y0 + 8
        o_sum, o_avg, o_m2, _, _ = oracledriver.render_tile(scene, camera, prm, x0, y0, x1, y1)
        rows = torch.arange(y0, y1).unsqueeze(1) * w + torch.arange(x0, x1).unsqueeze(0)
        got = film.sum[rows.flatten().cuda()].cpu().numpy()
        assert (_rel(got, o_sum[:, :n]).max(axis=1) <= 1e-3).mean() >= 0.98, (x0, y0)
    del film
    torch.cuda.empty_cache()


def test_merge_many_fused_epilogue(ctx):
    """The fused gather-merge kernel (the multi-GPU epilogue) on one device: three partial films of disjoint sample ranges
    -> merged planes equal the single render, images equal film_to_rgb of the merged film."""
    import torch
    w, h, depth = 40, 24, 4
    cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, 96, depth)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F32)
    n = scene.num_wavelengths
    dev = torch.device("cuda", 0)
    whole = film_mod.FilmPlanes(w, h, n, dev)
    ctx.render_device(oracledriver.params(w, h, 0, 96, depth, cfg.pixel_scheme, 9), whole.as_drt_film())
    parts = [ctx.film_alloc(w, h) for _ in range(3)]
    for f, (s0, s1) in zip(parts, ((0, 32), (32, 50), (50, 96))):
        ctx.render_device(oracledriver.params(w, h, s0, s1, depth, cfg.pixel_scheme, 9), f)
    merged = film_mod.FilmPlanes(w, h, n, dev)
    imgs = [torch.zeros(w * h, dtype=torch.int32, device=dev) for _ in range(3)]
    npix = w * h
    # two calls over disjoint pixel ranges, as two ranks would issue them
    ctx.film_merge_many(merged.as_drt_film(), parts, w, h, 0, npix // 2, bgra=[t.data_ptr() for t in imgs])
    ctx.film_merge_many(merged.as_drt_film(), parts, w, h, npix // 2, npix, bgra=[t.data_ptr() for t in imgs])
    torch.cuda.synchronize()
    assert torch.equal(merged.filter, whole.filter)
    for name in ("sum", "mean", "m2"):
        assert _rel(getattr(merged, name).cpu().numpy(), getattr(whole, name).cpu().numpy()).max() < 2e-4, name
    ref_img = torch.zeros(w * h, dtype=torch.int32, device=dev)
    for which in range(3):
        ctx.film_to_rgb(merged.as_drt_film(), w, h, which, None, ref_img.data_ptr())
        torch.cuda.synchronize()
        a, b = imgs[which].cpu().numpy().view(np.uint32), ref_img.cpu().numpy().view(np.uint32)
        worst = max(np.abs(((a >> s) & 255).astype(int) - ((b >> s) & 255).astype(int)).max() for s in (0, 8, 16))
        assert worst <= 1, which
    for f in parts:
        ctx.film_free(f)


@pytest.mark.parametrize("ranks,w,h", [(2, 40, 24), (3, 37, 23)])
def test_scattered_render_and_slice_merge(ctx, ranks, w, h):
    """The scattered exchange (multi-GPU default) emulated on one device: every "rank" renders its sample range of every pixel
    with drt_cuda_render_device_scatter into the owners' staging films, every owner merges its slice with
    drt_cuda_film_merge_slices -> the merged film equals the single render of all samples.  Ragged: 37*23 pixels over 3 ranks."""
    import torch
    depth, per = 4, 32
    cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, per * ranks, depth)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F32)
    n = scene.num_wavelengths
    dev = torch.device("cuda", 0)
    npix = w * h
    whole = film_mod.FilmPlanes(w, h, n, dev)
    ctx.render_device(oracledriver.params(w, h, 0, per * ranks, depth, cfg.pixel_scheme, 9), whole.as_drt_film())
    slice_px = -(-npix // ranks)
    rows = -(-slice_px * ranks // w)
    staging = [ctx.film_alloc(w, rows) for _ in range(ranks)]          # staging[o]: all ranks' partial films of owner o's slice
    for r in range(ranks):
        ctx.render_device_scatter(oracledriver.params(w, h, r * per, (r + 1) * per, depth, cfg.pixel_scheme, 9), staging, r, slice_px)
    merged = film_mod.FilmPlanes(w, h, n, dev)
    imgs = [torch.zeros(npix, dtype=torch.int32, device=dev) for _ in range(3)]
    for o in range(ranks):
        p0, p1 = min(npix, o * slice_px), min(npix, (o + 1) * slice_px)
        ctx.film_merge_slices(merged.as_drt_film(), staging[o], ranks, slice_px, w, h, p0, p1, bgra=[t.data_ptr() for t in imgs])
    torch.cuda.synchronize()
    assert torch.equal(merged.filter, whole.filter)
    for name in ("sum", "mean", "m2"):
        assert _rel(getattr(merged, name).cpu().numpy(), getattr(whole, name).cpu().numpy()).max() < 2e-4, name
    ref_img = torch.zeros(npix, dtype=torch.int32, device=dev)
    ctx.film_to_rgb(merged.as_drt_film(), w, h, 1, None, ref_img.data_ptr())
    torch.cuda.synchronize()
    a, b = imgs[1].cpu().numpy().view(np.uint32), ref_img.cpu().numpy().view(np.uint32)
    assert max(np.abs(((a >> s) & 255).astype(int) - ((b >> s) & 255).astype(int)).max() for s in (0, 8, 16)) <= 1
    for f in staging:
        ctx.film_free(f)


@pytest.mark.parametrize("ranks,w,h", [(2, 40, 24), (3, 37, 23), (8, 64, 33)])
def test_sharded_exchange_with_device_flags(ctx, ranks, w, h):
    """The round-2 exchange emulated on one device, two epochs back to back WITHOUT any host synchronisation in between: per "rank"
    render + scatter, drt_cuda_flags_signal; per owner drt_cuda_flags_wait, drt_cuda_film_merge_slices_local into the owner's own slice
    film (sharded result), drt_cuda_film_read_slice into one host film.  The assembled film equals the single render of all samples."""
    import torch
    depth, per = 4, 16
    cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, per * ranks, depth)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F32)
    n = scene.num_wavelengths
    dev = torch.device("cuda", 0)
    npix = w * h
    slice_px, parts = film_mod.slice_partition(npix, ranks)
    rows = -(-slice_px * ranks // w)
    staging = [[ctx.film_alloc(w, rows) for _ in range(ranks)] for _ in range(2)]       # [parity][owner]
    slices = [ctx.film_alloc(w, max(1, -(-slice_px // w))) for _ in range(ranks)]
    flags = [ctx.buffer_alloc(64 * 4) for _ in range(ranks)]                             # per owner: arrive[parity][rank]
    imgs = [torch.zeros(npix, dtype=torch.int32, device=dev) for _ in range(3)]
    host = {k: np.zeros((npix, n) if k != "filter" else npix, np.float32) for k in ("sum", "filter", "mean", "m2")}
    host_film = cuda.Film(host["sum"].ctypes.data, host["filter"].ctypes.data, host["mean"].ctypes.data, host["m2"].ctypes.data)
    streams = [torch.cuda.Stream() for _ in range(ranks)]                                # one stream per "rank", as on real peers
    for epoch in (1, 2):
        par, seed = epoch & 1, 9 + epoch
        # all signals are enqueued before any wait: on ONE device streams may share a hardware queue, and a wait queued ahead of the
        # signal it needs would then block it (real peers are separate devices; drt_cuda_render_host_multi enqueues in this order too)
        for r in range(ranks):
            sp = streams[r].cuda_stream
            ctx.render_device_scatter(oracledriver.params(w, h, r * per, (r + 1) * per, depth, cfg.pixel_scheme, seed), staging[par], r, slice_px, stream=sp)
            ctx.flags_signal([f + 4 * (16 * par + r) for f in flags], epoch, stream=sp)
        for r in range(ranks):
            p0, p1 = parts[r]
            sp = streams[r].cuda_stream
            ctx.flags_wait(flags[r] + 4 * 16 * par, ranks, epoch, stream=sp)
            ctx.film_merge_slices_local(slices[r], staging[par][r], ranks, slice_px, w, h, p0, p1, bgra=[t.data_ptr() for t in imgs], stream=sp)
            if epoch == 2:
                ctx.film_read_slice(slices[r], p0, p0, p1, host_film, stream=sp)
    torch.cuda.synchronize()
    assert ctx.flags_timeouts() == 0
    whole = film_mod.FilmPlanes(w, h, n, dev)
    ctx.render_device(oracledriver.params(w, h, 0, per * ranks, depth, cfg.pixel_scheme, 11), whole.as_drt_film())
    torch.cuda.synchronize()
    assert np.array_equal(host["filter"], whole.filter.cpu().numpy())
    for name in ("sum", "mean", "m2"):
        assert _rel(host[name], getattr(whole, name).cpu().numpy()).max() < 2e-4, name
    ref_img = torch.zeros(npix, dtype=torch.int32, device=dev)
    ctx.film_to_rgb(whole.as_drt_film(), w, h, 1, None, ref_img.data_ptr())
    torch.cuda.synchronize()
    a, b = imgs[1].cpu().numpy().view(np.uint32), ref_img.cpu().numpy().view(np.uint32)
    assert max(np.abs(((a >> s) & 255).astype(int) - ((b >> s) & 255).astype(int)).max() for s in (0, 8, 16)) <= 1
    for f in staging[0] + staging[1] + slices:
        ctx.film_free(f)
    for b_ in flags:
        ctx.buffer_free(b_)


@pytest.mark.parametrize("ranks,w,h", [(2, 40, 24), (3, 37, 23)])
def test_banded_scatter_merge_and_read_back(ctx, ranks, w, h):
    """The end-to-end pipeline of the sharded exchange, emulated on one device: the frame is rendered in BANDS (the same part of every
    owner's slice per band, drt_cuda_render_device_scatter_band) on a render stream per "rank"; on a copy stream per rank the band's part
    of the own slice is merged and read back while the next band renders.  The host film equals the single render of all samples, and
    the work counters of the bands add up."""
    import torch
    depth, per = 4, 32
    cfg, tables, scene, camera = common.load("cornell_plane_light", w, h, per * ranks, depth)
    ctx.upload_scene(scene, camera, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F32)
    n = scene.num_wavelengths
    npix = w * h
    slice_px, parts = film_mod.slice_partition(npix, ranks)
    rows = -(-slice_px * ranks // w)
    staging = [ctx.film_alloc(w, rows) for _ in range(ranks)]
    slices = [ctx.film_alloc(w, max(1, -(-slice_px // w))) for _ in range(ranks)]
    flags = [ctx.buffer_alloc(64 * 4) for _ in range(ranks)]
    host = {k: np.zeros((npix, n) if k != "filter" else npix, np.float32) for k in ("sum", "filter", "mean", "m2")}
    host_film = cuda.Film(host["sum"].ctypes.data, host["filter"].ctypes.data, host["mean"].ctypes.data, host["m2"].ctypes.data)
    render = [torch.cuda.Stream() for _ in range(ranks)]
    copy = [torch.cuda.Stream() for _ in range(ranks)]
    cuts = film_mod.ShardedFilmGroup.BAND_CUTS
    paths = 0
    for b in range(len(cuts) - 1):
        b0, b1 = slice_px * cuts[b] // 16, slice_px * cuts[b + 1] // 16
        if b1 <= b0:
            continue
        for r in range(ranks):
            prm = oracledriver.params(w, h, r * per, (r + 1) * per, depth, cfg.pixel_scheme, 21)
            ctx.render_device_scatter_band(prm, staging, r, slice_px, b0, b1, keep_stats=False, stream=render[r].cuda_stream)
            ctx.flags_signal([f + 4 * r for f in flags], b + 1, stream=render[r].cuda_stream)
            torch.cuda.synchronize()            # one context serves all emulated ranks: read each launch's counters before the next
            paths += ctx.stats().paths
        for r in range(ranks):
            p0, p1 = parts[r]
            q0, q1 = min(p1, p0 + b0), min(p1, p0 + b1)
            ctx.flags_wait(flags[r], ranks, b + 1, stream=copy[r].cuda_stream)
            if q1 > q0:
                ctx.film_merge_slices_local(slices[r], staging[r], ranks, slice_px, w, h, q0, q1, stream=copy[r].cuda_stream)
                ctx.film_read_slice(slices[r], p0, q0, q1, host_film, stream=copy[r].cuda_stream)
    torch.cuda.synchronize()
    assert ctx.flags_timeouts() == 0
    assert paths == npix * per * ranks
    whole = film_mod.FilmPlanes(w, h, n, torch.device("cuda", 0))
    ctx.render_device(oracledriver.params(w, h, 0, per * ranks, depth, cfg.pixel_scheme, 21), whole.as_drt_film())
    torch.cuda.synchronize()
    assert np.array_equal(host["filter"], whole.filter.cpu().numpy())
    for name in ("sum", "mean", "m2"):
        assert _rel(host[name], getattr(whole, name).cpu().numpy()).max() < 2e-4, name
    for f in staging + slices:
        ctx.film_free(f)
    for b_ in flags:
        ctx.buffer_free(b_)


@pytest.mark.parametrize("scene", ["cornell_plane_light", "stress_all"])
def test_large_render_is_finite(ctx, scene):
    """67 M paths through glass, gold and mirrors: no sample may poison a pixel with NaN/inf.  (Regression: with the
    approximate f32 sqrt, conductor Fresnel at kappa = 0 produced sqrt(-eps) for single wavelengths.)"""
    import torch
    w = h = 1024
    cfg, tables, sc, cam = common.load(scene, w, h, 64, 4)
    ctx.upload_scene(sc, cam, tables)
    ctx.set_geometry_precision(cuda.GEOMETRY_F32)
    film = film_mod.FilmPlanes(w, h, sc.num_wavelengths, torch.device("cuda", 0))
    ctx.render_device(oracledriver.params(w, h, 0, 64, 4, cfg.pixel_scheme, 5), film.as_drt_film())
    torch.cuda.synchronize()
    for name in ("sum", "mean", "m2"):
        assert bool(torch.isfinite(getattr(film, name)).all()), name
    assert float(film.sum.max()) > 0

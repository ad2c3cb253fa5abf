"""CPU tier: the host-side analyses behind the kernel's two exact cullings (drt_cuda_analyse_scene, plain arithmetic in
libdrt_cuda.so, no device needed) and the slice arithmetic of the scattered multi-GPU exchange.

* hit rectangle: every camera path of a pixel OUTSIDE it must end on the escape material at depth 0 -- checked by tracing those
  pixels with the oracle (which knows nothing of the rectangle) at several samples per pixel and three image shapes;
* boundary planes: the walls of the Cornell rooms are flagged, lights and objects inside are not."""
import importlib

import numpy as np
import pytest

import common
import oracledriver

cuda = importlib.import_module("daily-ray-trace_b200.cuda")
film = importlib.import_module("daily-ray-trace_b200.film")

SCENES = ["init_cornell", "cornell_plane_light", "cornell_large_box", "cornell_downward", "first_scene", "example_scene", "stress_all",
          "rotated_room", "classed_all"]


@pytest.mark.parametrize("w,h", [(64, 48), (40, 56), (33, 33)])
@pytest.mark.parametrize("name", SCENES)
def test_pixels_outside_the_hit_rectangle_see_nothing(name, w, h):
    spp, depth = 6, 4
    cfg, tables, scene, camera = common.load(name, w, h, spp, depth)
    (x0, y0, x1, y1), _ = cuda.analyse_scene(scene, camera, w, h)
    assert 0 <= x0 <= x1 <= w and 0 <= y0 <= y1 <= h
    prm = oracledriver.params(w, h, 0, spp, depth, cfg.pixel_scheme, 31)
    strips = [(0, 0, w, y0), (0, y1, w, h), (0, y0, x0, y1), (x1, y0, w, y1)]      # below, above, left, right of the rectangle
    outside = 0
    for sx0, sy0, sx1, sy1 in strips:
        if sx1 <= sx0 or sy1 <= sy0:
            continue
        total, _, _, _, cnt = oracledriver.render_tile(scene, camera, prm, sx0, sy0, sx1, sy1)
        n = scene.num_wavelengths
        outside += (sx1 - sx0) * (sy1 - sy0)
        assert not total[:, :n].any(), (name, (sx0, sy0, sx1, sy1))
        assert cnt.shaded_bounces == 0 and cnt.shadow_rays == 0 and cnt.closest_rays == cnt.paths == (sx1 - sx0) * (sy1 - sy0) * spp
        assert cnt.terminated_at_depth[0] == cnt.paths
    if name in ("init_cornell", "cornell_plane_light"):
        assert outside >= 0.2 * w * h            # these two boxes fill only part of the frame: the bound must find that


def test_emissive_escape_material_has_no_bound():
    """Q19: an emissive escape material is an environment light (init_scene forces it black-body, daily_ray_trace.c:148; cast_ray
    :452-456 adds throughput * emission when a ray leaves the scene), so pixels that see no surface are LIT: no pixel may be culled.
    sky_cornell = init_cornell with a sky; without the sky the same camera gets a proper rectangle."""
    w, h, spp = 64, 48, 2
    cfg, tables, scene, camera = common.load("sky_cornell", w, h, spp, 4)
    assert camera.aperture_radius == 0 and scene.materials[scene.escape_material].is_emissive
    rect, _ = cuda.analyse_scene(scene, camera, w, h)
    assert rect == (0, 0, w, h)
    # the oracle (bit-exact with the reference) finds light in the frame's corners, which see no surface at all
    prm = oracledriver.params(w, h, 0, spp, 4, cfg.pixel_scheme, 3)
    total, _, _, _, cnt = oracledriver.render_tile(scene, camera, prm, 0, 0, 4, 4)
    n = scene.num_wavelengths
    assert (total[:, :n] > 0).all() and cnt.shaded_bounces == 0
    scene.materials[scene.escape_material].is_emissive = 0
    rect, _ = cuda.analyse_scene(scene, camera, w, h)
    assert rect != (0, 0, w, h) and rect[0] > 4


@pytest.mark.parametrize("w,h", [(64, 48), (33, 33)])
@pytest.mark.parametrize("name", SCENES + ["sky_cornell"])
def test_nothing_lit_outside_the_hit_rectangle(name, w, h):
    """The culling's actual claim, stated on the result: no pixel outside the rectangle receives ANY radiance in the oracle's film
    (zero shaded bounces is not enough when the escape material emits)."""
    spp = 3
    cfg, tables, scene, camera = common.load(name, w, h, spp, 4)
    (x0, y0, x1, y1), _ = cuda.analyse_scene(scene, camera, w, h)
    prm = oracledriver.params(w, h, 0, spp, 4, cfg.pixel_scheme, 77)
    total, _, _, _, _ = oracledriver.render_tile(scene, camera, prm, 0, 0, w, h)
    n = scene.num_wavelengths
    lit = (np.abs(total[:, :n]).max(axis=1) > 0).reshape(h, w)
    inside = np.zeros((h, w), bool)
    inside[y0:y1, x0:x1] = True
    assert not (lit & ~inside).any(), (name, int((lit & ~inside).sum()))


def _copy_scene(scene):
    import ctypes as C
    out = type(scene)()
    C.memmove(C.byref(out), C.byref(scene), C.sizeof(scene))
    return out


def test_validate_scene_rejects_what_the_kernels_would_misindex():
    """drt_cuda_upload_scene is a public entry point; counts and ids that index device tables are range-checked
    (drt_cuda_validate_scene is the same check without a device)."""
    cfg, tables, scene, camera = common.load("cornell_plane_light", 16, 16, 1, 4)
    cuda.validate_scene(scene)

    def bad(mutate, code=-103):
        s = _copy_scene(scene)
        mutate(s)
        with pytest.raises(cuda.CudaError) as e:
            cuda.validate_scene(s)
        assert e.value.code == code, e.value

    def set_lobes(s): s.materials[2].num_lobes = 17
    def set_lobe_id(s): s.materials[2].lobes[0] = 7
    def set_dir(s): s.materials[2].dir_func = 6
    def set_base(s): s.base_material = s.num_materials
    def set_escape(s): s.escape_material = 40
    def set_type(s): s.surfaces[0].type = 9
    def set_mat(s): s.surfaces[1].material = -2
    def set_n(s): s.num_wavelengths = 1
    def set_nsurf(s): s.num_surfaces = 17
    def no_base(s): s.base_material = -1
    def grid(s): s.min_wl = 700.0
    for m in (set_lobes, set_lobe_id, set_dir, set_base, set_escape, set_type, set_mat, set_n, set_nsurf):
        bad(m)
    bad(no_base, -104)
    bad(grid, -104)


def test_thin_lens_has_no_bound():
    cfg, tables, scene, camera = common.load("stress_all", 48, 36, 1, 4)      # the stress scene has aperture > 0
    assert camera.aperture_radius > 0
    rect, _ = cuda.analyse_scene(scene, camera, 48, 36)
    assert rect == (0, 0, 48, 36)


@pytest.mark.parametrize("name", ["init_cornell", "cornell_plane_light", "cornell_large_box", "cornell_downward", "rotated_room"])
def test_room_walls_are_boundary_planes(name):
    cfg, tables, scene, camera = common.load(name, 32, 32, 1, 4)
    _, boundary = cuda.analyse_scene(scene, camera, 32, 32)
    types = [scene.surfaces[i].type for i in range(scene.num_surfaces)]
    planes = [i for i, t in enumerate(types) if t == 3]
    assert sum(boundary) == 5, boundary                      # floor, ceiling, back and side walls
    assert all(types[i] == 3 for i, b in enumerate(boundary) if b)
    # brute force: a flagged plane has every other surface's corners / extents on one side, an unflagged plane does not
    for i in planes:
        f = scene.surfaces[i]
        p, nrm = np.array(f.position[:]), np.array(f.normal[:])
        lo = hi = 0.0
        for j in range(scene.num_surfaces):
            if j == i:
                continue
            o = scene.surfaces[j]
            if o.type == 3:
                pts = [np.array(o.position[:]) + a * np.array(o.u[:]) + b * np.array(o.v[:]) for a in (0, 1) for b in (0, 1)]
                ds = [((q - p) @ nrm, 0.0) for q in pts]
            else:
                ds = [((np.array(o.position[:]) - p) @ nrm, o.radius if o.type == 2 else 0.0)]
            for d, r in ds:
                lo, hi = min(lo, d - r), max(hi, d + r)
        assert boundary[i] == (lo >= -1e-6 or hi <= 1e-6), (name, i)


@pytest.mark.parametrize("npix,world", [(1024 * 1024, 8), (37 * 23, 3), (5, 8), (4096 * 4096, 8), (640 * 480, 6)])
def test_slice_partition_covers_every_pixel_once(npix, world):
    per, parts = film.slice_partition(npix, world)
    assert per * world >= npix and len(parts) == world
    covered = np.zeros(npix, np.int32)
    for p0, p1 in parts:
        assert 0 <= p0 <= p1 <= npix and p1 - p0 <= per
        covered[p0:p1] += 1
    assert (covered == 1).all()


@pytest.mark.parametrize("npix,world", [(1024 * 1024, 8), (37 * 23, 3), (64 * 33, 8), (4096 * 4096, 8), (640 * 480, 6), (100, 2)])
def test_bands_of_the_scattered_render_cover_every_pixel_once(npix, world):
    """The banded exchange renders, per band, the same part [b0, b1) of EVERY owner's slice: task t of a band is pixel
    (t / chunk) * slice + b0 + t % chunk, tasks past the end of the image are empty (render_kernel, RenderLaunch::band_*).  This is
    that arithmetic in numpy with the band boundaries of film.ShardedFilmGroup / drt_cuda_render_host_multi: over all bands every
    pixel is rendered exactly once, and every owner merges exactly its own pixels."""
    slice_px, parts = film.slice_partition(npix, world)
    cuts = film.ShardedFilmGroup.BAND_CUTS
    rendered = np.zeros(npix, np.int32)
    merged = np.zeros(npix, np.int32)
    for b in range(len(cuts) - 1):
        b0, b1 = slice_px * cuts[b] // 16, slice_px * cuts[b + 1] // 16
        if b1 <= b0:
            continue
        chunk = b1 - b0
        t = np.arange(chunk * world, dtype=np.int64)                      # the band's tasks, as the kernel maps them
        pix = (t // chunk) * slice_px + b0 + t % chunk
        pix = pix[pix < npix]
        np.add.at(rendered, pix, 1)
        for r, (p0, p1) in enumerate(parts):                              # what owner r merges and reads back for this band
            q0, q1 = min(p1, p0 + b0), min(p1, p0 + b1)
            merged[q0:q1] += 1
            assert q0 >= p0 and q1 <= p1
    assert (rendered == 1).all() and (merged == 1).all()


@pytest.mark.parametrize("name,mode", [("init_cornell", 1), ("cornell_large_box", 1), ("cornell_downward", 2), ("first_scene", 1), ("example_scene", 1),
                                       ("rotated_room", 1), ("plane_light_all_plastic", 1), ("cornell_plane_light", 2), ("classed_all", 2),
                                       ("stress_all", 0), ("sky_cornell", 0)])
def test_kernel_mode_of_every_scene(name, mode):
    """Which render kernel a scene gets (drt_cuda_plan_scene, the decision drt_cuda_upload_scene takes; host arithmetic): plastic-only
    (1) for plastics under one light, classed (2) when specular / rough-conductor materials join them, general (0) for several lights,
    an emissive escape material or other lobe lists."""
    cfg, tables, scene, camera = common.load(name, 32, 32, 1, 4)
    got, classes, _ = cuda.plan_scene(scene, camera)
    assert got == mode, (name, got, classes)


def test_material_classes_and_specular_constants():
    """classed_all holds every class.  The constants (c0, c1) of a specular material's bdsf() sum c0 + c1 X follow from walking its lobe
    list the way bdsf() does (daily_ray_trace.c:215-229): a lobe that does not write leaves the previous lobe's value in the scratch
    spectrum, which is added AGAIN (Q7) -- so glass [R, T] gives 2 R under reflection and 1 - R under refraction."""
    cfg, tables, scene, camera = common.load("classed_all", 32, 32, 1, 4)
    mode, classes, consts = cuda.plan_scene(scene, camera)
    by_name = {scene.materials[m].name.decode(): m for m in range(scene.num_materials)}
    PLASTIC, SPECULAR, ROUGH = 0, 1, 2
    for nm in ("grey", "red", "sheen", "thick"):
        assert classes[by_name[nm]] == PLASTIC, nm
    for nm in ("chrome", "mirror", "glass", "clear"):
        assert classes[by_name[nm]] == SPECULAR, nm
    assert classes[by_name["rough_gold"]] == ROUGH
    none, refl, refr = 0, 1, 2
    assert consts[by_name["glass"]].tolist() == [[0, 0], [0, 2], [1, -1]]          # R then stale R again; 0 then 1 - R
    assert consts[by_name["clear"]].tolist() == [[0, 0], [0, 0], [1, -1]]          # transmittance only
    assert consts[by_name["mirror"]].tolist() == [[0, 0], [0, 1], [0, 0]]          # mirror_bdsf writes zero on a mismatch
    assert consts[by_name["chrome"]].tolist() == [[0, 0], [0, 1], [0, 0]]
    # stress_all's stale_mix = [bp_diffuse, fs_dielectric_reflectance, bp_diffuse] mixes a plastic lobe with a gated one: general
    cfg, tables, scene, camera = common.load("stress_all", 32, 32, 1, 4)
    mode, classes, _ = cuda.plan_scene(scene, camera)
    by_name = {scene.materials[m].name.decode(): m for m in range(scene.num_materials)}
    assert mode == 0 and classes[by_name["stale_mix"]] == 3

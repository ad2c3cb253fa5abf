"""TEST INFRASTRUCTURE: ctypes wrapper of oracle/_ref/libdrt_ref.so, the UNMODIFIED reference compiled by
oracle/Makefile with per-path RNG streams (oracle/ref_perpath.c).  Used only as a checker / CPU baseline."""
import ctypes as C
import os
import shutil

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(REPO, "oracle", "_ref", "libdrt_ref.so")
REF_BIN = os.path.join(REPO, "oracle", "_ref", "raytrace_ref")


def available():
    return os.path.exists(REF_LIB)


def make_root(root, assets, scene_text=None, scene_name="scene.scn"):
    """A run directory the reference can chdir into: spectra/ and scenes/ (plus an optional generated scene)."""
    os.makedirs(os.path.join(root, "scenes"), exist_ok=True)
    os.makedirs(os.path.join(root, "output"), exist_ok=True)
    if not os.path.exists(os.path.join(root, "spectra")):
        os.symlink(os.path.join(assets, "spectra"), os.path.join(root, "spectra"))
    for f in os.listdir(os.path.join(assets, "scenes")):
        dst = os.path.join(root, "scenes", f)
        if not os.path.exists(dst):
            shutil.copy(os.path.join(assets, "scenes", f), dst)
    if scene_text is not None:
        with open(os.path.join(root, "scenes", scene_name), "w") as fh:
            fh.write(scene_text)
    return root


class Ref:
    """One loaded (config, scene) in the reference library. The library holds global state: one at a time."""

    def __init__(self, root, config_text, seed=0):
        self.lib = L = C.CDLL(REF_LIB)
        L.ref_setup.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_set_seed.argtypes = [C.c_ulonglong]
        L.ref_draw_count.restype = C.c_ulonglong
        dp = C.POINTER(C.c_double)
        L.ref_sample.argtypes = [C.c_uint, C.c_uint, C.c_uint, dp, dp]
        L.ref_render_tile.argtypes = [C.c_uint] * 6 + [dp, dp, dp, dp]
        L.ref_get_camera.argtypes = [dp]
        L.ref_get_surface.argtypes = [C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_int), dp]
        L.ref_get_material.argtypes = [C.c_uint, C.c_char_p, C.POINTER(C.c_int), dp, C.POINTER(C.c_int), dp]
        L.ref_get_tables.argtypes = [dp]
        L.ref_spectrum_to_rgb.argtypes = [dp, dp]
        L.ref_rgb_to_spectrum.argtypes = [dp, dp]
        L.ref_rgb_to_u8.argtypes = [dp]
        L.ref_rgb_to_u8.restype = C.c_uint
        L.ref_blackbody.argtypes = [C.c_double, dp]
        L.ref_bdsf.argtypes = [C.c_uint, C.c_uint, C.c_uint, dp, dp, dp, dp]
        cfg_path = os.path.join(root, "config_ref.cfg")
        with open(cfg_path, "w") as fh:
            fh.write(config_text)
        cwd = os.getcwd()
        try:
            rc = L.ref_setup(root.encode(), b"config_ref.cfg")
        finally:
            os.chdir(cwd)
        if rc != 0:
            raise RuntimeError(f"ref_setup failed: {rc}")
        L.ref_set_seed(seed)
        self.n = L.ref_num_wavelengths()
        self.width, self.height = L.ref_width(), L.ref_height()

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def sample(self, x, y, s):
        out = np.zeros(self.n + 1)
        filt = C.c_double()
        self.lib.ref_sample(x, y, s, self._p(out), C.byref(filt))
        return out[:self.n].copy()

    def render_tile(self, x0, y0, x1, y1, s0, s1, want_paths=False):
        npx, n = (x1 - x0) * (y1 - y0), self.n
        total = np.zeros((npx, n + 1))
        avg = np.zeros((npx, n))
        m2 = np.zeros((npx, n))
        paths = np.zeros((npx, s1 - s0, n)) if want_paths else None
        self.lib.ref_render_tile(x0, y0, x1, y1, s0, s1, self._p(total), self._p(avg), self._p(m2),
                                 self._p(paths) if want_paths else None)
        return total, avg, m2, paths

    def camera(self):
        out = np.zeros(20)
        self.lib.ref_get_camera(self._p(out))
        return out

    def tables(self):
        out = np.zeros((11, self.n))
        self.lib.ref_get_tables(self._p(out))
        return out

    def surface(self, i):
        t, m = C.c_int(), C.c_int()
        geom = np.zeros(13)
        self.lib.ref_get_surface(i, C.byref(t), C.byref(m), self._p(geom))
        return t.value, m.value, geom

    def material(self, i):
        name = C.create_string_buffer(33)
        flags = (C.c_int * 5)()
        scalars = np.zeros(2)
        ids = (C.c_int * 16)()
        spds = np.zeros((6, self.n))
        self.lib.ref_get_material(i, name, flags, self._p(scalars), ids, self._p(spds))
        return dict(name=name.value.decode(), is_black_body=flags[0], is_emissive=flags[1], num_lobes=flags[2],
                    dir_func=flags[3], spd_mask=flags[4], shininess=scalars[0], roughness=scalars[1],
                    lobes=list(ids)[:flags[2]], spds=spds)

    def spectrum_to_rgb(self, spd):
        spd = np.ascontiguousarray(spd, dtype=np.float64)
        out = np.zeros(3)
        self.lib.ref_spectrum_to_rgb(self._p(spd), self._p(out))
        return out

    def rgb_to_spectrum(self, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.float64)
        out = np.zeros(self.n)
        self.lib.ref_rgb_to_spectrum(self._p(rgb), self._p(out))
        return out

    def rgb_to_u8(self, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.float64)
        return self.lib.ref_rgb_to_u8(self._p(rgb))

    def blackbody(self, temp):
        out = np.zeros(self.n)
        self.lib.ref_blackbody(temp, self._p(out))
        return out

    def bdsf(self, surf_mat, inc_mat, trans_mat, normal, out_dir, in_dir):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (normal, out_dir, in_dir)]
        res = np.zeros(self.n)
        self.lib.ref_bdsf(surf_mat, inc_mat, trans_mat, self._p(a[0]), self._p(a[1]), self._p(a[2]), self._p(res))
        return res

"""ctypes binding of libdrt_host.so (include/drt_host.h): the C host front-end of the render path."""
import ctypes as C
import os

from . import PACKAGE_DIR
from ._structs import (Camera, Config, Scene, SceneInput, Tables, PARSE_LEGACY_COMPAT)

_lib = None


class HostError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"drt_host error {code}: {message}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(PACKAGE_DIR, "libdrt_host.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `make -C {PACKAGE_DIR} host` (or __graft_entry__.build())")
        L = C.CDLL(path)
        L.drt_host_last_error.restype = C.c_char_p
        L.drt_parse_config.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Config)]
        L.drt_parse_config_file.argtypes = [C.c_char_p, C.POINTER(Config)]
        L.drt_parse_scene.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(SceneInput)]
        L.drt_scene_apply_compat.argtypes = [C.POINTER(SceneInput)]
        L.drt_scene_write.argtypes = [C.POINTER(SceneInput), C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.drt_load_csv_spectrum.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.drt_load_tables.argtypes = [C.POINTER(Config), C.c_char_p, C.POINTER(Tables)]
        L.drt_rgb_to_spectrum.argtypes = [C.POINTER(Tables), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.drt_rgb_to_spectrum.restype = None
        L.drt_blackbody_spectrum.argtypes = [C.POINTER(Tables), C.c_double, C.POINTER(C.c_double)]
        L.drt_blackbody_spectrum.restype = None
        L.drt_spectrum_to_rgb.argtypes = [C.POINTER(Tables), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.drt_spectrum_to_rgb.restype = None
        L.drt_rgb_to_bgra8.argtypes = [C.POINTER(C.c_double)]
        L.drt_rgb_to_bgra8.restype = C.c_uint32
        L.drt_build_scene.argtypes = [C.POINTER(SceneInput), C.POINTER(Tables), C.c_char_p, C.c_uint32, C.c_uint32,
                                      C.POINTER(Scene), C.POINTER(Camera)]
        L.drt_load_scene_file.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Tables), C.c_int, C.c_uint32, C.c_uint32,
                                          C.POINTER(Scene), C.POINTER(Camera)]
        L.drt_write_spd_sum.argtypes = [C.c_char_p, C.POINTER(Tables), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.drt_write_spd_plain.argtypes = [C.c_char_p, C.POINTER(Tables), C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        L.drt_spd_to_rgb.argtypes = [C.c_char_p, C.POINTER(Tables), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.POINTER(C.c_double))]
        L.drt_write_bmp.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.drt_write_bmp_rgb.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise HostError(rc, lib().drt_host_last_error().decode(errors="replace"))


def parse_config_text(text):
    cfg = Config()
    data = text.encode() if isinstance(text, str) else text
    _check(lib().drt_parse_config(data, len(data), C.byref(cfg)))
    return cfg


def parse_config_file(path):
    cfg = Config()
    _check(lib().drt_parse_config_file(os.fspath(path).encode(), C.byref(cfg)))
    return cfg


def load_tables(cfg, root_dir):
    t = Tables()
    _check(lib().drt_load_tables(C.byref(cfg), os.fspath(root_dir).encode(), C.byref(t)))
    return t


def parse_scene_text(text, flags=PARSE_LEGACY_COMPAT):
    s = SceneInput()
    data = text.encode() if isinstance(text, str) else text
    _check(lib().drt_parse_scene(data, len(data), flags, C.byref(s)))
    return s


def scene_to_text(scene_input):
    buf = C.create_string_buffer(1 << 16)
    n = C.c_size_t()
    _check(lib().drt_scene_write(C.byref(scene_input), buf, len(buf), C.byref(n)))
    return buf.raw[:n.value].decode()


def build_scene(scene_input, tables, root_dir, width, height):
    scene, cam = Scene(), Camera()
    _check(lib().drt_build_scene(C.byref(scene_input), C.byref(tables), os.fspath(root_dir).encode(), width, height,
                                 C.byref(scene), C.byref(cam)))
    return scene, cam


def load_scene_file(root_dir, scene_path, tables, width, height, flags=PARSE_LEGACY_COMPAT):
    scene, cam = Scene(), Camera()
    _check(lib().drt_load_scene_file(os.fspath(root_dir).encode(), os.fspath(scene_path).encode(), C.byref(tables), flags,
                                     width, height, C.byref(scene), C.byref(cam)))
    return scene, cam


DEFAULT_CONFIG_TEXT = """num_pixel_samples {spp}
max_cast_depth    {depth}
output_width      {width}
output_height     {height}
min_wl            {min_wl}
max_wl            {max_wl}
wl_interval       {wl_interval}
pixel_scheme      {scheme}
input_scene       {scene}
output_spd        output\\output.spd
average_spd       output\\average.spd
variance_spd      output\\variance.spd
output_bmp        output\\output.bmp
average_bmp       output\\average.bmp
variance_bmp      output\\variance.bmp
white_spd         spectra\\white_rgb_to_spd.csv
cmf_x             spectra\\cmf_x.csv
cmf_y             spectra\\cmf_y.csv
cmf_z             spectra\\cmf_z.csv
red_spd           spectra\\red_rgb_to_spd.csv
green_spd         spectra\\green_rgb_to_spd.csv
blue_spd          spectra\\blue_rgb_to_spd.csv
cyan_spd          spectra\\cyan_rgb_to_spd.csv
magenta_spd       spectra\\magenta_rgb_to_spd.csv
yellow_spd        spectra\\yellow_rgb_to_spd.csv
"""


def make_config_text(scene="scenes\\cornell_plane_light.scn", width=64, height=64, spp=1, depth=4, scheme="pixel_random",
                     min_wl=380.0, max_wl=720.0, wl_interval=5.0):
    """A config.cfg in the reference's own format (config.cfg:1-25) for the given workload."""
    return DEFAULT_CONFIG_TEXT.format(scene=scene, width=width, height=height, spp=spp, depth=depth, scheme=scheme,
                                      min_wl=f"{min_wl:.1f}", max_wl=f"{max_wl:.1f}", wl_interval=f"{wl_interval:.1f}")

"""Host-side mirror of the reference's render seams, over the C libraries.

    render_image(config)            <- void render_image(config_arguments*)  src/daily_ray_trace.c:635
    sample_scene(renderer, x, y, s) <- void sample_scene(...)                src/daily_ray_trace.c:571
Same argument meaning as the reference (a parsed config.cfg names the scene, image size, sample count, depth,
pixel scheme and output files); errors raise instead of exit(-1)."""
import os

import numpy as np

from . import ASSETS_DIR, cuda, host
from ._structs import PARSE_LEGACY_COMPAT, RenderParams


class Renderer:
    """A config + scene uploaded to one GPU."""

    def __init__(self, config, root_dir=ASSETS_DIR, device=0, seed=0, geometry=cuda.GEOMETRY_F32, scene_flags=PARSE_LEGACY_COMPAT):
        self.config = config
        self.root_dir = root_dir
        self.seed = seed
        self.tables = host.load_tables(config, root_dir)
        self.scene, self.camera = host.load_scene_file(root_dir, config.input_scene.decode(), self.tables,
                                                       config.output_width, config.output_height, scene_flags)
        self.ctx = cuda.Context(device)
        self.ctx.upload_scene(self.scene, self.camera, self.tables)
        self.ctx.set_geometry_precision(geometry)
        self.n = self.scene.num_wavelengths

    def params(self, sample_begin=0, sample_end=None):
        c = self.config
        return RenderParams(c.output_width, c.output_height, sample_begin,
                            c.num_pixel_samples if sample_end is None else sample_end,
                            c.max_cast_depth, c.pixel_scheme, self.seed)

    def render(self, sample_begin=0, sample_end=None):
        """All samples of every pixel; returns host films (the three .spd payloads before widening to f64)."""
        return self.ctx.render_host(self.params(sample_begin, sample_end))

    def sample_scene(self, x, y, sample):
        p = self.params(sample, sample + 1)
        return self.ctx.sample_paths(p, x, y, x + 1, y + 1)[0, 0]

    def close(self):
        self.ctx.close()


def render_image(config, root_dir=ASSETS_DIR, device=0, seed=0, write_files=True):
    """render_image + the three spd_file_to_bmp conversions of win32_main.c:146-152."""
    import ctypes as C
    r = Renderer(config, root_dir, device, seed)
    film = r.render()
    if write_files:
        L = host.lib()
        w, h = config.output_width, config.output_height

        def path(p):
            return os.path.join(root_dir, p.decode().replace("\\", "/")).encode()

        host._check(L.drt_write_spd_sum(path(config.output_spd), C.byref(r.tables), w, h, film["sum"].ctypes.data, film["filter"].ctypes.data))
        host._check(L.drt_write_spd_plain(path(config.average_spd), C.byref(r.tables), w, h, film["mean"].ctypes.data, 0))
        host._check(L.drt_write_spd_plain(path(config.variance_spd), C.byref(r.tables), w, h, film["m2"].ctypes.data, 1))
        for spd, bmp in ((config.output_spd, config.output_bmp), (config.average_spd, config.average_bmp),
                         (config.variance_spd, config.variance_bmp)):
            ww, hh = C.c_uint32(), C.c_uint32()
            rgb = C.POINTER(C.c_double)()
            host._check(L.drt_spd_to_rgb(path(spd), C.byref(r.tables), C.byref(ww), C.byref(hh), C.byref(rgb)))
            host._check(L.drt_write_bmp_rgb(path(bmp), ww.value, hh.value, rgb))
    r.close()
    return film

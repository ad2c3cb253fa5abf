"""CPU tier: the .spd / .bmp writers of the host library against files written by the verbatim reference binary
(oracle/_ref/raytrace_ref = unmodified win32_main.c): same header, same record layout, same conversion to BMP."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import refdriver

pytestmark = pytest.mark.skipif(not os.path.exists(refdriver.REF_BIN), reason="oracle/_ref not built")


def test_spd_and_bmp_bytes_match_reference_files(host, assets, tmp_path):
    root = refdriver.make_root(str(tmp_path), assets)
    w, h, spp = 24, 16, 2
    cfg_text = host.make_config_text(scene="scenes\\cornell_plane_light.scn", width=w, height=h, spp=spp)
    open(os.path.join(root, "config.cfg"), "w").write(cfg_text)
    subprocess.run([refdriver.REF_BIN], cwd=root, check=True, stdout=subprocess.DEVNULL)
    out = os.path.join(root, "output")
    ref_spd = {k: open(os.path.join(out, k + ".spd"), "rb").read() for k in ("output", "average", "variance")}
    ref_bmp = {k: open(os.path.join(out, k + ".bmp"), "rb").read() for k in ("output", "average", "variance")}

    cfg = host.parse_config_text(cfg_text)
    tables = host.load_tables(cfg, assets)
    n = tables.num_wavelengths
    L = host.lib()
    # feed the reference's own film values (narrowed to the f32 the device produces) through this repo's writers
    body = np.frombuffer(ref_spd["output"][40:], dtype=np.float64).reshape(w * h, n + 1)
    total = np.ascontiguousarray(body[:, :n].astype(np.float32))
    filt = np.ascontiguousarray(body[:, n].astype(np.float32))
    mean = np.ascontiguousarray(np.frombuffer(ref_spd["average"][40:], dtype=np.float64).reshape(w * h, n).astype(np.float32))
    mine = os.path.join(str(tmp_path), "mine")
    os.makedirs(mine)
    host._check(L.drt_write_spd_sum(os.path.join(mine, "output.spd").encode(), C.byref(tables), w, h, total.ctypes.data, filt.ctypes.data))
    host._check(L.drt_write_spd_plain(os.path.join(mine, "average.spd").encode(), C.byref(tables), w, h, mean.ctypes.data, 0))
    got = open(os.path.join(mine, "output.spd"), "rb").read()
    assert len(got) == len(ref_spd["output"]) == 40 + w * h * (n + 1) * 8
    assert got[:40] == ref_spd["output"][:40]                        # header bytes incl. id 0xedfeefbe, has_filter, padding
    assert open(os.path.join(mine, "average.spd"), "rb").read()[:40] == ref_spd["average"][:40]
    mine_body = np.frombuffer(got[40:], dtype=np.float64).reshape(w * h, n + 1)
    assert np.allclose(mine_body, body, rtol=1e-6, atol=0)           # f32 narrowing only

    # .spd -> .bmp: convert the REFERENCE's .spd with this repo's converter, compare with the reference's .bmp
    for k in ("output", "average", "variance"):
        ww, hh = C.c_uint32(), C.c_uint32()
        rgb = C.POINTER(C.c_double)()
        host._check(L.drt_spd_to_rgb(os.path.join(out, k + ".spd").encode(), C.byref(tables), C.byref(ww), C.byref(hh), C.byref(rgb)))
        bmp = os.path.join(mine, k + ".bmp")
        host._check(L.drt_write_bmp_rgb(bmp.encode(), ww.value, hh.value, rgb))
        assert open(bmp, "rb").read() == ref_bmp[k], k

"""Shared helpers for the parity tests: load a scene through the product's host front-end."""
import importlib
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(REPO, "assets")
GOLDEN = os.path.join(REPO, "tests", "golden")

host = importlib.import_module("daily-ray-trace_b200.host")
structs = importlib.import_module("daily-ray-trace_b200._structs")


def scene_path(name):
    for d in (os.path.join(ASSETS, "scenes"), os.path.join(GOLDEN, "scenes")):
        p = os.path.join(d, name + ".scn")
        if os.path.exists(p):
            return p
    raise FileNotFoundError(name)


def load(name, width, height, spp=1, depth=4, scheme="pixel_random", **grid):
    """Returns (config, tables, scene, camera) for a shipped or golden scene."""
    cfg = host.parse_config_text(host.make_config_text(width=width, height=height, spp=spp, depth=depth, scheme=scheme, **grid))
    tables = host.load_tables(cfg, ASSETS)
    parsed = host.parse_scene_text(open(scene_path(name)).read())
    scene, camera = host.build_scene(parsed, tables, ASSETS, width, height)
    return cfg, tables, scene, camera


def path_errors(gpu, ref):
    """Per-path max-over-wavelength relative error with the absolute floor of SURVEY.md 8d:
    |gpu - ref| / max(|ref|, 1e-6 * max|ref over the whole set|)."""
    import numpy as np
    floor = 1e-6 * max(float(np.nanmax(np.abs(ref))), 1e-30)
    err = np.abs(gpu.astype(np.float64) - ref) / np.maximum(np.abs(ref), floor)
    both_nan = np.isnan(gpu) & np.isnan(ref)
    err = np.where(both_nan, 0.0, err)
    err = np.where(np.isnan(err), np.inf, err)
    return err.max(axis=-1)

"""daily-ray-trace_b200: B200-native spectral path-tracing core behind the reference's render surface.

The product is C and CUDA (host/ and csrc/, built in-tree by the Makefile); this package is the thin
Python plumbing used by tests/, bench.py and torch.distributed launches:
  host.py   ctypes binding of libdrt_host.so  (config/.scn parsing, spectral setup, .spd/.bmp writers)
  cuda.py   ctypes binding of libdrt_cuda.so  (the C-ABI of include/drt_cuda.h; fails loudly if absent)
  render.py host-side mirror of the reference's render_image / sample_scene seams
  film.py   multi-GPU film merge (sample-sharded ranks) over torch.distributed
"""
import os

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PACKAGE_DIR)
ASSETS_DIR = os.path.join(REPO_ROOT, "assets")

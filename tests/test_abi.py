"""CPU tier: the C-ABI libraries load, export every symbol their headers declare, and refuse to run without a GPU."""
import ctypes as C
import importlib
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(REPO, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drt_[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_header(drt):
    cuda = importlib.import_module("daily-ray-trace_b200.cuda")
    L = cuda.lib()
    names = _declared("drt_cuda.h")
    assert len(names) >= 12
    for name in names:
        assert hasattr(L, name), name
    assert set(cuda.EXPORTS) == set(names)


def test_host_library_exports_header(host):
    L = host.lib()
    for name in _declared("drt_host.h"):
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_device():
    cuda = importlib.import_module("daily-ray-trace_b200.cuda")
    if cuda.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(cuda.CudaError) as e:
        cuda.Context(0)
    assert e.value.code == -101

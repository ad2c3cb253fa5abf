import importlib
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

PKG = "daily-ray-trace_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def drt():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def host():
    return importlib.import_module(PKG + ".host")


@pytest.fixture(scope="session")
def assets():
    return os.path.join(REPO, "assets")

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "julia-ocean-modelling_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False
    n = ctypes.c_int(0)
    return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0


@pytest.fixture(scope="session")
def have_gpu():
    return _have_gpu()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN

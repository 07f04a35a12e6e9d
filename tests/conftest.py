import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcudart.so.12")
        n = ctypes.c_int(0)
        return cuda.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU is a configuration error, not a pass
    if not HAS_GPU:
        skip = pytest.mark.skip(reason="no CUDA device in this container")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_impl():
    """The reference itself when oracle/_ref is built, else the pinned C++ port."""
    from oracle import ref, port
    return ref if ref.available() else port

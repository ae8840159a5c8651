import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _has_gpu():
    try:
        import ctypes
        from qrack_b200 import _abi
        lib = _abi.load()
        n = ctypes.c_int(0)
        return lib.b200sv_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_available():
    return _has_gpu()


def pytest_collection_modifyitems(config, items):
    # a `gpu` test on a box without a device is an error of the invocation, not a silent pass: skip loudly
    if any("gpu" in it.keywords for it in items) and not _has_gpu():
        skip = pytest.mark.skip(reason="no CUDA device visible (gpu-marked test)")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)

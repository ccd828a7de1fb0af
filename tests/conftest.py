import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # many tests put several sub-domains (two streams each, spinning flag waits) on ONE GPU: more hardware
    # queues than CUDA's default of 8, set before the first CUDA call of the session (capi.want_hardware_queues)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def _have_gpu():
    try:
        import ctypes
        from turbulent_lbm_multigpu_b200 import capi
        n = ctypes.c_int()
        return capi.load().lbmGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

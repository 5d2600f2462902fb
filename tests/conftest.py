import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    """Device detection through the product library itself (ctypes -> sarasmg_device_count): the CUDA path has no torch
    dependency, so a box with a CPU-only or missing torch must not turn the parity suite into silent skips."""
    try:
        import ctypes as C
        from saras_b200 import capi
        n = C.c_int(0)
        return capi.load().sarasmg_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    if os.environ.get("SARASMG_REQUIRE_GPU") and any("gpu" in item.keywords for item in items):
        raise pytest.UsageError("SARASMG_REQUIRE_GPU is set but libsarasmg.so sees no CUDA device: refusing to skip the GPU parity tests")
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

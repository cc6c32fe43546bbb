import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    """Devices visible to the CUDA runtime, without importing torch (ctypes on the library's own entry point)."""
    try:
        from sfb_b200 import _lib
        import ctypes
        n = ctypes.c_int32(0)
        return int(n.value) if _lib.load().sfb_device_count(ctypes.byref(n)) == 0 else 0
    except Exception:  # noqa: BLE001  (library not built, no driver, ...)
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a machine without a CUDA device."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (the product has no CPU fallback)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def sfb():
    import sfb_b200
    return sfb_b200


def relerr(a, b):
    """Norm-wise relative error, what Julia's `≈` with rtol means for arrays."""
    import numpy as np
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))

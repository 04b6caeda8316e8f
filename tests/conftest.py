import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (the .so files are git-ignored): compile it once, exactly as
    __graft_entry__.build() does, so that the suite does not depend on the order the driver runs its steps in."""
    from knaster_b200 import _ffi

    if not os.path.exists(_ffi.LIB_PATH):
        _ffi.build()


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device: on a box without one they are skipped, not failed (the engine has no CPU
    fallback, so every call would raise KGPU_ERR_CUDA and bury real CPU-side regressions)."""
    import pytest

    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    from knaster_b200 import _ffi

    try:
        n = int(_ffi.lib().kgpu_device_count())
    except Exception:
        n = 0
    if n == 0:
        skip = pytest.mark.skip(reason="no CUDA device on this box (kgpu_device_count() == 0)")
        for it in gpu_items:
            it.add_marker(skip)

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

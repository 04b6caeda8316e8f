"""CPU: the per-template kernel generator (csrc/jit.cpp).  NVRTC needs no GPU, so "every supported voice shape generates
and compiles for sm_100a" is checked here; what the kernels compute is checked on the GPU (tests/test_gpu_jit.py)."""
import os

import numpy as np
import pytest

import knaster_b200 as kn
from knaster_b200 import _ffi, banks
from knaster_b200.graph import Graph


def nvrtc_available():
    import ctypes

    for name in ("libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"):
        try:
            ctypes.CDLL(name)
            return True
        except OSError:
            pass
    return False


pytestmark = pytest.mark.skipif(not nvrtc_available(), reason="libnvrtc is not installed")


def test_chain_bank_generates_compiles_and_is_cached():
    g = Graph(0, 2, 64, 48000)
    banks.chain_bank(g, 8, 0.5)
    n, cached = _ffi.jit_compile(g)
    assert n == 1                                 # eight isomorphic voices = one template = one kernel
    n2, cached2 = _ffi.jit_compile(g)
    assert (n2, cached2) == (1, 1)                # second time: from the cubin cache next to the library
    cache = os.path.join(os.path.dirname(_ffi.LIB_PATH), "jit")
    assert any(f.endswith(".cubin") for f in os.listdir(cache))


def test_hand_written_recipes_are_not_regenerated():
    g = Graph(0, 2, 64, 48000)
    banks.subtractive_bank(g, 8, 0.5)
    assert _ffi.jit_compile(g) == (0, 0)


@pytest.mark.parametrize("seed", [11, 17])
def test_fuzz_voice_shapes_generate_and_compile(seed):
    # every source / filter / envelope / wrapper / audio-rate route the fuzz generator of test_gpu_fuzz.py draws from
    import test_gpu_fuzz as fz

    g = Graph(0, 2, 64, 48000)
    kn.reset_randomness_seed(0)
    r = np.random.Generator(np.random.PCG64(seed))
    with g.edit() as ge:
        for vi in range(20):
            fz.random_voice(ge, r, vi).out([0, 0]).to_graph_out()
    n, _ = _ffi.jit_compile(g, tap_outputs=True)
    assert n >= 19

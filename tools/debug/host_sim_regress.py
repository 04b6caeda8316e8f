"""Dumps a digest of the host control simulation's device-event streams (kgpu_debug_simulate) for a set of
graphs and launch splits.  Run it with KNASTER_GPU_LIB pointing at two builds and diff the output: the
event streams must be identical event for event (used while the simulation core was rewritten in round 2)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import knaster_b200 as kn
from knaster_b200 import _ffi, banks
from knaster_b200.graph import Graph


def digest(graph, n_blocks, bpc):
    ev = graph.take_events()
    evs, nodes, info = _ffi.debug_simulate(graph, ev, n_blocks, bpc, cap=1 << 21)
    h = hashlib.sha256(repr(evs).encode()).hexdigest()[:16]
    return h, len(evs), info["dropped_changes"], info["ignored_delays"], info["device_events"]


def cases():
    for wl, nv, secs in (("subtractive", 300, 1.0), ("subtractive_seg", 200, 1.0), ("additive", 200, 1.0), ("fm", 50, 0.5)):
        for bpc in (0, 1, 7, 100):
            if bpc == 1 and nv > 100:
                nvv = 40
            else:
                nvv = nv
            g = Graph(0, 2, 64, 48000)
            banks.bank_builder(wl, secs)(g, nvv, 0, nvv)
            yield f"{wl}/{nvv}/{bpc}", g, int(secs * 48000) // 64, bpc
    # the fuzz generator's voice shapes (wrappers of every kind, smoothing + precise timing nests, AR routes)
    os.environ.setdefault("KGPU_NO_GPU_IMPORT", "1")
    import importlib

    fz = importlib.import_module("test_gpu_fuzz")
    for seed in range(11, 27):
        for bpc in (0, 3, 16):
            g = Graph(0, 2, 64, 48000)
            kn.reset_randomness_seed(0)
            r = np.random.Generator(np.random.PCG64(seed))
            with g.edit() as ge:
                for vi in range(20):
                    sig = fz.random_voice(ge, r, vi)
                    sig.out([0, 0]).to_graph_out()
            yield f"fuzz{seed}/{bpc}", g, fz.N_BLOCKS, bpc


for name, g, nb, bpc in cases():
    print(name, *digest(g, nb, bpc))

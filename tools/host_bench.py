"""Times the host half of a render call (validation, bucketing, control-rate simulation of the parameter-change
queue -> device events) without a GPU, driven launch by launch the way kgpu_render drives it.
usage: host_bench.py [workload] [voices] [seconds] [steps] [threads]   (threads 3 = one GPU's share of a 32-core box with 8 GPUs)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from knaster_b200 import _ffi, banks
from knaster_b200.graph import Graph

workload = sys.argv[1] if len(sys.argv) > 1 else "subtractive"
voices = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
threads = int(sys.argv[5]) if len(sys.argv) > 5 else 3
g = Graph(0, 2, 64, 48000)
banks.bank_builder(workload, seconds)(g, voices, 0, voices)
ev = np.ascontiguousarray(g.take_events())
n_blocks = int(seconds * 48000) // 64
L = _ffi.lib()
L.kgpu_debug_host_bench.argtypes = [C.POINTER(_ffi.GraphDesc), C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
gd, keep = _ffi.graph_desc(g)
out = np.zeros((steps, 4))
_ffi.check(L.kgpu_debug_host_bench(C.byref(gd), ev.ctypes.data, len(ev), n_blocks, int(os.environ.get("BPL", "2048")), steps, threads, out.ctypes.data))
print(f"{workload} {voices} voices x {seconds:g} s, {len(ev)} events/step, {threads} worker threads; ms per step:")
print("   push   begin  first-launch-ready  all-launches")
for r in out:
    print("  %6.2f %6.2f %12.2f %14.2f" % tuple(r))

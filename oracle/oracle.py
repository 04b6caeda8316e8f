"""ctypes binding of the CPU oracle (oracle/knaster_oracle.cpp).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product
package (knaster_b200) never does.

``OracleProcessor`` renders a ``knaster_b200.graph.Graph`` with the oracle and
mirrors ``AudioProcessor``'s surface (processor.rs:119-197) so that a parity test
can run the same graph + events through both and compare.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build(force: bool = False) -> None:
    """Compile the oracle (both the parity and the -O3 baseline build)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


class _WrapperDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("capacity", C.c_uint32), ("value", C.c_double)]


class _NodeDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("mode", C.c_uint32), ("channels", C.c_uint32), ("flags", C.c_uint32),
                ("args", C.c_double * 4), ("n_wrappers", C.c_uint32), ("n_segments", C.c_uint32),
                ("wrappers", C.POINTER(_WrapperDesc)), ("segments", C.POINTER(C.c_double))]


def _cpu_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " avx2 " in line + " "
    except OSError:
        pass
    return False


_libs = {}


def load(fast: bool = False) -> C.CDLL:
    name = "libknaster_oracle_fast.so" if (fast and _cpu_has_avx2()) else "libknaster_oracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_BUILD, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    lib.ko_graph_create.restype = vp
    lib.ko_graph_create.argtypes = [u32, u32, u32, u32, u32]
    lib.ko_graph_destroy.argtypes = [vp]
    lib.ko_last_error.restype = C.c_char_p
    lib.ko_push.argtypes = [vp, C.POINTER(_NodeDesc)]
    lib.ko_node_inputs.argtypes = [vp, i32]
    lib.ko_node_outputs.argtypes = [vp, i32]
    lib.ko_node_parameters.argtypes = [vp, i32]
    lib.ko_set_input_edge.argtypes = [vp, i32, u32, i32, u32]
    lib.ko_set_output_edge.argtypes = [vp, u32, i32, u32]
    lib.ko_set_param_edge.argtypes = [vp, i32, u32, i32, u32]
    lib.ko_commit.argtypes = [vp]
    lib.ko_send_event.argtypes = [vp, vp]
    lib.ko_run_block.argtypes = [vp, vp]
    lib.ko_output_block.restype = C.POINTER(C.c_float)
    lib.ko_output_block.argtypes = [vp]
    lib.ko_frame_clock.restype = u64
    lib.ko_frame_clock.argtypes = [vp]
    lib.ko_log_count.restype = u64
    lib.ko_log_count.argtypes = [vp]
    lib.ko_add_tap.argtypes = [vp, i32, u32]
    lib.ko_render.argtypes = [vp, u64, vp, vp, C.c_size_t, vp]
    lib.ko_seconds_from_samples.argtypes = [u64, u64, C.POINTER(u32), C.POINTER(u32)]
    lib.ko_seconds_to_samples.restype = u64
    lib.ko_seconds_to_samples.argtypes = [u32, u32, u64]
    lib.ko_seconds_from_secs_f64.argtypes = [C.c_double, C.POINTER(u32), C.POINTER(u32)]
    lib.ko_seconds_to_tesimals.restype = u64
    lib.ko_seconds_to_tesimals.argtypes = [u32, u32]
    lib.ko_seconds_from_tesimals.argtypes = [u64, C.POINTER(u32), C.POINTER(u32)]
    lib.ko_seconds_add.argtypes = [u32, u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
    lib.ko_ugen_create.restype = vp
    lib.ko_ugen_create.argtypes = [C.POINTER(_NodeDesc), u32, u32]
    lib.ko_ugen_destroy.argtypes = [vp]
    lib.ko_ugen_set_delay.argtypes = [vp, u32, u32]
    lib.ko_ugen_param.argtypes = [vp, u32, vp]
    lib.ko_ugen_process_block.argtypes = [vp, vp, vp, u32]
    lib.ko_ugen_process.argtypes = [vp, vp, vp]
    _libs[name] = lib
    return lib


class OracleError(RuntimeError):
    pass


def _node_desc(ug) -> Tuple[_NodeDesc, list]:
    """knaster_b200.ugens.UGen -> ko_node_desc (+ keep-alive list)."""
    keep = []
    d = _NodeDesc()
    d.kind, d.mode, d.channels, d.flags = ug.kind, ug.mode, ug.channels, ug.flags
    for i, a in enumerate(ug.args[:4]):
        d.args[i] = a
    d.n_wrappers = len(ug.wrappers)
    if ug.wrappers:
        arr = (_WrapperDesc * len(ug.wrappers))()
        for i, w in enumerate(ug.wrappers):
            arr[i].kind, arr[i].capacity, arr[i].value = w.kind, w.capacity, w.value
        keep.append(arr)
        d.wrappers = C.cast(arr, C.POINTER(_WrapperDesc))
    d.n_segments = len(ug.segments)
    if ug.segments:
        flat = (C.c_double * (2 * len(ug.segments)))()
        for i, (du, va) in enumerate(ug.segments):
            flat[2 * i], flat[2 * i + 1] = du, va
        keep.append(flat)
        d.segments = C.cast(flat, C.POINTER(C.c_double))
    return d, keep


class OracleUGen:
    """A bare (wrapped) UGen outside any graph, as the reference's wrappers_core.rs tests use."""

    def __init__(self, ugen, sample_rate: int = 48000, block_size: int = 16):
        self.lib = load()
        d, keep = _node_desc(ugen)
        self.h = self.lib.ko_ugen_create(C.byref(d), sample_rate, block_size)
        if not self.h:
            raise OracleError(self.lib.ko_last_error().decode())
        self.n_in, self.n_out = ugen.inputs(), ugen.outputs()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ko_ugen_destroy(self.h)
            self.h = None

    def set_delay_within_block_for_param(self, param: int, delay: int) -> None:
        self.lib.ko_ugen_set_delay(self.h, param, delay)

    def param(self, param: int, value) -> None:
        from knaster_b200.graph import EVENT_DTYPE, _value_kind

        ev = np.zeros(1, dtype=EVENT_DTYPE)
        k, v = _value_kind(value)
        ev["value_kind"], ev["value"] = k, v
        if self.lib.ko_ugen_param(self.h, param, ev.ctypes.data):
            raise OracleError(self.lib.ko_last_error().decode())

    def process(self, inputs: Sequence[float] = ()) -> np.ndarray:
        fin = np.asarray(list(inputs) + [0.0], dtype=np.float32)
        out = np.zeros(max(self.n_out, 1), dtype=np.float32)
        self.lib.ko_ugen_process(self.h, fin.ctypes.data, out.ctypes.data)
        return out[: self.n_out]

    def process_block(self, inputs: np.ndarray, frames: int) -> np.ndarray:
        fin = np.ascontiguousarray(inputs, dtype=np.float32).reshape(max(self.n_in, 1), frames)
        out = np.zeros((max(self.n_out, 1), frames), dtype=np.float32)
        self.lib.ko_ugen_process_block(self.h, fin.ctypes.data, out.ctypes.data, frames)
        return out[: self.n_out]


class OracleProcessor:
    """Oracle counterpart of AudioProcessor for a knaster_b200.graph.Graph."""

    def __init__(self, graph, ring_buffer_size: int = 1000, fast: bool = False):
        self.lib = load(fast)
        self.graph = graph
        self.block_size = graph.block_size
        self.n_out = graph.num_outputs
        self.n_in = graph.num_inputs
        self.ring_buffer_size = ring_buffer_size
        self.h = None
        self._built_version = -1
        self._n_nodes_built = 0
        self._taps: List[Tuple[int, int]] = []

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ko_graph_destroy(self.h)
            self.h = None

    def _err(self) -> OracleError:
        return OracleError(self.lib.ko_last_error().decode())

    def _sync(self) -> None:
        g = self.graph
        g.commit_changes()
        if self.h is None:
            self.h = self.lib.ko_graph_create(g.sample_rate, g.block_size, g.num_inputs, g.num_outputs,
                                              self.ring_buffer_size)
            if not self.h:
                raise self._err()
        if self._built_version == g.version:
            return
        # push new nodes (state of existing nodes is kept, like TaskData hand-over task.rs:107-111)
        for i in range(self._n_nodes_built, len(g.nodes)):
            d, _keep = _node_desc(g.nodes[i].ugen)
            if self.lib.ko_push(self.h, C.byref(d)) < 0:
                raise self._err()
        self._n_nodes_built = len(g.nodes)
        for sink, edges in enumerate(g.node_input_edges):
            for ch, e in enumerate(edges):
                src, sc = e if e is not None else (-1, 0)
                if self.lib.ko_set_input_edge(self.h, sink, ch, src, sc):
                    raise self._err()
        for sink, edges in enumerate(g.node_parameter_edges):
            for (p, s, c) in edges:
                if self.lib.ko_set_param_edge(self.h, sink, p, s, c):
                    raise self._err()
        for ch, e in enumerate(g.output_edges):
            src, sc = e if e is not None else (-1, 0)
            if self.lib.ko_set_output_edge(self.h, ch, src, sc):
                raise self._err()
        if self.lib.ko_commit(self.h):
            raise self._err()
        self._built_version = g.version

    # -- AudioProcessor surface
    def run_without_inputs(self) -> None:
        assert self.n_in == 0
        self.run([])

    def run(self, inputs: Sequence[np.ndarray]) -> None:
        self._sync()
        ev = self.graph.take_events()
        for i in range(len(ev)):
            rc = self.lib.ko_send_event(self.h, ev[i : i + 1].ctypes.data)
            if rc == -1:
                raise self._err()
        ptrs = None
        if self.n_in:
            bufs = [np.ascontiguousarray(x, dtype=np.float32) for x in inputs]
            assert len(bufs) == self.n_in
            arr = (C.c_void_p * self.n_in)(*[b.ctypes.data for b in bufs])
            ptrs = C.cast(arr, C.c_void_p)
        if self.lib.ko_run_block(self.h, ptrs):
            raise self._err()

    def output_block(self) -> np.ndarray:
        p = self.lib.ko_output_block(self.h)
        return np.ctypeslib.as_array(p, shape=(self.n_out, self.block_size)).copy()

    def frame_clock(self) -> int:
        return int(self.lib.ko_frame_clock(self.h)) if self.h else 0

    def log_count(self) -> int:
        return int(self.lib.ko_log_count(self.h)) if self.h else 0

    # -- batched render (just-in-time event feed), optional per-node taps
    def add_tap(self, node: int, channel: int = 0) -> int:
        self._sync()
        t = self.lib.ko_add_tap(self.h, int(node), channel)
        if t < 0:
            raise self._err()
        self._taps.append((int(node), channel))
        return t

    def render(self, n_blocks: int, want_output: bool = True):
        """Returns (out [n_blocks, n_out, block], taps [n_taps, n_blocks*block])."""
        self._sync()
        ev = self.graph.take_events()
        out = np.zeros((n_blocks, self.n_out, self.block_size), dtype=np.float32) if want_output else None
        taps = np.zeros((len(self._taps), n_blocks * self.block_size), dtype=np.float32) if self._taps else None
        rc = self.lib.ko_render(self.h, n_blocks, out.ctypes.data if out is not None else None,
                                ev.ctypes.data if len(ev) else None, len(ev),
                                taps.ctypes.data if taps is not None else None)
        if rc:
            raise self._err()
        return out, taps

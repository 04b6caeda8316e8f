"""Host-side mirror of knaster_graph's control side: Graph / GraphEdit / handles /
Parameter / Time / SchedulingEvent.

This is the *builder* half of the drop-in boundary.  It keeps the
reference's names and semantics:

* ``Graph.edit``            graph.rs:1410 (commit on leaving the edit, graph_edit.rs:258-262)
* ``GraphEdit.push``        graph_edit.rs:88
* ``SH.out / to / >> / | / to_graph_out / to_graph_out_channels / link / param``
                            graph_edit.rs:280-463,735-796,1145-1259
* operators ``* + - /``     graph_edit.rs:936-1225: push ``Constant`` + ``MathUGen`` nodes,
                            constant-on-the-left keeps the UGen as operand 0 (:1183-1192)
* additive connections      graph.rs:768-881: a second source on the same input/output
                            inserts ``MathUGen<Add>`` => left-fold chain in call order
* ``Parameter.set* / smooth* / trig*``   graph_edit.rs:1700-1886
* ``Graph.set / set_many``  graph.rs:1348-1404

The graph held here is what knaster's ``Graph`` holds at ``commit_changes`` time
(nodes, per-sink input edges, parameter edges, output edges).  A backend (the
CUDA engine in product code, the CPU oracle in tests) receives it lowered.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import ugens as U

GRAPH = -2  # NodeOrGraph::Graph (graph input as a source / graph output as a sink)
NONE = -1

SUBSECOND_TESIMALS_PER_SECOND = 282_240_000  # knaster_primitives/src/time.rs:10


class GraphError(Exception):
    """knaster_graph/src/graph.rs:2222-2274"""


class ParameterError(GraphError):
    """knaster_core/src/parameters.rs:65-85"""


@dataclass(frozen=True)
class Seconds:
    """knaster_primitives/src/time.rs:25-90: fixed-point seconds."""

    seconds: int = 0
    subsecond_tesimals: int = 0

    @staticmethod
    def from_samples(samples: int, sample_rate: int) -> "Seconds":  # time.rs:76-84
        return Seconds(samples // sample_rate,
                       (samples % sample_rate) * SUBSECOND_TESIMALS_PER_SECOND // sample_rate)

    def to_samples(self, sample_rate: int) -> int:  # time.rs:86-90
        return self.seconds * sample_rate + (self.subsecond_tesimals * sample_rate) // SUBSECOND_TESIMALS_PER_SECOND

    @staticmethod
    def from_secs_f64(s: float) -> "Seconds":  # time.rs:58-63
        import math

        sec = int(math.floor(s))
        fract = s - math.trunc(s)
        return Seconds(sec, int(fract * SUBSECOND_TESIMALS_PER_SECOND))


@dataclass(frozen=True)
class Time:
    """knaster_graph/src/scheduling.rs:73-139"""

    seconds: Seconds
    absolute: bool

    @staticmethod
    def at(secs: Seconds) -> "Time":
        return Time(secs, True)

    @staticmethod
    def after(secs: Seconds) -> "Time":
        return Time(secs, False)

    @staticmethod
    def asap() -> "Time":
        return Time(Seconds(0, 0), False)


@dataclass(frozen=True)
class ParameterSmoothing:
    """knaster_core/src/parameters/types.rs:108-114.  kind: 'none' | 'linear'."""

    kind: str = "none"
    seconds: float = 0.0

    @staticmethod
    def NoSmoothing() -> "ParameterSmoothing":
        return ParameterSmoothing("none", 0.0)

    @staticmethod
    def Linear(seconds: float) -> "ParameterSmoothing":
        return ParameterSmoothing("linear", float(seconds))


class PTrigger:
    """knaster_core parameters: marker for ParameterValue::Trigger."""


@dataclass
class SchedulingEvent:
    """knaster_graph/src/scheduling.rs:29-36 (token unsupported: activate() is todo!())."""

    node: int
    parameter: int
    value_kind: int = 0  # 0 none, 1 float, 2 trigger, 3 integer, 4 bool
    value: float = 0.0
    smoothing: Optional[ParameterSmoothing] = None
    time: Optional[Time] = None


# Wire format of one scheduling event; the same layout as ``kgpu_event`` in
# include/knaster_gpu.h (the oracle declares an identical struct of its own).
EVENT_DTYPE = np.dtype([
    ("node", "<u4"), ("param", "<u4"),
    ("value_kind", "<u4"),       # 0 none, 1 float, 2 trigger, 3 integer, 4 bool
    ("smoothing_kind", "<u4"),   # 0 no smoothing field, 1 ParameterSmoothing::None, 2 Linear
    ("value", "<f8"),
    ("smooth_seconds", "<f4"), ("smooth_rate", "<u4"),
    ("time_kind", "<u4"),        # 0 None, 1 Time::at, 2 Time::after
    ("seconds", "<u4"), ("subsec", "<u4"), ("_pad", "<u4"),
])
assert EVENT_DTYPE.itemsize == 48


def events_to_array(events: Sequence["SchedulingEvent"]) -> np.ndarray:
    arr = np.zeros(len(events), dtype=EVENT_DTYPE)
    for i, e in enumerate(events):
        r = arr[i]
        r["node"] = e.node
        r["param"] = e.parameter
        r["value_kind"] = e.value_kind
        r["value"] = e.value
        if e.smoothing is not None:
            r["smoothing_kind"] = 2 if e.smoothing.kind == "linear" else 1
            r["smooth_seconds"] = e.smoothing.seconds
        if e.time is not None:
            r["time_kind"] = 1 if e.time.absolute else 2
            r["seconds"] = e.time.seconds.seconds
            r["subsec"] = e.time.seconds.subsecond_tesimals
    return arr


def _value_kind(value) -> Tuple[int, float]:
    """ParameterValue::from (types.rs:39-64)."""
    if value is PTrigger or isinstance(value, PTrigger):
        return 2, 0.0
    if isinstance(value, bool):
        return 4, float(value)
    if isinstance(value, int):
        return 3, float(value)
    return 1, float(value)


@dataclass
class _Node:
    ugen: U.UGen
    auto_math_node: bool = False  # graph.rs:874-881
    name: str = ""


class Graph:
    """knaster_graph/src/graph.rs:159-199 (control side only; static graphs)."""

    def __init__(self, inputs: int, outputs: int, block_size: int, sample_rate: int):
        self.num_inputs = inputs
        self.num_outputs = outputs
        self.block_size = block_size
        self.sample_rate = sample_rate
        self.nodes: List[_Node] = []
        self.node_input_edges: List[List[Optional[Tuple[int, int]]]] = []  # per sink: (source, channel)
        self.node_parameter_edges: List[List[Tuple[int, int, int]]] = []  # (param, source, channel)
        self.output_edges: List[Optional[Tuple[int, int]]] = [None] * outputs
        self.pending_events: List[SchedulingEvent] = []
        self.pending_event_arrays: List[np.ndarray] = []  # bulk path (EVENT_DTYPE), see schedule_bulk
        self.recalculation_required = False
        self.version = 0  # bumped on every structural commit
        self._commit_listeners: List[Callable[["Graph"], None]] = []

    # -- graph.rs:373-389, 462-475
    def push_internal(self, ugen: U.UGen, auto_math_node: bool = False) -> int:
        if not isinstance(ugen, U.UGen):
            raise GraphError("push expects a UGen description")
        self.nodes.append(_Node(ugen, auto_math_node))
        self.node_input_edges.append([None] * ugen.inputs())
        self.node_parameter_edges.append([])
        self.recalculation_required = True
        return len(self.nodes) - 1

    def _check_source(self, source: int, channel: int) -> None:
        if source == GRAPH:
            if channel >= self.num_inputs:
                raise GraphError(f"GraphInputOutOfBounds({channel})")
            return
        if source < 0 or source >= len(self.nodes):
            raise GraphError("NodeNotFound")
        if channel >= self.nodes[source].ugen.outputs():
            raise GraphError(f"OutputOutOfBounds({channel})")

    def new_additive_node(self) -> int:  # graph.rs:874-881
        return self.push_internal(U.MathUGen(1, U.MathOp.Add), auto_math_node=True)

    # -- graph.rs:768-826 (connect_to_node_internal) / 827-872 (connect_to_output_internal)
    def connect2(self, source: int, source_channel: int, sink_channel: int, sink: int,
                 additive: bool = True) -> None:
        self._check_source(source, source_channel)
        self.recalculation_required = True
        if sink == GRAPH:
            if sink_channel >= self.num_outputs:
                raise GraphError(f"GraphOutputOutOfBounds({sink_channel})")
            existing = self.output_edges[sink_channel]
            if additive and existing is not None:
                add = self.new_additive_node()
                self.node_input_edges[add][0] = existing
                self.node_input_edges[add][1] = (source, source_channel)
                self.output_edges[sink_channel] = (add, 0)
            else:
                self.output_edges[sink_channel] = (source, source_channel)
            return
        if sink < 0 or sink >= len(self.nodes):
            raise GraphError("NodeNotFound")
        if sink_channel >= self.nodes[sink].ugen.inputs():
            raise GraphError(f"InputOutOfBounds({sink_channel})")
        existing = self.node_input_edges[sink][sink_channel]
        if additive and existing is not None:
            add = self.new_additive_node()
            self.node_input_edges[add][0] = existing
            self.node_input_edges[add][1] = (source, source_channel)
            self.node_input_edges[sink][sink_channel] = (add, 0)
        else:
            self.node_input_edges[sink][sink_channel] = (source, source_channel)

    def connect2_replace(self, source: int, source_channel: int, sink_channel: int, sink: int) -> None:
        self.connect2(source, source_channel, sink_channel, sink, additive=False)

    # -- graph.rs:997-1053
    def disconnect_output_from_source(self, source: int, source_channel: int) -> None:
        self._check_source(source, source_channel)
        for edges in self.node_input_edges:
            for ch, e in enumerate(edges):
                if e is not None and e == (source, source_channel):
                    edges[ch] = None
        self.recalculation_required = True

    # -- graph.rs:1055-1100
    def disconnect_input_to_sink(self, sink_channel: int, sink: int) -> None:
        if sink == GRAPH:
            if sink_channel >= self.num_outputs:
                raise GraphError(f"GraphOutputOutOfBounds({sink_channel})")
            self.output_edges[sink_channel] = None
        else:
            if sink < 0 or sink >= len(self.nodes):
                raise GraphError("NodeNotFound")
            if sink_channel >= len(self.node_input_edges[sink]):
                raise GraphError(f"InputOutOfBounds({sink_channel})")
            self.node_input_edges[sink][sink_channel] = None
        self.recalculation_required = True

    def param_index(self, node: int, param: Union[int, str]) -> int:
        if node < 0 or node >= len(self.nodes):
            raise GraphError("NodeNotFound")
        descs = self.nodes[node].ugen.param_descriptions()
        if isinstance(param, str):
            if param not in descs:
                raise ParameterError(f"DescriptionNotFound({param})")
            return descs.index(param)
        if param >= len(descs):
            raise ParameterError("ParameterIndexOutOfBounds")
        return int(param)

    # -- graph.rs:620-760 (connect_node_to_parameter)
    def connect_replace_to_parameter(self, source: int, source_channel: int, parameter: Union[int, str],
                                     sink: int) -> None:
        self._connect_node_to_parameter(source, source_channel, parameter, sink, additive=False)

    def connect_to_parameter(self, source: int, source_channel: int, parameter: Union[int, str],
                             sink: int) -> None:
        self._connect_node_to_parameter(source, source_channel, parameter, sink, additive=True)

    def _connect_node_to_parameter(self, source, source_channel, parameter, sink, additive) -> None:
        if source == GRAPH:
            raise GraphError("Graph inputs are not supported as parameter inputs")
        self._check_source(source, source_channel)
        idx = self.param_index(sink, parameter)
        edges = self.node_parameter_edges[sink]
        pos = next((i for i, e in enumerate(edges) if e[0] == idx), None)
        if pos is not None:
            old = edges.pop(pos)
            if additive:
                add = self.new_additive_node()
                self.node_input_edges[add][0] = (old[1], old[2])
                self.node_input_edges[add][1] = (source, source_channel)
                source, source_channel = add, 0
        edges.append((idx, source, source_channel))
        self.recalculation_required = True

    # -- graph.rs:1348-1404
    def set(self, node: int, param: Union[int, str], value, t: Time) -> None:
        idx = self.param_index(node, param)
        k, v = _value_kind(value)
        self.pending_events.append(SchedulingEvent(node, idx, k, v, None, t))

    def set_many(self, changes: Sequence[Tuple[int, Union[int, str], object]], time: Time) -> None:
        for node, param, value in changes:
            self.set(node, param, value, time)

    def schedule_bulk(self, nodes, params, value_kinds, values, frames, smooth_seconds=None) -> None:
        """Bulk equivalent of many ``Parameter.set_at(value, Seconds::from_samples(frame, sr))``
        (or ``smooth_at(Linear(s), ..)`` where smooth_seconds is not NaN) calls, in array order
        (extension for synthetic voice banks; same event semantics)."""
        n = len(nodes)
        arr = np.zeros(n, dtype=EVENT_DTYPE)
        arr["node"] = nodes
        arr["param"] = params
        arr["value_kind"] = value_kinds
        arr["value"] = values
        if smooth_seconds is not None:
            ss = np.asarray(smooth_seconds, dtype=np.float32)
            has = ~np.isnan(ss)
            arr["smoothing_kind"] = np.where(has, 2, 0)
            arr["smooth_seconds"] = np.where(has, ss, 0.0)
        fr = np.asarray(frames, dtype=np.uint64)
        sr = np.uint64(self.sample_rate)
        arr["time_kind"] = 1
        arr["seconds"] = (fr // sr).astype(np.uint32)
        arr["subsec"] = ((fr % sr) * np.uint64(SUBSECOND_TESIMALS_PER_SECOND) // sr).astype(np.uint32)
        self.flush_events_to_arrays()
        self.pending_event_arrays.append(arr)

    def flush_events_to_arrays(self) -> None:
        if self.pending_events:
            self.pending_event_arrays.append(events_to_array(self.pending_events))
            self.pending_events = []

    def take_events(self) -> np.ndarray:
        """Drain everything sent so far, in send order (what the audio thread would pop
        from the scheduling ring, graph_gen.rs:143-166)."""
        self.flush_events_to_arrays()
        if not self.pending_event_arrays:
            return np.zeros(0, dtype=EVENT_DTYPE)
        out = np.concatenate(self.pending_event_arrays) if len(self.pending_event_arrays) > 1 else self.pending_event_arrays[0]
        self.pending_event_arrays = []
        return np.ascontiguousarray(out)

    # -- graph.rs:1410
    def edit(self, c: Optional[Callable[["GraphEdit"], object]] = None):
        """``graph.edit(|g| { ... })``.  Call with a function, or use as a context manager:
        ``with graph.edit() as g: ...`` -- changes are committed on exit (Drop)."""
        if c is None:
            return GraphEdit(self)
        ge = GraphEdit(self)
        try:
            return c(ge)
        finally:
            self.commit_changes()

    # -- graph.rs:1707-1726
    def commit_changes(self) -> None:
        if self.recalculation_required:
            self.version += 1
            self.recalculation_required = False
            for cb in self._commit_listeners:
                cb(self)

    # convenience used by plan compilers / tests
    def lowered(self):
        """(node descs, input edges [(source, so_ch, sink, si_ch)], param edges
        [(source, so_ch, sink, param)], output edges [(source, so_ch, out_ch)])."""
        in_edges = []
        for sink, edges in enumerate(self.node_input_edges):
            for ch, e in enumerate(edges):
                if e is not None:
                    in_edges.append((e[0], e[1], sink, ch))
        p_edges = []
        for sink, edges in enumerate(self.node_parameter_edges):
            for (p, s, c) in edges:
                p_edges.append((s, c, sink, p))
        o_edges = [(e[0], e[1], ch) for ch, e in enumerate(self.output_edges) if e is not None]
        return [n.ugen for n in self.nodes], in_edges, p_edges, o_edges


Channels = Union[int, Sequence[int]]


def _chan_list(c: Channels) -> List[int]:
    return [int(c)] if isinstance(c, int) else [int(x) for x in c]


class GraphEdit:
    """knaster_graph/src/graph_edit.rs:77-262"""

    def __init__(self, graph: Graph):
        self.graph = graph

    def __enter__(self) -> "GraphEdit":
        return self

    def __exit__(self, exc_type, exc, tb) -> None:
        if exc_type is None:
            self.graph.commit_changes()

    def push(self, ugen: U.UGen) -> "SH":
        node = self.graph.push_internal(ugen)
        return SH(self, [(node, c) for c in range(ugen.outputs())],
                  [(node, c) for c in range(ugen.inputs())], node)

    def handle(self, node_id: int) -> Optional["SH"]:
        if node_id < 0 or node_id >= len(self.graph.nodes):
            return None
        ug = self.graph.nodes[node_id].ugen
        return SH(self, [(node_id, c) for c in range(ug.outputs())],
                  [(node_id, c) for c in range(ug.inputs())], node_id)

    def handle_from_name(self, name: str) -> Optional["SH"]:
        for i, n in enumerate(self.graph.nodes):
            if n.name == name:
                return self.handle(i)
        return None

    def set(self, node, param, value, t: Time) -> None:
        self.graph.set(int(node), param, value, t)

    def from_inputs(self, source_channels: Channels) -> "SH":
        chans = _chan_list(source_channels)
        for c in chans:
            if c >= self.graph.num_inputs:
                raise GraphError(f"GraphInputOutOfBounds({c})")
        return SH(self, [(GRAPH, c) for c in chans], [], None)


class SH:
    """Static handle (graph_edit.rs:266-463): a set of source channels and sink channels."""

    def __init__(self, edit: GraphEdit, outputs, inputs, node_id: Optional[int]):
        self._edit = edit
        self._outputs: List[Tuple[int, int]] = list(outputs)
        self._inputs: List[Tuple[int, int]] = list(inputs)
        self._node = node_id

    @property
    def _g(self) -> Graph:
        return self._edit.graph

    def __int__(self) -> int:
        return self.id()

    def id(self) -> int:
        if self._node is None:
            raise GraphError("handle does not refer to a single node")
        return self._node

    def out(self, source_channels: Channels) -> "SH":  # graph_edit.rs:280-292
        chans = _chan_list(source_channels)
        return SH(self._edit, [self._outputs[c] for c in chans], [], None)

    def to(self, n: "SH") -> "SH":  # graph_edit.rs:295-310
        if len(self._outputs) != len(n._inputs):
            raise GraphError("channel count mismatch: Inputs must be Same<Outputs>")
        for (src, sc), (sink, kc) in zip(self._outputs, n._inputs):
            self._g.connect2(src, sc, kc, sink)
        return n

    def to_replace(self, n: "SH") -> "SH":  # graph_edit.rs:330-345
        if len(self._outputs) != len(n._inputs):
            raise GraphError("channel count mismatch: Inputs must be Same<Outputs>")
        for (src, sc), (sink, kc) in zip(self._outputs, n._inputs):
            self._g.connect2_replace(src, sc, kc, sink)
        return n

    def __rshift__(self, n: "SH") -> "SH":  # graph_edit.rs:1227-1237
        return self.to(n)

    def stack(self, s: "SH") -> "SH":  # graph_edit.rs:420-430
        return SH(self._edit, self._outputs + s._outputs, self._inputs + s._inputs, None)

    def __or__(self, s: "SH") -> "SH":  # graph_edit.rs:1238-1244
        return self.stack(s)

    def to_graph_out(self) -> None:  # graph_edit.rs:363-369
        for i, (src, sc) in enumerate(self._outputs):
            self._g.connect2(src, sc, i, GRAPH)

    def to_graph_out_replace(self) -> None:  # graph_edit.rs:372-378
        for i, (src, sc) in enumerate(self._outputs):
            self._g.connect2_replace(src, sc, i, GRAPH)

    def to_graph_out_channels(self, sink_channels: Channels) -> None:  # graph_edit.rs:381-392
        chans = _chan_list(sink_channels)
        if len(chans) != len(self._outputs):
            raise GraphError("channel count mismatch")
        for (src, sc), kc in zip(self._outputs, chans):
            self._g.connect2(src, sc, kc, GRAPH)

    def to_graph_out_channels_replace(self, sink_channels: Channels) -> None:
        chans = _chan_list(sink_channels)
        for (src, sc), kc in zip(self._outputs, chans):
            self._g.connect2_replace(src, sc, kc, GRAPH)

    def link(self, p: Union[int, str], source: "SH") -> "SH":  # graph_edit.rs:735-756
        if len(source._outputs) != 1:
            raise GraphError("link expects a single-output source")
        src, sc = source._outputs[0]
        self._g.connect_replace_to_parameter(src, sc, p, self.id())
        return self

    def param(self, p: Union[int, str]) -> "Parameter":  # graph_edit.rs:763-796
        return Parameter(self._g, self.id(), self._g.param_index(self.id(), p))

    def disconnect_output(self, source_channel: int) -> None:  # graph_edit.rs:394-404
        src, sc = self._outputs[source_channel]
        self._g.disconnect_output_from_source(src, sc)

    def disconnect_input(self, sink_channel: int) -> None:  # graph_edit.rs:406-416
        sink, kc = self._inputs[sink_channel]
        self._g.disconnect_input_to_sink(kc, sink)

    def name(self, n: str) -> "SH":
        self._g.nodes[self.id()].name = n
        return self

    # arithmetic: graph_edit.rs:936-1225
    def _math(self, rhs, op: U.MathOp) -> "SH":
        g = self._g
        outs = []
        if isinstance(rhs, SH):
            if len(self._outputs) != len(rhs._outputs):
                raise GraphError("channel count mismatch: Outputs must be Same")
            for (s0, c0), (s1, c1) in zip(self._outputs, rhs._outputs):
                m = g.push_internal(U.MathUGen(1, op))
                g.connect2(s0, c0, 0, m)
                g.connect2(s1, c1, 1, m)
                outs.append((m, 0))
        else:
            c = g.push_internal(U.Constant(float(rhs)))  # one Constant, graph_edit.rs:1046-1047
            for (s0, c0) in self._outputs:
                m = g.push_internal(U.MathUGen(1, op))
                g.connect2(s0, c0, 0, m)
                g.connect2(c, 0, 1, m)
                outs.append((m, 0))
        return SH(self._edit, outs, [], None)

    def __mul__(self, rhs): return self._math(rhs, U.MathOp.Mul)
    def __add__(self, rhs): return self._math(rhs, U.MathOp.Add)
    def __sub__(self, rhs): return self._math(rhs, U.MathOp.Sub)
    def __truediv__(self, rhs): return self._math(rhs, U.MathOp.Div)
    # constant on the left keeps the UGen as operand 0 (graph_edit.rs:1183-1192): 2.0 - a == a - 2.0
    def __rmul__(self, lhs): return self._math(lhs, U.MathOp.Mul)
    def __radd__(self, lhs): return self._math(lhs, U.MathOp.Add)
    def __rsub__(self, lhs): return self._math(lhs, U.MathOp.Sub)
    def __rtruediv__(self, lhs): return self._math(lhs, U.MathOp.Div)

    def pow(self, rhs: "SH") -> "SH":  # graph_edit.rs:446-458
        return self._math(rhs, U.MathOp.Pow)


class Parameter:
    """knaster_graph/src/graph_edit.rs:1700-1886"""

    def __init__(self, graph: Graph, node: int, param_index: int):
        self._graph = graph
        self.node = node
        self.param_index = param_index

    def _send(self, value=None, smoothing=None, time: Optional[Time] = None) -> None:
        if value is None:
            k, v = 0, 0.0
        else:
            k, v = _value_kind(value)
        self._graph.pending_events.append(SchedulingEvent(self.node, self.param_index, k, v, smoothing, time))

    def set(self, value) -> None: self._send(value=value)
    def set_time(self, value, t: Time) -> None: self._send(value=value, time=t)
    def set_at(self, value, t: Seconds) -> None: self._send(value=value, time=Time.at(t))
    def set_after(self, value, t: Seconds) -> None: self._send(value=value, time=Time.after(t))
    def smooth(self, s: ParameterSmoothing) -> None: self._send(smoothing=s)
    def smooth_time(self, s: ParameterSmoothing, t: Time) -> None: self._send(smoothing=s, time=t)
    def smooth_at(self, s: ParameterSmoothing, t: Seconds) -> None: self._send(smoothing=s, time=Time.at(t))
    def smooth_after(self, s: ParameterSmoothing, t: Seconds) -> None: self._send(smoothing=s, time=Time.after(t))
    def trig(self) -> None: self._send(value=PTrigger)
    def trig_time(self, t: Time) -> None: self._send(value=PTrigger, time=t)
    def trig_at(self, t: Seconds) -> None: self._send(value=PTrigger, time=Time.at(t))
    def trig_after(self, t: Seconds) -> None: self._send(value=PTrigger, time=Time.after(t))

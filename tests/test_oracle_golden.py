"""Pin the CPU oracle against every known-answer test the reference holds for the
hot path.  Each test cites the reference test it restates.  CPU only."""
import numpy as np
import pytest

import knaster_b200 as kn
from knaster_b200.graph import GRAPH, Graph
from oracle.oracle import OracleProcessor, OracleUGen, load


def new_proc(inputs, outputs, block_size=16, sample_rate=48000, ring=50):
    g = Graph(inputs, outputs, block_size, sample_rate)
    return g, OracleProcessor(g, ring_buffer_size=ring)


# ---- knaster_core_dsp/src/wrappers_core.rs:167-200 sample_accurate_parameters_test
GOLDEN_PRECISE = [0., 0., 0., 0., 0., 5., 6., 6., 8., 9., 10., 10., 10., 10., 10., 10.]


def test_sample_accurate_parameters():
    g = OracleUGen(kn.TestInPlusParamUGen().precise_timing(10), 48000, 16)
    for d in (5, 6, 8, 9, 10):
        g.set_delay_within_block_for_param(0, d)
        g.param(0, float(d))
    out = g.process_block(np.zeros((1, 16), np.float32), 16)
    assert out[0].tolist() == GOLDEN_PRECISE


# ---- wrappers_core.rs:202-250 sample_accurate_parameters_with_wrappers_test (WrClosure is out of scope)
def test_sample_accurate_parameters_with_wrappers():
    u = kn.TestInPlusParamUGen().precise_timing(10)
    u = u.wr_add(0.0).wr_sub(0.0).wr_div(1.0).wr_mul(1.0).wr_powf(1.0).wr_powi(1)
    g = OracleUGen(u, 48000, 16)
    for d in (5, 6, 8, 9, 10):
        g.set_delay_within_block_for_param(0, d)
        g.param(0, float(d))
    out = g.process_block(np.zeros((1, 16), np.float32), 16)
    np.testing.assert_allclose(out[0], GOLDEN_PRECISE, atol=2e-4)


# ---- wrappers_core.rs:124-164 wrapper_arithmetic
def test_wrapper_arithmetic():
    def one(u):
        return float(OracleUGen(u, 48000, 4).process()[0])

    assert one(kn.TestNumUGen(2.5).wr_add(2.5)) == 5.0
    assert one(kn.TestNumUGen(2.5).wr_mul(3.0)) == 7.5
    assert one(kn.TestNumUGen(2.5).wr_div(5.0)) == 0.5
    assert one(kn.TestNumUGen(2.5).wr_v_div_gen(5.0)) == 2.0
    assert one(kn.TestNumUGen(6.0).wr_sub(7.0)) == -1.0
    assert one(kn.TestNumUGen(6.0).wr_v_sub_gen(7.0)) == 1.0
    assert abs(one(kn.TestNumUGen(6.0).wr_powf(2.0)) - 36.0) < np.finfo(np.float32).eps * 10
    assert abs(one(kn.TestNumUGen(6.0).wr_powi(2)) - 36.0) <= np.finfo(np.float32).eps * 2 * 36


# ---- knaster_core_dsp/src/ugens/math.rs:317-356 gen_arithmetics
@pytest.mark.parametrize("op,expect", [(kn.MathOp.Add, 5.0), (kn.MathOp.Sub, 1.0), (kn.MathOp.Div, 1.5),
                                       (kn.MathOp.Mul, 6.0)])
def test_gen_arithmetics(op, expect):
    m = OracleUGen(kn.MathUGen(1, op), 48000, 4)
    assert float(m.process([3.0, 2.0])[0]) == expect
    b0 = np.zeros((2, 4), np.float32)
    b0[0] = 3.0
    b0[1] = 2.0
    out = m.process_block(b0, 4)
    assert out[0].tolist() == [expect] * 4


# ---- math.rs:357-389 gen_arithmetics_multichannel: inputs laid out [a0,a1,b0,b1]
def test_gen_arithmetics_multichannel():
    m = OracleUGen(kn.MathUGen(2, kn.MathOp.Add), 48000, 4)
    assert m.process([3.0, 7.0, 2.0, 4.0]).tolist() == [5.0, 11.0]
    b0 = np.zeros((4, 4), np.float32)
    for c, v in enumerate((3.0, 7.0, 2.0, 4.0)):
        b0[c] = v
    out = m.process_block(b0, 4)
    assert out[0].tolist() == [5.0] * 4 and out[1].tolist() == [11.0] * 4


# ---- knaster_primitives/src/time.rs:461-503
def test_seconds_sample_conversion():
    import ctypes as C

    lib = load()

    def from_samples(n, sr):
        s, t = C.c_uint32(), C.c_uint32()
        lib.ko_seconds_from_samples(n, sr, C.byref(s), C.byref(t))
        return s.value, t.value

    def to_samples(st, sr):
        return lib.ko_seconds_to_samples(st[0], st[1], sr)

    def from_f64(v):
        s, t = C.c_uint32(), C.c_uint32()
        lib.ko_seconds_from_secs_f64(v, C.byref(s), C.byref(t))
        return s.value, t.value

    assert to_samples(from_samples(1, 44100), 88200) == 2
    for n in (1, 2, 3, 4):
        assert to_samples(from_samples(n, 44100), 44100) == n
    assert from_samples(22050, 44100) == from_f64(0.5)
    assert to_samples(from_samples(44100, 44100), 88200) == 88200
    assert to_samples(from_samples(44100 * 3 + 1, 44100), 88200) == 3 * 88200 + 2
    assert to_samples(from_samples(96000 * 3 + 8, 96000), 88200) == 3 * 88200 + 7
    assert to_samples((0, 0), 48000) == 0
    assert to_samples(from_f64(0.0), 48000) == 0
    # convert_to_u64_and_back
    T = 282_240_000
    as_u64 = lib.ko_seconds_to_tesimals(8347, T - 5)
    s, t = C.c_uint32(), C.c_uint32()
    lib.ko_seconds_from_tesimals(as_u64, C.byref(s), C.byref(t))
    assert (s.value, t.value) == (8347, T - 5)
    # arithmetic
    lib.ko_seconds_add(0, T - 1, 1, 1, C.byref(s), C.byref(t))
    assert (s.value, t.value) == (2, 0)
    # the host mirror agrees with the oracle (and is lossless at 48 kHz)
    for n in (0, 1, 63, 64, 47999, 48000, 479999, 123456789):
        assert from_samples(n, 48000) == (kn.Seconds.from_samples(n, 48000).seconds,
                                          kn.Seconds.from_samples(n, 48000).subsecond_tesimals)
        assert kn.Seconds.from_samples(n, 48000).to_samples(48000) == n == to_samples(from_samples(n, 48000), 48000)


# ---- knaster_graph/src/tests/graph_tests.rs:13-47
@pytest.mark.parametrize("n_out", [1, 4])
def test_graph_empty_graph_zero_output(n_out):
    _g, p = new_proc(0, n_out, ring=2)
    p.run_without_inputs()
    assert not p.output_block().any()


# ---- graph_tests.rs:49-79
def test_graph_inputs_to_outputs():
    graph, p = new_proc(3, 3)
    with graph.edit() as g:
        g.from_inputs(1).to_graph_out_channels(0)
        g.from_inputs(2).to_graph_out_channels(1)
    p.run([np.ones(16, np.float32)] * 3)
    out = p.output_block()
    assert (out[0, 0], out[1, 0], out[2, 0]) == (1.0, 1.0, 0.0)


# ---- graph_tests.rs:81-126
def test_graph_inputs_to_nodes_to_outputs():
    graph, p = new_proc(3, 3)
    with graph.edit() as g:
        g.from_inputs([0, 0]).to_graph_out_channels([1, 2])
        g0 = g.push(kn.TestInPlusParamUGen())
        g1 = g.push(kn.TestInPlusParamUGen())
        g0.param("number").set(0.75)
        g1.param("number").set(0.5)
        g0.to_graph_out_channels(2)
        g.from_inputs(2).to(g1).to_graph_out_channels(0)
    p.run([np.full(16, 2.0, np.float32)] * 3)
    out = p.output_block()
    assert (out[0, 0], out[1, 0], out[2, 0]) == (2.5, 2.0, 2.75)


# ---- graph_tests.rs:128-183 multichannel_nodes (first half; the second half edits a running graph)
def test_multichannel_nodes():
    graph, p = new_proc(3, 2)
    with graph.edit() as g:
        v0_0 = g.push(kn.TestNumUGen(0.125))
        v0_1 = g.push(kn.TestNumUGen(1.0))
        v1_0 = g.push(kn.TestNumUGen(0.5))
        v1_1 = g.push(kn.TestNumUGen(4.125))
        m = g.push(kn.MathUGen(2, kn.MathOp.Add))
        (v0_0 | v0_1 | v1_0 | v1_1).to(m).to_graph_out()
    p.run([np.ones(16, np.float32)] * 3)
    out = p.output_block()
    assert (out[0, 0], out[1, 0]) == (0.625, 5.125)
    with graph.edit() as g:
        m2 = g.push(kn.MathUGen(1, kn.MathOp.Mul))
        m3 = g.push(kn.MathUGen(1, kn.MathOp.Mul))
        (g.handle(m.id()).out(0) | g.handle(v1_0.id())).to(m2)
        (g.handle(m.id()).out(1) | g.handle(v0_0.id())).to(m3)
        (m2 | m3).to_graph_out_replace()
    p.run([np.ones(16, np.float32)] * 3)
    out = p.output_block()
    assert (out[0, 0], out[1, 0]) == (0.625 * 0.5, 5.125 * 0.125)


# ---- graph_tests.rs:256-297 disconnect  /  graph_edit.rs:2079-2122
def test_disconnect():
    graph, p = new_proc(0, 1)
    with graph.edit() as g:
        n1 = g.push(kn.TestInPlusParamUGen()).name("n1")
        g.set(n1, 0, 0.5, kn.Time.asap())
        n2 = g.push(kn.TestInPlusParamUGen()).name("n2")
        g.set(n2, 0, 1.25, kn.Time.asap())
        n3 = g.push(kn.TestInPlusParamUGen()).name("n3")
        g.set(n3, 0, 0.125, kn.Time.asap())
        (n1 >> n2 >> n3).to_graph_out()
    p.run_without_inputs()
    assert p.output_block()[0, 0] == 0.5 + 1.25 + 0.125
    with graph.edit() as g:
        g.handle_from_name("n1").disconnect_output(0)
    p.run_without_inputs()
    assert p.output_block()[0, 0] == 1.25 + 0.125
    with graph.edit() as g:
        g.handle_from_name("n3").disconnect_input(0)
    p.run_without_inputs()
    assert p.output_block()[0, 0] == 0.125


# ---- knaster_benchmarks/benches/wrappers_vs_nodes.rs:56-113: additive graph-out chain sums exactly
def test_100_wr_mul_and_100_mathgen_mul():
    graph, p = new_proc(0, 1, block_size=32)
    with graph.edit() as g:
        for _ in range(100):
            g.push(kn.TestNumUGen(2.0).wr_mul(0.5)).to_graph_out()
    p.run_without_inputs()
    assert p.output_block()[0, 31] == 100.0
    # 99 auto Add nodes in a left-fold chain (graph.rs:850-864)
    assert sum(n.auto_math_node for n in graph.nodes) == 99

    graph, p = new_proc(0, 1, block_size=32)
    with graph.edit() as g:
        for _ in range(100):
            a = g.push(kn.TestNumUGen(2.0))
            v = g.push(kn.TestNumUGen(0.5))
            (a * v).to_graph_out()
    p.run_without_inputs()
    assert p.output_block()[0, 31] == 100.0


# ---- README.md:35-47 + `sine * 0.2` lowering (graph_edit.rs:1036-1069)
def test_readme_example_structure_and_first_samples():
    graph, p = new_proc(0, 2, block_size=64)
    with graph.edit() as g:
        sine = g.push(kn.SinWt(440.0))
        sig = sine * 0.2
        sig.out([0, 0]).to_graph_out()
    kinds = [n.ugen.kind for n in graph.nodes]
    from knaster_b200 import ugens as U

    assert kinds == [U.KIND_SIN_WT, U.KIND_CONSTANT, U.KIND_MATH]
    assert graph.output_edges == [(2, 0), (2, 0)]
    p.run_without_inputs()
    out = p.output_block()
    # independent numpy restatement of osc.rs:127-130,151-156 + wavetable.rs:27-32,134-136
    table = np.sin((np.arange(16384, dtype=np.float64) / 16384.0) * np.pi * 2.0).astype(np.float32)
    k = 16384.0 * 65536.0 * (1.0 / 48000.0)
    inc = int(float(np.float32(440.0)) * k)
    phase = (np.arange(64, dtype=np.uint64) * inc) & 0xFFFFFFFF
    expect = table[(phase >> 16) & 0x3FFF] * np.float32(0.2)
    assert np.array_equal(out[0], expect) and np.array_equal(out[1], expect)


# ---- event timing through the graph (graph_gen.rs:269-305, scheduling.rs:95-121):
# an absolute event at frame 70 with block 16 lands in block 4 at in-block delay 6.
def test_absolute_event_is_sample_accurate_with_precise_timing():
    graph, p = new_proc(0, 1, block_size=16)
    with graph.edit() as g:
        n = g.push(kn.TestInPlusParamUGen().precise_timing(4))
        n.to_graph_out()
        n.param("number").set_at(3.0, kn.Seconds.from_samples(70, 48000))
    out, _ = p.render(6)
    flat = out[:, 0, :].reshape(-1)
    assert flat[:70].tolist() == [0.0] * 70 and flat[70:].tolist() == [3.0] * 26


# ---- Appendix B1: WrPreciseTiming.next_delay is sticky (precise_timing.rs:111-113, graph_gen.rs:291-293)
def test_sticky_next_delay_quirk():
    graph, p = new_proc(0, 1, block_size=16)
    with graph.edit() as g:
        n = g.push(kn.TestInPlusParamUGen().precise_timing(4))
        n.to_graph_out()
        n.param(0).set_at(1.0, kn.Seconds.from_samples(5, 48000))    # delay 5 in block 0
        n.param(0).set_at(2.0, kn.Seconds.from_samples(32, 48000))   # block-aligned: delay 0 -> reuses 5
    out, _ = p.render(4)
    flat = out[:, 0, :].reshape(-1)
    expect = [0.0] * 5 + [1.0] * (32 + 5 - 5) + [2.0] * (64 - 37)
    assert flat.tolist() == expect


# ---- without WrPreciseTiming the delay cannot be honoured: applied at the block start + a log warning
def test_event_without_precise_timing_applies_at_block_start():
    graph, p = new_proc(0, 1, block_size=16)
    with graph.edit() as g:
        n = g.push(kn.TestInPlusParamUGen())
        n.to_graph_out()
        n.param(0).set_at(1.0, kn.Seconds.from_samples(21, 48000))
    out, _ = p.render(3)
    flat = out[:, 0, :].reshape(-1)
    assert flat.tolist() == [0.0] * 16 + [1.0] * 32
    assert p.log_count() == 1


def test_envelope_builder_time_scale_and_looping():
    # envelopes.rs:386-395: the builders set time_scale / looping before init(); :423-433 the ramp
    # advances by time_scale * (1 / sr) per frame.  A 0.01 s segment at time_scale 2 ends after 240 frames.
    seg = [kn.EnvelopeSegment(0.01, 1.0)]
    outs = {}
    for ts in (1.0, 2.0):
        u = OracleUGen(kn.Envelope(0.0, seg).time_scale(ts), 48000, 64)
        u.param(2, kn.PTrigger)  # t_restart
        outs[ts] = np.concatenate([u.process_block(np.zeros((1, 64), np.float32), 64)[0] for _ in range(10)])
    assert outs[1.0][240] == pytest.approx(0.5, abs=1e-6) and outs[1.0][481] == 1.0
    assert outs[2.0][120] == pytest.approx(0.5, abs=1e-6) and outs[2.0][241] == 1.0
    assert np.array_equal(outs[2.0][:240], outs[1.0][:480:2])
    lo = OracleUGen(kn.Envelope(0.0, seg).looping(True), 48000, 64)
    lo.param(2, kn.PTrigger)
    y = np.concatenate([lo.process_block(np.zeros((1, 64), np.float32), 64)[0] for _ in range(20)])
    # :446-452 looping restarts segment and time but keeps from_value = the last target: 1 -> 1 from then on
    assert y[481] == 1.0 and np.all(y[481:] == 1.0)

"""Internal signals (VERDICT r1 "graph shapes rejected at plan time"): nodes that read the SUM of many voices (post-mix
processing -- `mix * 0.5`, a master filter; graph_edit.rs:1145-1225 over the additive chain of graph.rs:850-864) and sources
shared by many voices (one LFO into every voice's cutoff).  The plan cuts the graph at those points: the sum / the shared
source is reduced into a signal buffer one level before the voices that read it (plan.cpp, HostPlan::signal_level)."""
import numpy as np
import pytest

import knaster_b200 as kn
from knaster_b200.graph import Graph
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions
from oracle.oracle import OracleProcessor

pytestmark = pytest.mark.gpu
SR = 48000
NV = 40


def at(f):
    return kn.Seconds.from_samples(int(f), SR)


def voice(g, i, cutoff_from=None):
    saw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 70.0 + 23.0 * i).precise_timing(4))
    svf_u = kn.SvfFilter(kn.SvfFilterType.Low, 600.0 + 90.0 * i, 1.0 + 0.1 * (i % 7), 0.0)
    svf = g.push(svf_u.ar_params() if cutoff_from is not None else svf_u.precise_timing(4))
    if cutoff_from is not None:
        svf.link("cutoff_freq", cutoff_from * (200.0 + 10.0 * i) + (900.0 + 40.0 * i))
    env = g.push(kn.EnvAsr(0.004 + 0.001 * i, 0.05 + 0.002 * i).wr_mul(1.0 / NV).precise_timing(4))
    for k in range(3):
        env.param("t_restart").trig_at(at(500 + 9000 * k + 31 * i))
        saw.param("freq").set_at(90.0 + 17.0 * i + 40.0 * k, at(500 + 9000 * k + 31 * i))
        env.param("t_release").trig_at(at(500 + 9000 * k + 4000 + 13 * i))
    return (saw >> svf) * env


def mix_gain(g):
    acc = voice(g, 0)
    for i in range(1, NV):
        acc = acc + voice(g, i)
    master = acc * 0.5
    master.out([0, 0]).to_graph_out()
    return [master._outputs[0][0]]


def mix_master_filter(g):
    acc = voice(g, 0)
    for i in range(1, NV):
        acc = acc + voice(g, i)
    flt = g.push(kn.SvfFilter(kn.SvfFilterType.Low, 2500.0, 0.9, 0.0).precise_timing(4))
    flt.param("cutoff_freq").set_at(1200.0, at(7001))
    post = (acc >> flt) * 0.8
    post.to_graph_out_channels(0)
    acc.to_graph_out_channels(1)               # the dry sum on the other channel: the same Add chain read twice
    return [post._outputs[0][0]]


def shared_lfo(g):
    lfo = g.push(kn.SinWt(1.7))
    ids = []
    for i in range(NV):
        sig = voice(g, i, cutoff_from=lfo)
        sig.out([0, 0]).to_graph_out()
        ids.append(sig._outputs[0][0])
    return ids[:6]


def shared_lfo_and_master(g):
    lfo = g.push(kn.SinWt(0.9))
    acc = voice(g, 0, cutoff_from=lfo)
    for i in range(1, NV):
        acc = acc + voice(g, i, cutoff_from=lfo)
    hp = g.push(kn.OnePoleHpf())
    hp.param("cutoff_freq").set(60.0)
    post = (acc >> hp) * 0.7
    post.out([0, 0]).to_graph_out()
    return [post._outputs[0][0]]


@pytest.mark.parametrize("shape", [mix_gain, mix_master_filter, shared_lfo, shared_lfo_and_master])
@pytest.mark.parametrize("jit", [False, True])
def test_post_mix_nodes_and_shared_sources(shape, jit):
    n_blocks = 450

    def build(graph):
        with graph.edit() as g:
            return shape(g)

    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(force_jit=jit))
    ids = build(graph)
    for i in ids:
        proc.add_tap(i, 0)
    out = proc.render(n_blocks)
    taps = proc.read_taps()
    info = proc.info()
    assert info["n_voices"] >= NV + 1 and info["n_groups"] >= 2, info      # the voices, and what reads / feeds them, are separate groups
    # block-by-block == batched across the level structure
    g3, p3 = AudioProcessor.new(0, 2, AudioProcessorOptions(force_jit=jit))
    build(g3)
    blocks = []
    for _ in range(60):
        p3.run_without_inputs()
        blocks.append(p3.output_block())
    assert np.array_equal(np.stack(blocks), out[:60])
    g2 = Graph(0, 2, 64, SR)
    ids2 = build(g2)
    orc = OracleProcessor(g2, ring_buffer_size=1 << 22)
    for i in ids2:
        orc.add_tap(i, 0)
    ref, ref_taps = orc.render(n_blocks)
    assert np.abs(ref).max() > 1e-3 and np.isfinite(ref).all()
    err_bus = float(np.abs(out - ref).max())
    err_tap = float(np.abs(taps - ref_taps).max())
    print(f"{shape.__name__} jit={jit}: kernels {sorted(set(info['kernels']))}, bus err {err_bus:.2e}, tap err {err_tap:.2e}")
    assert err_bus <= 1e-5 and err_tap <= 1e-5


# ---- graph inputs: AudioProcessor::run(&[&[F]]) (processor.rs:119-141) ---------------------------------------------
def test_reference_graph_input_tests_on_gpu():
    # knaster_graph/src/tests/graph_tests.rs:49-79: inputs straight to outputs
    graph, p = AudioProcessor.new(3, 3, AudioProcessorOptions(block_size=16))
    with graph.edit() as g:
        g.from_inputs(1).to_graph_out_channels(0)
        g.from_inputs(2).to_graph_out_channels(1)
    p.run([np.ones(16, np.float32)] * 3)
    out = p.output_block()
    assert (out[0, 0], out[1, 0], out[2, 0]) == (1.0, 1.0, 0.0)
    # graph_tests.rs:81-126: inputs through nodes, and summed with nodes on an output
    graph, p = AudioProcessor.new(3, 3, AudioProcessorOptions(block_size=16))
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
    with pytest.raises(Exception):
        p.run_without_inputs()                     # processor.rs:143: run_without_inputs on a graph with inputs


def test_graph_input_through_a_filter_bank_matches_the_oracle():
    # an external signal (noise) through 24 different band filters with gain envelopes: the batched form of run()
    n_blocks = 300
    r = np.random.Generator(np.random.PCG64(9))
    x = (r.standard_normal((n_blocks, 1, 64)) * 0.3).astype(np.float32)

    def build(graph):
        ids = []
        with graph.edit() as g:
            for i in range(24):
                flt = g.push(kn.SvfFilter(kn.SvfFilterType.Band, 200.0 * (i + 1), 4.0, 0.0).precise_timing(4))
                env = g.push(kn.EnvAsr(0.01, 0.1).wr_mul(1.0 / 24))
                env.param("t_restart").trig_at(at(100 + 400 * i))
                env.param("t_release").trig_at(at(9000 + 100 * i))
                flt.param("cutoff_freq").set_at(250.0 * (i + 1), at(5000 + 7 * i))
                sig = g.from_inputs(0).to(flt) * env
                sig.out([0, 0]).to_graph_out()
                ids.append(sig._outputs[0][0])
        return ids

    graph, proc = AudioProcessor.new(1, 2, AudioProcessorOptions())
    ids = build(graph)
    for i in ids:
        proc.add_tap(i, 0)
    out = proc.render(n_blocks, inputs=x)
    taps = proc.read_taps()
    g2 = Graph(1, 2, 64, SR)
    ids2 = build(g2)
    orc = OracleProcessor(g2, ring_buffer_size=1 << 22)
    for i in ids2:
        orc.add_tap(i, 0)
    ref = np.zeros_like(out)
    ref_taps = np.zeros_like(taps)
    for b in range(n_blocks):                      # the oracle takes its inputs block by block, like knaster
        orc.run([x[b, 0]])
        ref[b] = orc.output_block()
    assert np.abs(ref).max() > 1e-3
    assert np.abs(out - ref).max() <= 1e-5

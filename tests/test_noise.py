"""Noise UGens (knaster_core_dsp/src/ugens/noise.rs).  The random stream is the fastrand 2.3.0 crate
(wyrand), which is not under the reference tree and which no reference test pins: PARITY UNPINNED for
the stream itself.  What is checked here: (CPU) the oracle's restatement of wyrand against
an independent restatement in Python integers, the distributions and state machines noise.rs
prescribes, construction-order seeding; (GPU) the CUDA engine bit-identical to the oracle."""
import numpy as np
import pytest

import knaster_b200 as kn
from knaster_b200.graph import Graph
from oracle.oracle import OracleProcessor, OracleUGen

SR = 48000
M64 = (1 << 64) - 1


def wyrand_f32(seed, n):
    """fastrand 2.3.0 Rng::f32(), restated independently of oracle/ and csrc/ (Python integers)."""
    s, out = seed, np.empty(n, np.float32)
    for i in range(n):
        s = (s + 0x2D358DCCAA6C78A5) & M64
        t = s * (s ^ 0x8BB84B93962EACC9)
        u = ((t & M64) ^ (t >> 64)) & 0xFFFFFFFF
        out[i] = np.array([0x3F800000 + (u >> 9)], np.uint32).view(np.float32)[0] - np.float32(1.0)
    return out


def block(ugen, n_blocks=8, bs=64):
    u = OracleUGen(ugen, SR, bs)
    return np.concatenate([u.process_block(np.zeros((1, bs), np.float32), bs)[0] for _ in range(n_blocks)])


def test_white_noise_is_wyrand_scaled_and_seeded_in_construction_order():
    kn.reset_randomness_seed(0)
    a, b = kn.WhiteNoise(), kn.WhiteNoise()            # seeds 0 and 1 (noise.rs:20-22)
    assert (a.args[0], b.args[0]) == (0.0, 1.0)
    for seed, u in ((0, a), (1, b)):
        ref = wyrand_f32(seed, 512) * np.float32(2.0) - np.float32(1.0)
        assert np.array_equal(block(u), ref)
    kn.reset_randomness_seed(0)
    assert kn.RandomLin(2.0).args[1] == 53.0 and kn.RandomLin(2.0).args[1] == 147.0   # seed * 94 + 53 (:173)


def test_white_noise_distribution():
    kn.reset_randomness_seed(7)
    y = block(kn.WhiteNoise(), 1024)
    assert y.min() >= -1.0 and y.max() < 1.0
    assert abs(y.mean()) < 0.02 and abs(y.var() - 1.0 / 3.0) < 0.01
    assert abs(np.corrcoef(y[:-1], y[1:])[0, 1]) < 0.02


def test_pink_noise_structure():
    kn.reset_randomness_seed(3)
    seed = 3
    y = block(kn.PinkNoise(), 64)
    # replay noise.rs:94-114 on the independent stream
    w = wyrand_f32(seed, 2 * len(y)) * np.float32(2.0) - np.float32(1.0)
    white, always, counter, pink = np.zeros(9, np.float32), np.float32(0), 1, np.float32(0)
    ref = np.empty_like(y)
    for i in range(len(y)):
        idx = (counter & -counter).bit_length() - 1
        pink = np.float32(pink - white[idx]); white[idx] = w[2 * i]; pink = np.float32(pink + white[idx])
        pink = np.float32(pink - always); always = w[2 * i + 1]; pink = np.float32(pink + always)
        counter = (counter & 255) + 1
        ref[i] = np.float32(pink / np.float32(10.0))
    assert np.array_equal(y, ref)
    assert np.abs(y).max() < 1.0
    spec = np.abs(np.fft.rfft(y[:4096] * np.hanning(4096))) ** 2
    assert spec[4:40].mean() > 8 * spec[400:2000].mean()       # energy falls with frequency


def test_brown_noise_is_clamped_integrated_white():
    kn.reset_randomness_seed(11)
    y = block(kn.BrownNoise(), 256)
    w = wyrand_f32(11, len(y)) * np.float32(2.0) - np.float32(1.0)
    last, ref = np.float32(0), np.empty_like(y)
    for i in range(len(y)):
        last = np.float32(last + np.float32(w[i] * np.float32(0.1)))
        last = min(max(last, np.float32(-1)), np.float32(1))
        ref[i] = last
    assert np.array_equal(y, ref)
    assert np.abs(y).max() <= 1.0 and np.abs(np.diff(y)).max() <= 0.1 + 1e-6


def test_random_lin_interpolates_between_uniform_values():
    kn.reset_randomness_seed(0)
    u = kn.RandomLin(100.0)                                # a new value every 480 frames
    y = block(u, 64)
    assert y.min() >= 0.0 and y.max() <= 1.0
    d2 = np.abs(np.diff(y, 2))
    kinks = np.nonzero(d2 > 1e-5)[0]
    assert 7 <= len(kinks) <= 9 and np.all(np.abs(np.diff(kinks) - 480) <= 1)
    r = wyrand_f32(53, 3)
    assert y[0] == r[0]                                     # new(): current_value; init(): first target r[1]
    assert abs(y[480] - r[1]) < 3e-3


def noise_graph(graph):
    kn.reset_randomness_seed(0)
    ids = []
    with graph.edit() as g:
        for i in range(12):
            w = g.push(kn.WhiteNoise().wr_mul(0.5))
            p = g.push(kn.PinkNoise())
            b = g.push(kn.BrownNoise().precise_timing(4))
            r = g.push(kn.RandomLin(50.0 + 30.0 * i).precise_timing(4))
            f = g.push(kn.SvfFilter(kn.SvfFilterType.Low, 800.0 + 100.0 * i, 1.5, 0.0))
            sig = (w >> f) * r + p * 0.25 + b * 0.1
            sig.out([0, 0]).to_graph_out()
            r.param("freq").set_at(400.0, kn.Seconds.from_samples(3001 + i, SR))
            w.param("wr_mul").set_at(0.25, kn.Seconds.from_samples(5000, SR))
            ids += [w.id(), p.id(), b.id(), r.id(), sig._outputs[0][0]]
    return ids


def test_noise_graph_renders_in_the_oracle():
    g = Graph(0, 2, 64, SR)
    ids = noise_graph(g)
    orc = OracleProcessor(g, ring_buffer_size=1 << 20)
    for i in ids:
        orc.add_tap(i, 0)
    out, taps = orc.render(100)
    assert np.isfinite(out).all() and np.abs(out).max() > 0.1
    assert np.abs(taps[0]).max() <= 0.5 and taps[3].min() >= 0.0


@pytest.mark.gpu
def test_noise_ugens_bit_identical_on_the_gpu():
    from knaster_b200.processor import AudioProcessor, AudioProcessorOptions

    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(sample_rate=SR))
    ids = noise_graph(graph)
    ev = graph.take_events()
    graph.pending_event_arrays = [ev.copy()]
    for i in ids:
        proc.add_tap(i, 0)
    gpu = proc.render(150)
    gt = proc.read_taps()
    g2 = Graph(0, 2, 64, SR)
    ids2 = noise_graph(g2)
    g2.take_events()
    g2.pending_event_arrays = [ev.copy()]
    orc = OracleProcessor(g2, ring_buffer_size=1 << 20)
    for i in ids2:
        orc.add_tap(i, 0)
    ref, rt = orc.render(150)
    for k in range(len(ids)):
        if k % 5 != 4:
            assert np.array_equal(gt[k], rt[k]), f"noise node tap {k} differs"   # integer RNG + single f32 ops
    assert np.abs(gt - rt).max() <= 1e-4      # through the SVF
    assert np.abs(gpu - ref).max() <= 1e-5
    # rendering on: the RNG state persists across render calls
    gpu2 = proc.render(50)
    ref2, _ = orc.render(50)
    assert np.abs(gpu2 - ref2).max() <= 1e-5 and np.abs(gpu2).max() > 0.05

"""Pan2 (knaster_core_dsp/src/ugens/pan.rs).  Its gains are fastapprox 0.3.1's fast::cos / fast::sin, a crate
outside the reference tree that no reference test pins: PARITY UNPINNED for those two functions (they are
restated from the crate's published source in the oracle and, independently, on the engine's host side).
Checked here: the pan law's shape, the `pan` parameter, and the engine against the oracle on the reference's
own `many_sines` voice shape (knaster/examples/many_sines.rs:51-60: (EnvAr * SinWt.wr_mul) >> Pan2 -> out)."""
import math

import numpy as np
import pytest

import knaster_b200 as kn
from knaster_b200.graph import Graph
from oracle.oracle import OracleProcessor, OracleUGen

SR = 48000


def gains(pan):
    u = OracleUGen(kn.Pan2(pan), SR, 16)
    out = u.process_block(np.ones((1, 16), np.float32), 16)
    return float(out[0][0]), float(out[1][0])


def test_pan_law_is_the_fast_cos_sin_of_a_quarter_turn():
    for pan in (-1.0, -0.5, 0.0, 0.3, 1.0):
        l, r = gains(pan)
        a = (pan * 0.5 + 0.5) * math.pi / 2
        assert abs(l - math.cos(a)) < 2e-3 and abs(r - math.sin(a)) < 2e-3   # fastapprox: ~1e-3 absolute
    l, r = gains(0.0)
    assert abs(l - r) < 2e-3 and abs(l * l + r * r - 1.0) < 4e-3


def many_sines(graph, n=24):
    ids = []
    with graph.edit() as g:
        for i in range(n):
            env = g.push(kn.EnvAr(0.01, 0.1))
            sine = g.push(kn.SinWt(300.0 + 37.0 * i).wr_mul(0.01 + 0.001 * i))
            pan = g.push(kn.Pan2(-1.0 + 2.0 * i / (n - 1)))
            sig = (env * sine) >> pan
            sig.to_graph_out()
            env.param("t_restart").trig_at(kn.Seconds.from_samples(100 + 40 * i, SR))
            env.param("t_restart").trig_at(kn.Seconds.from_samples(6000 + 17 * i, SR))
            pan.param("pan").set_at(0.9 - 0.07 * i, kn.Seconds.from_samples(3000 + i, SR))
            ids.append(pan.id())
    return ids


def test_many_sines_shape_renders_in_the_oracle():
    g = Graph(0, 2, 64, SR)
    many_sines(g)
    out, _ = OracleProcessor(g, ring_buffer_size=1 << 20).render(150)
    assert np.isfinite(out).all()
    assert np.abs(out[:, 0]).max() > 0.01 and np.abs(out[:, 1]).max() > 0.01
    assert not np.array_equal(out[:, 0], out[:, 1])


@pytest.mark.gpu
def test_many_sines_shape_on_the_gpu():
    from knaster_b200.processor import AudioProcessor, AudioProcessorOptions

    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(sample_rate=SR))
    ids = many_sines(graph)
    ev = graph.take_events()
    graph.pending_event_arrays = [ev.copy()]
    for i in ids:
        proc.add_tap(i, 0)
        proc.add_tap(i, 1)
    gpu = proc.render(150)
    gt = proc.read_taps()
    g2 = Graph(0, 2, 64, SR)
    ids2 = many_sines(g2)
    g2.take_events()
    g2.pending_event_arrays = [ev.copy()]
    orc = OracleProcessor(g2, ring_buffer_size=1 << 20)
    for i in ids2:
        orc.add_tap(i, 0)
        orc.add_tap(i, 1)
    ref, rt = orc.render(150)
    assert np.array_equal(gt, rt)                  # integer-phase table lookup, f32 products, host-evaluated gains
    assert np.abs(gpu - ref).max() <= 1e-6
    assert np.abs(ref).max() > 0.01

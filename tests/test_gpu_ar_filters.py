"""Audio-rate routes into filter parameters (WrArParams, audio_rate.rs:42-57 over svf.rs:81-157 / onepole.rs:35-46,135-139): the
canonical subtractive patch -- an LFO or an envelope driving the cutoff.  knaster calls the parameter setter every frame, so
the coefficients (tanf / expf) are recomputed per sample; the engine does the same on the device (nodes.cuh svf_coeffs_dev)
with the correctly rounded f64 functions, held to the 1e-4 filter budget over 10 s."""
import numpy as np
import pytest

import knaster_b200 as kn
from knaster_b200.graph import Graph
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions
from oracle.oracle import OracleProcessor

pytestmark = pytest.mark.gpu
SR = 48000


def at(f):
    return kn.Seconds.from_samples(int(f), SR)


def lfo_cutoff(g, i):
    saw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 80.0 + 37.0 * i))
    lfo = g.push(kn.SinWt(0.3 + 0.7 * i))
    svf = g.push(kn.SvfFilter(kn.SvfFilterType(i % 9), 1000.0, 0.7 + 0.6 * (i % 4), 3.0).ar_params())
    svf.link("cutoff_freq", lfo * (400.0 + 100.0 * i) + (900.0 + 150.0 * i))
    return (saw >> svf) * (1.0 / 12)


def env_cutoff(g, i):
    saw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 55.0 * (i + 1)))
    env = g.push(kn.EnvAsr(0.01 + 0.02 * i, 0.3))
    svf = g.push(kn.SvfFilter(kn.SvfFilterType.Low, 500.0, 2.0 + i, 0.0).ar_params())
    svf.link("cutoff_freq", env * 6000.0 + 200.0)
    for k in range(8):
        env.param("t_restart").trig_at(at(1000 + 60000 * k + 17 * i))
        env.param("t_release").trig_at(at(1000 + 60000 * k + 20000))
    return (saw >> svf) * (1.0 / 12)


def lfo_q_and_onepole(g, i):
    saw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 110.0 + 13.0 * i))
    lfo = g.push(kn.SinWt(2.0 + i))
    svf = g.push(kn.SvfFilter(kn.SvfFilterType.Band, 700.0 + 300.0 * i, 1.0, 0.0).ar_params())
    svf.link("q", lfo * 0.4 + 1.5)
    svf.param("cutoff_freq").set_at(1500.0 + 100.0 * i, at(30000 + i))        # a plain event beside the routed parameter
    lp = g.push(kn.OnePoleLpf(2000.0).ar_params())
    lp.link("cutoff_freq", lfo * 800.0 + 1800.0)
    return (saw >> svf >> lp) * (1.0 / 12)


def lfo_envelope_times(g, i):
    """envelopes.rs:84-111 under WrArParams: attack_time / release_time re-set every frame from an LFO"""
    saw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 70.0 + 29.0 * i))
    lfo = g.push(kn.SinWt(1.5 + 0.5 * i))
    env = g.push((kn.EnvAsr(0.05, 0.2) if i % 2 == 0 else kn.EnvAr(0.05, 0.2)).ar_params())
    env.link("attack_time", lfo * 0.02 + 0.03)
    env.link("release_time", lfo * 0.1 + 0.25)
    for k in range(8):
        env.param("t_restart").trig_at(at(1000 + 60000 * k + 17 * i))
        if i % 2 == 0:
            env.param("t_release").trig_at(at(1000 + 60000 * k + 20000))
    return (saw * env) * (1.0 / 12)


@pytest.mark.parametrize("voice", [lfo_cutoff, env_cutoff, lfo_q_and_onepole, lfo_envelope_times])
@pytest.mark.parametrize("jit", [False, True])
def test_audio_rate_routes_into_filter_parameters(voice, jit):
    n_blocks = 7500                                  # 10 s

    def build(graph):
        ids = []
        with graph.edit() as g:
            for i in range(12):
                sig = voice(g, i)
                sig.out([0, 0]).to_graph_out()
                ids.append(sig._outputs[0][0])
        return ids

    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(force_jit=jit))
    ids = build(graph)
    for i in ids:
        proc.add_tap(i, 0)
    out = proc.render(n_blocks)
    taps = proc.read_taps()
    kernels = proc.info()["kernels"]
    assert all(k == ("render_jit" if jit else "render_interp") for k in kernels), kernels
    g2 = Graph(0, 2, 64, SR)
    ids2 = build(g2)
    orc = OracleProcessor(g2, ring_buffer_size=1 << 22)
    for i in ids2:
        orc.add_tap(i, 0)
    ref, ref_taps = orc.render(n_blocks)
    assert np.isfinite(ref_taps).all() and np.abs(ref_taps).max() > 1e-2
    scale = 12.0 / np.maximum(1.0, 12.0 * np.abs(ref_taps).max(axis=1, keepdims=True))   # unit voice gain; resonant peaks relative
    err = np.abs(taps - ref_taps) * scale
    print(f"{voice.__name__} jit={jit}: max err at unit voice gain {err.max():.3e} (last second {err[:, -SR:].max():.3e})")
    assert err.max() <= 1e-4
    assert np.abs(out - ref).max() <= 1e-5 * max(1.0, float(np.abs(ref).max()) * 12)

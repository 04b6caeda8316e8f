"""Randomised parity: seeded random voice shapes (every supported source, filter, envelope, wrapper and
audio-rate route, with random parameter events) rendered by the CUDA engine (interpreter, through the C
ABI) and by the oracle.  One graph holds ~20 different voice templates, i.e. ~20 kernel groups.  The
shapes are drawn from what the engine documents as supported; a shape it rejects is a test failure."""
import numpy as np
import pytest

import knaster_b200 as kn
from test_gpu_parity import SR, both

pytestmark = pytest.mark.gpu

N_BLOCKS = 120            # 7680 frames
N_FRAMES = N_BLOCKS * 64


def at(frame):
    return kn.Seconds.from_samples(int(frame), SR)


def random_voice(g, r, vi):
    """Push one random voice into GraphEdit `g`; returns the SH of its output."""
    def frames(n):
        return sorted(int(x) for x in r.integers(0, N_FRAMES, n))

    def wrap(u, allow_smooth=True):
        k = r.integers(0, 5)
        if k == 1:
            u = u.wr_mul(float(r.uniform(0.3, 1.2)))
        elif k == 2:
            u = u.wr_add(float(r.uniform(-0.2, 0.2)))
        elif k == 3:
            u = u.wr_mul(float(r.uniform(0.5, 1.0))).wr_sub(float(r.uniform(-0.1, 0.1)))
        t = r.integers(0, 4)
        if t == 1 and allow_smooth:
            u = u.smooth_params()
        elif t == 2:
            u = u.precise_timing(int(r.integers(2, 9)))
        elif t == 3 and allow_smooth:
            u = u.smooth_params().precise_timing(int(r.integers(2, 9)))
        return u

    f0 = float(r.uniform(40.0, 3000.0))
    sk = int(r.integers(0, 9))
    ar = r.random() < 0.3
    lfo = None
    if sk == 0:
        src = g.push(kn.SinWt(f0).ar_params() if ar else wrap(kn.SinWt(f0)))
        fpar = "freq"
    elif sk == 1:
        src = g.push(kn.SinNumeric(f0).ar_params() if ar else wrap(kn.SinNumeric(f0)))
        fpar = "freq"
    elif sk in (2, 3):
        wf = kn.Waveform(int(r.integers(0, 14)))
        src = g.push(kn.PolyBlep(wf, f0).ar_params() if ar else wrap(kn.PolyBlep(wf, f0)))
        fpar = "freq" if r.random() < 0.7 else "pulse_width"
    elif sk == 4:
        src, ar, fpar = g.push(wrap(kn.WhiteNoise(), allow_smooth=False)), False, None
    elif sk == 5:
        src, ar, fpar = g.push(wrap(kn.PinkNoise(), allow_smooth=False)), False, None
    elif sk == 6:
        src, ar, fpar = g.push(wrap(kn.BrownNoise(), allow_smooth=False)), False, None
    elif sk == 7:
        src, ar, fpar = g.push(wrap(kn.RandomLin(float(r.uniform(5.0, 900.0))))), False, "freq"
    else:
        src, ar, fpar = g.push(wrap(kn.Phasor(f0))), False, "freq"
    if ar:
        lfo = g.push(kn.SinWt(float(r.uniform(1.0, 40.0))))
        if fpar == "pulse_width":
            src.link(fpar, lfo * 0.3 + 0.5)
        else:
            src.link(fpar, lfo * float(r.uniform(0.0, 0.5) * f0) + f0)
    elif fpar == "freq":
        for fr in frames(int(r.integers(0, 4))):
            src.param("freq").set_at(float(r.uniform(30.0, 4000.0)), at(fr))
    elif fpar == "pulse_width":
        for fr in frames(int(r.integers(0, 3))):
            src.param("pulse_width").set_at(float(r.uniform(0.1, 0.9)), at(fr))

    sig = src
    for _ in range(int(r.integers(0, 3))):
        pk = int(r.integers(0, 4))
        if pk == 0:
            ty = kn.SvfFilterType(int(r.integers(0, 9)))
            flt = g.push(wrap(kn.SvfFilter(ty, float(r.uniform(80.0, 9000.0)), float(r.uniform(0.5, 6.0)), float(r.uniform(-6.0, 6.0)))))
            for fr in frames(int(r.integers(0, 3))):
                flt.param("cutoff_freq").set_at(float(r.uniform(80.0, 9000.0)), at(fr))
            if r.random() < 0.3:
                flt.param("q").set_at(float(r.uniform(0.5, 6.0)), at(frames(1)[0]))
        elif pk == 1:
            flt = g.push(wrap(kn.OnePoleLpf(float(r.uniform(50.0, 8000.0)))))
            for fr in frames(int(r.integers(0, 2))):
                flt.param("cutoff_freq").set_at(float(r.uniform(50.0, 8000.0)), at(fr))
        elif pk == 2:
            flt = g.push(kn.OnePoleHpf())
            flt.param("cutoff_freq").set(float(r.uniform(20.0, 2000.0)))
        else:
            flt = g.push(kn.Math1UGen(kn.Math1Op(int(r.choice([0, 2, 3, 4])))))   # Ceil, Floor, Trunc, Fract
        sig = sig >> flt

    ek = int(r.integers(0, 4))
    if ek:
        a, rel = float(r.uniform(0.001, 0.03)), float(r.uniform(0.005, 0.08))
        if ek == 1:
            env = g.push(wrap(kn.EnvAsr(a, rel), allow_smooth=False))
            ons = frames(int(r.integers(1, 4)))
            for fr in ons:
                env.param("t_restart").trig_at(at(fr))
                env.param("t_release").trig_at(at(fr + int(r.integers(50, 2500))))
        elif ek == 2:
            env = g.push(wrap(kn.EnvAr(a, rel), allow_smooth=False))
            for fr in frames(int(r.integers(1, 4))):
                env.param("t_restart").trig_at(at(fr))
        else:
            segs = [kn.EnvelopeSegment(a, 1.0), kn.EnvelopeSegment(float(r.uniform(0.005, 0.03)), float(r.uniform(0.2, 0.8))),
                    kn.EnvelopeSegment(rel, 0.0)]
            envu = kn.Envelope(0.0, segs)
            if r.random() < 0.3:
                envu = envu.looping(True)
            env = g.push(wrap(envu, allow_smooth=False))
            for fr in frames(int(r.integers(1, 4))):
                env.param("t_restart").trig_at(at(fr))
                if r.random() < 0.4:
                    env.param("t_stop").trig_at(at(fr + int(r.integers(20, 1500))))
                if r.random() < 0.3:
                    env.param("jump_to_segment").set_at(int(r.integers(0, 3)), at(fr + int(r.integers(1500, 3000))))
        sig = sig * env
    if r.random() < 0.5:
        sig = sig * float(r.uniform(0.1, 0.9))
    return sig


@pytest.mark.parametrize("seed", list(range(11, 27)))
def test_random_voice_shapes_match_the_oracle(seed):
    def build(graph):
        kn.reset_randomness_seed(0)
        r = np.random.Generator(np.random.PCG64(seed))
        ids = []
        with graph.edit() as g:
            for vi in range(20):
                sig = random_voice(g, r, vi)
                sig.out([0, 0]).to_graph_out()
                ids.append(sig._outputs[0][0])
        return ids

    gpu, ref, gt, rt, proc = both(build, N_BLOCKS)
    assert np.isfinite(rt).all(), "the generator produced a non-finite reference: narrow its ranges"
    scale = np.maximum(1.0, np.abs(rt).max(axis=1, keepdims=True))
    err = np.abs(gt - rt) / scale
    worst = int(np.argmax(err.max(axis=1)))
    assert err.max() <= 1e-4, f"seed {seed}: voice {worst} differs by {err.max():.3e} (first at frame {int(np.argmax(err[worst] > 1e-4))})"
    assert np.abs(gpu - ref).max() <= 1e-5 * max(1.0, float(np.abs(ref).max()))
    assert proc.info()["dropped_changes"] == 0

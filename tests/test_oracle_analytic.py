"""The oracle against ANALYTIC truths -- an anchor for the DSP bodies that is independent of anyone's reading of the
reference (no reference test pins them; knaster itself cannot be built in this image or on the GPU box:
neither has cargo / rustc).  Each check states the closed form it uses.

  SvfFilter   magnitude and phase of the steady-state response against the transfer function of the trapezoidal
              (Cytomic / Simper) SVF: the analogue prototype H(s) with s = j tan(pi f / sr) / tan(pi fc / sr), k = 1 / q
  OnePoleLpf  against H(z) = a0 / (1 - b1 z^-1), b1 = exp(-2 pi fc / sr)
  EnvAsr      a 1 s attack ends after 47 970 samples, not 48 000 (f32 accumulation of 1 / 48000);
              the release is t^3 * scale
  PolyBlep    sawtooth: no DC, harmonics 2 / (pi k)
  SinWt       within one table step (2 pi / 16384) of the f64 sine over a second: no interpolation (wavetable.rs:322-324)
  SinNumeric  within 2e-5 of the f64 sine over 10 s of accumulated f32 phase
"""
import numpy as np
import pytest

import knaster_b200 as kn
from oracle.oracle import OracleUGen

SR = 48000


def run_source(ugen, n, block=64, setup=None):
    u = OracleUGen(ugen, SR, block)
    if setup:
        setup(u)
    out = [u.process_block(np.zeros((1, block), np.float32), block)[0] for _ in range(n // block)]
    return np.concatenate(out).astype(np.float64)


def run_filter(ugen, x, block=64):
    u = OracleUGen(ugen, SR, block)
    out = [u.process_block(x[i:i + block].astype(np.float32)[None, :], block)[0] for i in range(0, len(x), block)]
    return np.concatenate(out).astype(np.float64)


def fit(y, f, n0):
    """Amplitude and phase of the component at frequency f in y[n0:] (least squares on sin / cos)."""
    n = np.arange(n0, len(y))
    w = 2 * np.pi * f / SR
    A = np.stack([np.sin(w * n), np.cos(w * n)], 1)
    (a, b), *_ = np.linalg.lstsq(A, y[n0:], rcond=None)
    return np.hypot(a, b), np.arctan2(b, a)


def svf_h(ty, f, fc, q):
    s = 1j * np.tan(np.pi * f / SR) / np.tan(np.pi * fc / SR)
    k = 1.0 / q
    den = s * s + k * s + 1.0
    low, band, high = 1.0 / den, s / den, s * s / den
    return {"Low": low, "High": high, "Band": band, "Notch": low + high, "Peak": high - low, "All": 1.0 - 2.0 * k * band}[ty]


@pytest.mark.parametrize("ty", ["Low", "High", "Band", "Notch", "Peak", "All"])
@pytest.mark.parametrize("fc,q", [(1000.0, 0.7071), (250.0, 4.0), (6000.0, 1.5), (12000.0, 8.0)])
def test_svf_frequency_response_matches_the_trapezoidal_svf_transfer_function(ty, fc, q):
    n = 48000
    for f in (110.0, 0.5 * fc, fc, 1.7 * fc if 1.7 * fc < 20000 else 19000.0, 15000.0):
        x = np.sin(2 * np.pi * f * np.arange(n) / SR)
        y = run_filter(kn.SvfFilter(getattr(kn.SvfFilterType, ty), fc, q, 0.0), x)
        amp, ph = fit(y, f, n // 2)                     # the transient has died: poles at radius < 0.999 for these (fc, q)
        h = svf_h(ty, f, fc, q)
        assert abs(amp - abs(h)) <= 2e-4 * max(1.0, abs(h)), (ty, fc, q, f, amp, abs(h))
        if abs(h) > 1e-2:
            dphi = np.angle(np.exp(1j * (ph - np.angle(h))))
            assert abs(dphi) <= 2e-3, (ty, fc, q, f, ph, np.angle(h))


@pytest.mark.parametrize("fc", [50.0, 800.0, 9000.0])
def test_onepole_lowpass_frequency_response(fc):
    b1 = np.exp(-2 * np.pi * fc / SR)
    a0 = 1.0 - b1
    n = 48000
    for f in (30.0, fc, 4 * fc if 4 * fc < 20000 else 18000.0):
        x = np.sin(2 * np.pi * f * np.arange(n) / SR)
        y = run_filter(kn.OnePoleLpf(fc), x)
        amp, ph = fit(y, f, n // 2)
        h = a0 / (1.0 - b1 * np.exp(-2j * np.pi * f / SR))
        assert abs(amp - abs(h)) <= 1e-4
        assert abs(np.angle(np.exp(1j * (ph - np.angle(h))))) <= 1e-3


def test_envasr_one_second_attack_ends_after_47970_samples_and_releases_as_a_cube():
    def setup(u):
        u.param(3, kn.PTrigger())                        # t_restart

    y = run_source(kn.EnvAsr(1.0, 0.5), 48000 + 64, setup=setup)
    first_one = int(np.argmax(y >= 1.0))
    assert first_one == 47970                           # f32 accumulation of 1 / 48000 overshoots early
    assert np.all(np.diff(y[:first_one]) > 0) and y[0] == 0.0
    assert abs(y[24000] - 0.5) < 2e-3                   # linear, with the accumulated f32 drift the survey measured
    # release: t^3 * scale from t = 1, scale = the level at release
    u = OracleUGen(kn.EnvAsr(0.01, 0.25), SR, 64)
    u.param(3, kn.PTrigger())
    pre = np.concatenate([u.process_block(np.zeros((1, 64), np.float32), 64)[0] for _ in range(20)])
    assert pre[-1] == 1.0                               # sustaining
    u.param(2, kn.PTrigger())                            # t_release
    rel = np.concatenate([u.process_block(np.zeros((1, 64), np.float32), 64)[0] for _ in range(200)]).astype(np.float64)
    t = 1.0 - np.arange(len(rel)) / (0.25 * SR)
    want = np.where(t > 0, t ** 3, 0.0)
    assert np.abs(rel - want).max() < 1e-3
    assert rel[int(0.25 * SR) + 8] == 0.0


@pytest.mark.parametrize("f0", [110.0, 440.0, 1500.0])
def test_polyblep_sawtooth_has_no_dc_and_harmonics_two_over_pi_k(f0):
    n = 48000
    y = run_source(kn.PolyBlep(kn.Waveform.Sawtooth, f0), n)
    periods = int(f0 * n / SR)                          # f0 divides 48000 / s evenly for these: whole periods
    assert abs(y[: int(periods * SR / f0)].mean()) < 2e-3
    for k in (1, 2, 3, 5):
        amp, _ = fit(y, k * f0, 0)
        # the 2-sample polynomial BLEP residual smooths the step like a short triangular window: the k-th harmonic loses
        # (pi k f0 / sr)^2 / 3 of its level to first order (2.9 % at 4.5 kHz)
        x = np.pi * k * f0 / SR
        want = 2.0 / (np.pi * k) * (1.0 - x * x / 3.0)
        assert abs(amp - want) <= 0.004 * want, (f0, k, amp, want)


def test_sinwt_is_the_f64_sine_to_one_table_step():
    for f in (55.0, 440.0, 3333.3):
        y = run_source(kn.SinWt(f), 48000)
        ff = float(np.float32(f))
        want = np.sin(2 * np.pi * ff * np.arange(48000) / SR)
        # nearest-below lookup: at most one table step of phase (2 pi / 16384 = 3.8e-4), plus the truncation of the u32
        # increment (osc.rs:127-130): up to 2^-30 of a cycle per sample, i.e. 2 pi n / 2^30 of phase after n samples
        bound = 2 * np.pi / 16384 + 2 * np.pi * np.arange(48000) / 2.0 ** 30 + 1e-6
        assert np.all(np.abs(y - want) <= bound)
    assert np.abs(run_source(kn.SinWt(440.0), 4800)).max() > 0.999


def test_sinnumeric_tracks_the_f64_sine_over_ten_seconds():
    f = 440.0
    n = 480000
    y = run_source(kn.SinNumeric(f), n)
    # the f32 phase accumulates rounding (that drift is part of knaster's render); the WAVEFORM at the
    # accumulated phase is the libm sine: compare against the f64 sine of the same f32 phase sequence
    inc = np.float32(f) / np.float32(SR)
    ph = np.empty(n, np.float32)
    p = np.float32(0.0)
    for i in range(n):
        ph[i] = p
        p = np.float32(p + inc)
        if p > np.float32(1.0):
            p = np.float32(p - np.float32(1.0))
    want = np.sin(np.float64(np.float32(ph * np.float32(2 * np.pi))))
    assert np.abs(y - want).max() <= 2e-7
    # and the drift against the ideal sine stays inside what the survey measured for a 440 Hz phase (< 3e-2 over 10 s)
    ideal = np.sin(2 * np.pi * f * np.arange(n) / SR)
    assert np.abs(y - ideal).max() < 3e-2

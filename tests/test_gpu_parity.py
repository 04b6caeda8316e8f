"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on the same
graph + events.  Tolerances are the north star's: bit-exact for integer-phase / pure
arithmetic paths, max abs error <= 1e-5 for oscillators, <= 1e-4 for IIR filters (the f32
recurrences are evaluated sequentially in the reference's rounding order, so in practice
they are far tighter).  The mix bus is summed in a different (tree) order than knaster's
left fold, so bus comparisons use normalised amplitudes and <= 1e-5."""
import numpy as np
import pytest

import knaster_b200 as kn
from knaster_b200 import banks
from knaster_b200.graph import Graph
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions
from oracle.oracle import OracleProcessor

pytestmark = pytest.mark.gpu
SR = 48000


def both(build, n_blocks, outputs=2, block_size=64, taps=True, force_interpreter=False, no_scan=False):
    """Run `build(graph) -> tap node ids` through the GPU engine and the oracle."""
    opts = AudioProcessorOptions(block_size=block_size, sample_rate=SR, force_interpreter=force_interpreter, no_scan=no_scan)
    graph, proc = AudioProcessor.new(0, outputs, opts)
    ids = build(graph)
    ev = graph.take_events()
    graph.pending_event_arrays = [ev.copy()] if len(ev) else []
    if taps:
        for i in ids:
            proc.add_tap(i, 0)
    gpu = proc.render(n_blocks)
    gpu_taps = proc.read_taps() if (taps and ids) else None

    g2 = Graph(0, outputs, block_size, SR)
    ids2 = build(g2)
    g2.take_events()
    g2.pending_event_arrays = [ev.copy()] if len(ev) else []
    orc = OracleProcessor(g2, ring_buffer_size=1 << 22)
    if taps:
        for i in ids2:
            orc.add_tap(i, 0)
    ref, ref_taps = orc.render(n_blocks)
    return gpu, ref, gpu_taps, ref_taps, proc


@pytest.mark.parametrize("force_interp", [False, True])
def test_readme_sine_bit_identical(force_interp):
    # configs[0]: SinWt 440 Hz * 0.2 -> stereo, block 64, f32.  Integer phase => bit-identical.
    gpu, ref, gt, rt, _ = both(banks.readme_sine, 750, force_interpreter=force_interp)
    assert np.array_equal(gpu, ref)
    assert np.array_equal(gt, rt)
    assert np.abs(ref).max() > 0.19


def test_run_without_inputs_block_by_block_equals_batched_render():
    def build(graph):
        return banks.subtractive_bank(graph, 8, 0.2, n_notes=3)

    opts = AudioProcessorOptions()
    g1, p1 = AudioProcessor.new(0, 2, opts)
    build(g1)
    blocks = []
    for _ in range(150):
        p1.run_without_inputs()
        blocks.append(p1.output_block())
    g2, p2 = AudioProcessor.new(0, 2, opts)
    build(g2)
    batched = p2.render(150)
    assert np.array_equal(np.stack(blocks), batched)
    assert p1.frame_clock() == 150 * 64 == p2.frame_clock()


def test_reference_graph_tests_on_gpu():
    # knaster_graph/src/tests/graph_tests.rs:13-47: empty graph -> zeros
    _g, p = AudioProcessor.new(0, 4, AudioProcessorOptions(block_size=16))
    p.run_without_inputs()
    assert not p.output_block().any()
    # knaster_benchmarks/benches/wrappers_vs_nodes.rs:65-113
    graph, p = AudioProcessor.new(0, 1, AudioProcessorOptions(block_size=32))
    with graph.edit() as g:
        for _ in range(100):
            g.push(kn.TestNumUGen(2.0).wr_mul(0.5)).to_graph_out()
    p.run_without_inputs()
    assert p.output_block()[0, 31] == 100.0
    graph, p = AudioProcessor.new(0, 1, AudioProcessorOptions(block_size=32))
    with graph.edit() as g:
        for _ in range(100):
            a = g.push(kn.TestNumUGen(2.0))
            v = g.push(kn.TestNumUGen(0.5))
            (a * v).to_graph_out()
    p.run_without_inputs()
    assert p.output_block()[0, 31] == 100.0
    # graph_tests.rs:128-160 multichannel_nodes
    graph, p = AudioProcessor.new(0, 2, AudioProcessorOptions(block_size=16))
    with graph.edit() as g:
        v = [g.push(kn.TestNumUGen(x)) for x in (0.125, 1.0, 0.5, 4.125)]
        m = g.push(kn.MathUGen(2, kn.MathOp.Add))
        (v[0] | v[1] | v[2] | v[3]).to(m).to_graph_out()
    p.run_without_inputs()
    out = p.output_block()
    assert (out[0, 0], out[1, 0]) == (0.625, 5.125)
    # graph_edit.rs:2079-2100: parameters set before the first block
    graph, p = AudioProcessor.new(0, 1, AudioProcessorOptions(block_size=16))
    with graph.edit() as g:
        n1 = g.push(kn.TestInPlusParamUGen())
        g.set(n1, 0, 0.5, kn.Time.asap())
        n2 = g.push(kn.TestInPlusParamUGen())
        g.set(n2, 0, 1.25, kn.Time.asap())
        n3 = g.push(kn.TestInPlusParamUGen())
        g.set(n3, 0, 0.125, kn.Time.asap())
        (n1 >> n2 >> n3).to_graph_out()
    p.run_without_inputs()
    assert p.output_block()[0, 0] == 0.5 + 1.25 + 0.125


def test_precise_timing_golden_vector_through_the_engine():
    # knaster_core_dsp/src/wrappers_core.rs:167-200, driven by absolute-time events
    graph, p = AudioProcessor.new(0, 1, AudioProcessorOptions(block_size=16))
    with graph.edit() as g:
        n = g.push(kn.TestInPlusParamUGen().precise_timing(10))
        n.to_graph_out()
        for d in (5, 6, 8, 9, 10):
            n.param(0).set_at(float(d), kn.Seconds.from_samples(d, SR))
    p.run_without_inputs()
    assert p.output_block()[0].tolist() == [0., 0., 0., 0., 0., 5., 6., 6., 8., 9., 10., 10., 10., 10., 10., 10.]


def test_additive_bank_with_smoothing():
    # configs[1] at reduced size: integer phase + block-rate amplitude ramps => per-voice bit-exact
    def build(graph):
        return banks.additive_bank(graph, 96, 1.0)

    gpu, ref, gt, rt, proc = both(build, 750)
    assert np.array_equal(gt, rt)
    assert np.abs(gpu - ref).max() <= 1e-5
    assert np.abs(ref).max() > 1e-3
    assert proc.info()["n_groups"] == 1


@pytest.mark.parametrize("envelope", ["asr", "segments"])
@pytest.mark.parametrize("force_interp", [False, True])
def test_subtractive_bank_with_note_events(envelope, force_interp):
    # configs[2] at reduced size: saw -> SVF lowpass -> envelope -> VCA, sample-accurate events
    def build(graph):
        return banks.subtractive_bank(graph, 70, 1.0, envelope=envelope)

    gpu, ref, gt, rt, proc = both(build, 750, force_interpreter=force_interp)
    assert np.abs(rt).max() > 1e-4
    assert np.abs(gt - rt).max() <= 1e-4          # IIR budget (sequential evaluation: expect ~0)
    assert np.abs(gpu - ref).max() <= 1e-5
    assert proc.info()["dropped_changes"] == 0


def test_subtractive_segments_variant_runs_the_fused_kernel_bit_identical_to_the_interpreter():
    # configs[2] variant B: Envelope (f64 segments) in place of EnvAsr => recipe "render_sub_seg"
    def run(force_interp):
        opts = AudioProcessorOptions(sample_rate=SR, force_interpreter=force_interp)
        graph, proc = AudioProcessor.new(0, 2, opts)
        ids = banks.subtractive_bank(graph, 70, 1.0, envelope="segments")
        for i in ids:
            proc.add_tap(i, 0)
        out = proc.render(750)
        return out, proc.read_taps(), proc.info()["kernels"]

    fo, ft, fk = run(False)
    io, it, ik = run(True)
    assert fk == ["render_sub_seg"] and ik == ["render_interp"]
    assert np.array_equal(ft, it)              # per-voice: same arithmetic in the same order
    assert np.abs(fo - io).max() <= 1e-6       # bus: the summation trees differ
    opts = AudioProcessorOptions(sample_rate=SR)
    graph, proc = AudioProcessor.new(0, 2, opts)   # and without taps (the TAPS=false instantiation)
    banks.subtractive_bank(graph, 70, 1.0, envelope="segments")
    assert np.array_equal(proc.render(750), fo)


def test_envelope_choreography_in_the_fused_voice_shape():
    # Envelope state machine edge cases inside the render_sub_seg voice shape: looping, a shorter-than-a-sample
    # segment, five segments, time_scale changes (incl. 0 and negative), jump_to_segment while stopped
    # and while running, t_stop in the middle of a segment, two events on one frame, restart while running
    def build(graph):
        ids = []
        with graph.edit() as g:
            for i in range(40):
                segs = [kn.EnvelopeSegment(0.004 + 0.001 * (i % 7), 1.0), kn.EnvelopeSegment(1e-5 if i % 5 == 0 else 0.01, 0.7),
                        kn.EnvelopeSegment(0.02, 0.3 + 0.01 * i), kn.EnvelopeSegment(0.003, 0.9), kn.EnvelopeSegment(0.05, 0.0)]
                envu = kn.Envelope(0.1 if i % 3 == 0 else 0.0, segs[: 2 + i % 4])
                if i % 4 == 1:
                    envu = envu.looping(True)
                if i % 6 == 2:
                    envu = envu.time_scale(1.7)
                saw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 110.0 * (1 + i % 9)).precise_timing(8))
                svf = g.push(kn.SvfFilter(kn.SvfFilterType.Low, 700.0 + 90.0 * i, 2.0, 0.0).precise_timing(8))
                env = g.push(envu.wr_mul(0.05).precise_timing(8))
                sig = (saw >> svf) * env
                sig.out([0, 0]).to_graph_out()
                at = lambda n: kn.Seconds.from_samples(n, SR)
                env.param("t_restart").trig_at(at(100 + 13 * i))
                if i % 2 == 0:
                    env.param("time_scale").set_at([0.5, 0.0, -0.05, 3.0][(i // 2) % 4], at(600 + i))
                    env.param("time_scale").set_at(1.0, at(2500 + i))
                if i % 3 == 1:
                    env.param("t_stop").trig_at(at(900 + 7 * i))
                    env.param("jump_to_segment").set_at(i % 4, at(1500 + i))      # while stopped
                if i % 3 == 2:
                    env.param("jump_to_segment").set_at(7, at(1000 + i))          # clamps to the last segment
                    env.param("jump_to_segment").set_at(0, at(4000))
                    env.param("t_stop").trig_at(at(4000))                         # same frame, after the jump
                env.param("t_restart").trig_at(at(6000 + 11 * i))                # restart while running / stopped
                env.param("wr_mul").set_at(0.02, at(7000 + i))
                ids.append(sig._outputs[0][0])
        return ids

    gpu, ref, gt, rt, proc = both(build, 200)
    kernels = proc.info()["kernels"]           # one group per (segment count, looping) template
    assert len(kernels) > 1 and set(kernels) == {"render_sub_seg"}
    assert np.abs(rt).max() > 1e-3
    assert np.abs(gt - rt).max() <= 1e-4
    assert np.abs(gpu - ref).max() <= 1e-5
    _, _, gi, _, _ = both(build, 200, force_interpreter=True)
    assert np.array_equal(gt, gi)


def test_envasr_choreography_in_the_fused_voice_shape():
    # EnvAsr state machine edge cases inside the render_sub_asr voice shape (the straight-line groups carry an attack and a
    # release through their ends, fused.cu SUB_SAT_ATTACK / SUB_SAT_RELEASE): t_restart while sustaining (one frame of the
    # kept first t >= 1), while attacking and while releasing, t_release during the attack, attack_time / release_time
    # changed in the middle of a ramp, a zero-length attack, a release that is never started, several launches per render
    def build(graph):
        ids = []
        with graph.edit() as g:
            for i in range(48):
                att = [0.0, 0.0005, 0.003, 0.011, 0.03][i % 5]
                rel = [0.002, 0.02, 0.07][i % 3]
                saw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 82.0 * (1 + i % 11)).precise_timing(8))
                svf = g.push(kn.SvfFilter(kn.SvfFilterType.Low, 500.0 + 120.0 * i, 1.5, 0.0).precise_timing(8))
                env = g.push(kn.EnvAsr(att, rel).wr_mul(0.04).precise_timing(8))
                sig = (saw >> svf) * env
                sig.out([0, 0]).to_graph_out()
                at = lambda n: kn.Seconds.from_samples(n, SR)
                env.param("t_restart").trig_at(at(90 + 17 * i))
                if i % 4 == 0:      # restart while sustaining, twice
                    env.param("t_restart").trig_at(at(3000 + i))
                    env.param("t_restart").trig_at(at(3001 + i))
                if i % 4 == 1:      # restart in the middle of the attack, release during the next attack
                    env.param("t_restart").trig_at(at(120 + 17 * i))
                    env.param("t_release").trig_at(at(4000 + i))
                    env.param("t_restart").trig_at(at(4100 + i))
                    env.param("t_release").trig_at(at(4130 + i))
                if i % 4 == 2:      # attack_time changed during the attack, release_time during the release
                    env.param("attack_time").set_at(0.05, at(100 + 17 * i))
                    env.param("attack_time").set_at(0.001, at(400 + 17 * i))
                    env.param("t_release").trig_at(at(5000))
                    env.param("release_time").set_at(0.3, at(5040 + i))
                    env.param("t_restart").trig_at(at(5600 + i))     # restart while releasing
                if i % 4 == 3:      # release at the frame the attack ends, and one frame around it
                    n_att = int(att * SR)
                    env.param("t_release").trig_at(at(90 + 17 * i + n_att + (i % 3) - 1))
                    env.param("t_restart").trig_at(at(6000 + i))     # restart after Stopped, never released again
                env.param("wr_mul").set_at(0.02, at(7000 + i))
                ids.append(sig._outputs[0][0])
        return ids

    gpu, ref, gt, rt, proc = both(build, 160, no_scan=True)
    assert proc.info()["kernels"] == ["render_sub_asr"]
    assert np.abs(rt).max() > 1e-3
    assert np.abs(gt - rt).max() <= 1e-4
    assert np.abs(gpu - ref).max() <= 1e-5
    _, _, gi, _, _ = both(build, 160, force_interpreter=True)
    assert np.array_equal(gt, gi)
    # the small-bank scan kernel renders the chunks that hold these events in reference order
    gs, _, gst, _, ps = both(build, 160)
    assert ps.info()["kernels"][0].startswith("render_sub_scan")
    assert np.abs(gst - rt).max() <= 1e-4 and np.abs(gs - ref).max() <= 1e-5
    # block-by-block calls (every call stores and reloads the envelope state) render the same samples
    opts = AudioProcessorOptions(block_size=64, sample_rate=SR, no_scan=True)
    graph, p2 = AudioProcessor.new(0, 2, opts)
    build(graph)
    blocks = []
    for _ in range(160):
        p2.run_without_inputs()
        blocks.append(p2.output_block())
    assert np.array_equal(np.stack(blocks), gpu)


def test_subtractive_intermediate_nodes_match():
    # tap every node of a voice (forces the interpreter): oscillators <= 1e-5, filter <= 1e-4
    ids = {}

    def build(graph):
        with graph.edit() as g:
            saw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 220.0).precise_timing(8))
            svf = g.push(kn.SvfFilter(kn.SvfFilterType.Low, 900.0, 4.0, 0.0).precise_timing(8))
            env = g.push(kn.EnvAsr(0.01, 0.2).precise_timing(8))
            lp = g.push(kn.OnePoleLpf(2000.0))
            sig = ((saw >> svf >> lp) * env)
            sig.to_graph_out()
            env.param("t_restart").trig_at(kn.Seconds.from_samples(1000, SR))
            saw.param("freq").set_at(13000.0, kn.Seconds.from_samples(9001, SR))  # >= sr/4: sine guard path
            saw.param("freq").set_at(330.0, kn.Seconds.from_samples(12345, SR))
            svf.param("cutoff_freq").set_at(3000.0, kn.Seconds.from_samples(15000, SR))
            svf.param("q").set_at(0.8, kn.Seconds.from_samples(15001, SR))
            lp.param("cutoff_freq").set_at(500.0, kn.Seconds.from_samples(20000, SR))
            env.param("t_release").trig_at(kn.Seconds.from_samples(30000, SR))
        return [saw.id(), svf.id(), env.id(), lp.id(), sig._outputs[0][0]]

    gpu, ref, gt, rt, _ = both(build, 750, outputs=1)
    assert np.array_equal(gt[2], rt[2])                    # envelope: pure f32 recurrence
    assert np.abs(gt[0] - rt[0]).max() <= 1e-5             # oscillator
    assert np.abs(gt[1] - rt[1]).max() <= 1e-4             # SVF
    assert np.abs(gt[3] - rt[3]).max() <= 1e-4             # one-pole
    assert np.abs(gpu - ref).max() <= 1e-4
    assert np.abs(rt[4]).max() > 0.05


@pytest.mark.parametrize("force_interp", [False, True])
def test_fm_bank_audio_rate_routes(force_interp):
    # configs[3] at reduced size: SinNumeric -> (*idx + fc) -> SinNumeric.ar_params() freq
    def build(graph):
        return banks.fm_bank(graph, 40)

    gpu, ref, gt, rt, _ = both(build, 750, force_interpreter=force_interp)
    assert np.abs(gt - rt).max() <= 1e-5
    assert np.abs(gpu - ref).max() <= 1e-5
    assert np.abs(ref).max() > 1e-2


def test_fm_voice_full_scale_over_10_seconds():
    # the hard case for float-phase parity: a carrier integrates its modulator for 480 000 samples
    def build(graph):
        with graph.edit() as g:
            mod = g.push(kn.SinNumeric(311.0))
            car = g.push(kn.SinNumeric(207.0).ar_params())
            car.link("freq", mod * 900.0 + 207.0)
            car.to_graph_out()
        return [car.id(), mod.id()]

    gpu, ref, gt, rt, _ = both(build, 7500, outputs=1)
    assert np.abs(gt[1] - rt[1]).max() <= 1e-5
    assert np.abs(gt[0] - rt[0]).max() <= 1e-5


def test_all_svf_types_and_math_ops():
    def build(graph):
        ids = []
        with graph.edit() as g:
            for ty in kn.SvfFilterType:
                saw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 100.0 + 30 * int(ty)))
                f = g.push(kn.SvfFilter(ty, 800.0 + 100 * int(ty), 1.5, 3.0))
                sig = ((saw >> f) * 0.05 + 0.001 - 0.002) / 2.0
                sig.to_graph_out()
                ids.append(sig._outputs[0][0])
            a = g.push(kn.SinWt(100.0).wr_add(1.5).wr_sub(0.25).wr_v_sub_gen(3.0))
            b = g.push(kn.TestInPlusParamUGen().wr_div(2.0).wr_v_div_gen(0.1).wr_mul(0.5))
            a.to(b).to_graph_out()
            ids.append(b.id())
        return ids

    gpu, ref, gt, rt, proc = both(build, 200, outputs=1)
    assert np.abs(gt - rt).max() <= 1e-4
    assert np.abs(gpu - ref).max() <= 1e-4
    assert proc.info()["n_groups"] == 10   # 9 filter types are 9 different voice shapes + the wrapper chain


def test_all_polyblep_waveforms():
    # polyblep.rs:90-120: all 14 waveforms, three pitches each (one above the sr/4 sine guard), with
    # pulse-width changes, a waveform switch by event and an audio-rate pulse width
    def build(graph):
        ids = []
        with graph.edit() as g:
            for wf in kn.Waveform:
                for k, f in enumerate((55.0, 1234.5, 12500.0)):
                    osc = g.push(kn.PolyBlep(wf, f).precise_timing(4))
                    osc.to_graph_out()
                    osc.param("pulse_width").set_at(0.2 + 0.1 * k, kn.Seconds.from_samples(3000 + 7 * int(wf), SR))
                    osc.param("pulse_width").set_at(0.93, kn.Seconds.from_samples(9000, SR))
                    osc.param("freq").set_at(f * 1.5 if k < 2 else 300.0, kn.Seconds.from_samples(6000 + k, SR))
                    ids.append(osc.id())
            sw = g.push(kn.PolyBlep(kn.Waveform.Sawtooth, 220.0).precise_timing(4))
            sw.to_graph_out()
            for i, wf in enumerate(kn.Waveform):
                sw.param("waveform").set_at(int(wf), kn.Seconds.from_samples(800 * (i + 1) + i, SR))
            ids.append(sw.id())
            lfo = g.push(kn.SinWt(3.0))
            pwm = g.push(kn.PolyBlep(kn.Waveform.Rectangle, 110.0).ar_params())
            pwm.link("pulse_width", lfo * 0.4 + 0.5)
            pwm.to_graph_out()
            ids.append(pwm.id())
        return ids

    gpu, ref, gt, rt, proc = both(build, 200, outputs=1)
    assert np.isfinite(rt).all()
    assert np.abs(gt - rt).max() <= 1e-5
    assert np.abs(gpu - ref).max() <= 1e-4     # 44 voices summed in a different order
    assert (np.abs(rt).max(axis=1) > 0.3).all()


def test_math1_pow_phasor_and_pow_wrappers():
    # math.rs:72-85,172-243 (Pow, Ceil/Sqrt/Floor/Trunc/Fract/Exp), osc.rs:170-213 (Phasor),
    # wrappers_core/math.rs:507-661 (WrPowf, WrPowi).  powf / expf come from libm in knaster and from
    # CUDA's math library here (<= 2 ulp): compared relative to the signal peak.
    def build(graph):
        ids = []
        with graph.edit() as g:
            ph = g.push(kn.Phasor(3.7).precise_timing(2))
            ph.param("freq").set_at(911.3, kn.Seconds.from_samples(5000, SR))
            ids.append(ph.id())
            ph.to_graph_out()
            for op in kn.Math1Op:
                src = g.push(kn.SinWt(97.0 + 11 * int(op)).wr_mul(3.0).wr_add(3.5 if op == kn.Math1Op.Sqrt else 0.0))
                m = g.push(kn.Math1UGen(op))
                src.to(m).to_graph_out()
                ids.append(m.id())
            a = g.push(kn.SinWt(220.0).wr_mul(0.5).wr_add(1.0))      # base in [0.5, 1.5]
            b = g.push(kn.Phasor(2.0).wr_mul(4.0).wr_sub(2.0))        # exponent in [-2, 2]
            p = a.pow(b)
            p.to_graph_out()
            ids.append(p._outputs[0][0])
            w1 = g.push(kn.SinWt(330.0).wr_mul(0.4).wr_add(0.6).wr_powf(2.5))
            w2 = g.push(kn.SinWt(331.0).wr_powi(5))
            w3 = g.push(kn.SinWt(332.0).wr_add(1.5).wr_powi(-3))
            for w in (w1, w2, w3):
                w.to_graph_out()
                ids.append(w.id())
        return ids

    gpu, ref, gt, rt, _ = both(build, 300, outputs=1)
    assert np.isfinite(rt).all() and np.isfinite(gt).all()
    peak = np.maximum(1.0, np.abs(rt).max(axis=1, keepdims=True))
    assert (np.abs(gt - rt) / peak).max() <= 1e-6
    assert np.array_equal(gt[0], rt[0])            # Phasor: f64 phase, bit-identical
    for i in (1, 2, 3, 4, 5):                      # Ceil, Sqrt, Floor, Trunc, Fract: IEEE-exact
        assert np.array_equal(gt[i], rt[i])
    assert np.array_equal(gt[-2], rt[-2]) and np.array_equal(gt[-1], rt[-1])   # powi: multiplications only
    assert np.abs(gpu - ref).max() <= 1e-5 * np.abs(ref).max()


def test_unsupported_graph_fails_loudly_on_gpu_too():
    graph, p = AudioProcessor.new(0, 1, AudioProcessorOptions())
    with graph.edit() as g:
        lfo = g.push(kn.SinWt(3.0))
        e = g.push(kn.EnvAsr(0.01, 0.1).ar_params())
        e.link("t_restart", lfo * 0.001 + 0.01)      # audio-rate route into a trigger: rejected (a type error in knaster too)
        (g.push(kn.SinWt(100.0)) * e).to_graph_out()
    from knaster_b200._ffi import KgpuError

    with pytest.raises(KgpuError):
        p.run_without_inputs()

#!/usr/bin/env python3
"""Regenerate tests/golden/*.  Run from the repo root: python tests/golden/make_golden.py

Two kinds of fixture live here:

* reference_kat.json -- the known-answer vectors the REFERENCE's own tests hold for the hot path
  (SURVEY 8c), transcribed by hand with the file:line they come from.  These pin the oracle.
* oracle_cases.npz -- small renders of every bench configuration and of the widened UGen set made by
  the CPU oracle (oracle/knaster_oracle.cpp) at the commit that added them.  knaster is Rust and cannot
  be built in this image (no cargo/rustc), so these are NOT outputs of the reference: for the DSP bodies
  no reference test pins (SinWt, PolyBlep, SvfFilter, EnvAsr, Envelope, smoothing, noise...) parity is
  unpinned and these vectors only freeze the restatement, so that a later edit of the oracle or of the
  engine cannot drift silently (tests/test_golden_fixtures.py checks both against them).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
SR = 48000

REFERENCE_KAT = {
    "_comment": "known-answer vectors held by the reference's own tests for the hot path (paths relative to /root/reference)",
    "sample_accurate_parameters_test": {
        "source": "knaster_core_dsp/src/wrappers_core.rs:167-200",
        "block_size": 16, "delays": [5, 6, 8, 9, 10], "values": [5.0, 6.0, 8.0, 9.0, 10.0],
        "expected": [0, 0, 0, 0, 0, 5, 6, 6, 8, 9, 10, 10, 10, 10, 10, 10]},
    "sample_accurate_parameters_with_wrappers_test": {
        "source": "knaster_core_dsp/src/wrappers_core.rs:202-250", "tolerance": 2e-4,
        "expected": [0, 0, 0, 0, 0, 5, 6, 6, 8, 9, 10, 10, 10, 10, 10, 10]},
    "wrapper_arithmetic": {
        "source": "knaster_core_dsp/src/wrappers_core.rs:124-164", "input": 2.0,
        "cases": {"wr_add(3)": 5.0, "wr_mul(3)": 6.0, "wr_div(4)": 0.5, "wr_v_div(4)": 2.0, "wr_sub(3)": -1.0,
                  "wr_v_sub(3)": 1.0, "wr_powf(3)": 8.0, "wr_powi(3)": 8.0}},
    "gen_arithmetics": {
        "source": "knaster_core_dsp/src/ugens/math.rs:317-352", "a": 3.0, "b": 2.0,
        "cases": {"Add": 5.0, "Sub": 1.0, "Mul": 6.0, "Div": 1.5}},
    "gen_arithmetics_multichannel": {
        "source": "knaster_core_dsp/src/ugens/math.rs:354-389", "layout": "[a0, a1, b0, b1]"},
    "time_sample_conversion": {
        "source": "knaster_primitives/src/time.rs:461-503", "subsample_tesimals_per_second": 282240000},
    "wrappers_vs_nodes_100": {
        "source": "knaster_benchmarks/benches/wrappers_vs_nodes.rs:75-113", "expected_out_31": 100.0},
}


def render(build, n_blocks, outputs=2, block=64):
    from knaster_b200.graph import Graph
    from oracle.oracle import OracleProcessor

    g = Graph(0, outputs, block, SR)
    ids = build(g)
    orc = OracleProcessor(g, ring_buffer_size=1 << 20)
    for i in ids:
        orc.add_tap(i, 0)
    out, taps = orc.render(n_blocks)
    return out, (taps if ids else np.zeros((0, n_blocks * block), np.float32))


def cases():
    """name -> (build(graph) -> tap ids, n_blocks).  Shared with tests/test_golden_fixtures.py."""
    import knaster_b200 as kn
    from knaster_b200 import banks

    def noise(g):
        kn.reset_randomness_seed(0)
        ids = []
        with g.edit() as e:
            for u in (kn.WhiteNoise(), kn.PinkNoise(), kn.BrownNoise(), kn.RandomLin(700.0)):
                n = e.push(u)
                n.out([0, 0]).to_graph_out()
                ids.append(n.id())
        return ids

    def waveforms(g):
        ids = []
        with g.edit() as e:
            for wf in kn.Waveform:
                n = e.push(kn.PolyBlep(wf, 331.0 + 17.0 * int(wf)))
                n.out([0, 0]).to_graph_out()
                ids.append(n.id())
        return ids

    def many_sines(g):  # knaster/examples/many_sines.rs:51-60: (EnvAr * SinWt.wr_mul) >> Pan2 -> stereo out
        ids = []
        with g.edit() as e:
            for i in range(6):
                env = e.push(kn.EnvAr(0.01, 0.05))
                sine = e.push(kn.SinWt(400.0 + 111.0 * i).wr_mul(0.05))
                pan = e.push(kn.Pan2(-1.0 + 0.4 * i))
                ((env * sine) >> pan).to_graph_out()
                env.param("t_restart").trig_at(kn.Seconds.from_samples(50 + 30 * i, SR))
                ids.append(pan.id())
        return ids

    return {
        "many_sines_pan2": (many_sines, 40),
        "readme_sine": (banks.readme_sine, 8),
        "additive_4": (lambda g: banks.additive_bank(g, 4, 0.05), 37),
        "subtractive_asr_3": (lambda g: banks.subtractive_bank(g, 3, 0.15, n_notes=2), 112),
        "subtractive_segments_3": (lambda g: banks.subtractive_bank(g, 3, 0.15, n_notes=2, envelope="segments"), 112),
        "fm_3": (lambda g: banks.fm_bank(g, 3), 30),
        "noise_4": (noise, 20),
        "polyblep_waveforms": (waveforms, 12),
    }


def main():
    with open(os.path.join(HERE, "reference_kat.json"), "w") as f:
        json.dump(REFERENCE_KAT, f, indent=1)
    arrays = {}
    for name, (build, n_blocks) in cases().items():
        out, taps = render(build, n_blocks)
        arrays[name + "/bus"] = out
        arrays[name + "/taps"] = taps
        print(f"{name}: bus {out.shape} peak {np.abs(out).max():.4f}, taps {taps.shape}")
    np.savez_compressed(os.path.join(HERE, "oracle_cases.npz"), **arrays)
    print("wrote", os.path.getsize(os.path.join(HERE, "oracle_cases.npz")), "bytes")


if __name__ == "__main__":
    main()

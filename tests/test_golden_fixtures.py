"""The committed fixtures of tests/golden/ (see make_golden.py for what they are and are not).

CPU: the oracle must still reproduce oracle_cases.npz bit for bit (a changed oracle has to be a
deliberate, reviewed regeneration), and reference_kat.json must agree with the vectors
test_oracle_golden.py checks.  GPU: the CUDA engine against the same fixtures, WITHOUT the oracle in the
loop -- the tolerances are the north star's (oscillators 1e-5, IIR 1e-4, normalised bus 1e-5)."""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SR = 48000

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
FIX = np.load(os.path.join(HERE, "oracle_cases.npz"))
CASES = make_golden.cases()


def test_reference_kat_file_matches_the_vectors_the_oracle_tests_use():
    import test_oracle_golden as T

    kat = json.load(open(os.path.join(HERE, "reference_kat.json")))
    assert kat["sample_accurate_parameters_test"]["expected"] == T.GOLDEN_PRECISE
    assert kat["sample_accurate_parameters_with_wrappers_test"]["expected"] == T.GOLDEN_PRECISE
    for entry in kat.values():
        if isinstance(entry, dict):
            assert ".rs:" in entry["source"]        # every vector cites the reference test it comes from


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_committed_cases(name):
    build, n_blocks = CASES[name]
    out, taps = make_golden.render(build, n_blocks)
    assert np.array_equal(out, FIX[name + "/bus"]), "oracle output changed: regenerate tests/golden deliberately"
    assert np.array_equal(taps, FIX[name + "/taps"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_engine_matches_the_committed_cases(name):
    from knaster_b200.processor import AudioProcessor, AudioProcessorOptions

    build, n_blocks = CASES[name]
    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(sample_rate=SR))
    ids = build(graph)
    for i in ids:
        proc.add_tap(i, 0)
    out = proc.render(n_blocks)
    taps = proc.read_taps()
    ref_out, ref_taps = FIX[name + "/bus"], FIX[name + "/taps"]
    exact = name in ("readme_sine", "additive_4", "noise_4", "many_sines_pan2")      # integer phase / integer RNG + single f32 ops
    if exact:
        assert np.array_equal(taps, ref_taps)
    else:
        assert np.abs(taps - ref_taps).max() <= (1e-4 if name.startswith("subtractive") else 1e-5)
    scale = max(1.0, float(np.abs(ref_out).max()))
    assert np.abs(out - ref_out).max() <= 1e-5 * scale
    assert np.abs(ref_out).max() > 1e-2

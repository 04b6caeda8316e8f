"""One small render per kernel of the engine, for compute-sanitizer (memcheck / racecheck / initcheck):
render_sub_asr, render_sub_asr2, render_sub_seg, render_sub_scan, render_fm2 (+ _wide), render_add_wt, render_interp, render_jit,
reduce_bus.  usage: compute-sanitizer --tool racecheck python tools/debug/sanitize.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from knaster_b200 import banks
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions

N_BLOCKS = 24
seen = []
for name, build, opts in [
    ("sub_asr", lambda g: banks.subtractive_bank(g, 70, 0.05, n_notes=3), dict(no_scan=True)),
    ("sub_scan", lambda g: banks.subtractive_bank(g, 5, 0.05, n_notes=3), {}),
    ("sub_seg", lambda g: banks.subtractive_bank(g, 70, 0.05, n_notes=3, envelope="segments"), {}),
    ("fm2", lambda g: banks.fm_bank(g, 40), {}),
    ("add_wt", lambda g: banks.additive_bank(g, 200, 0.05), {}),
    ("interp", lambda g: banks.chain_bank(g, 40, 0.05, n_notes=3), dict(force_interpreter=True)),
    ("jit", lambda g: banks.chain_bank(g, 40, 0.05, n_notes=3), dict(force_jit=True)),
]:
    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(**opts))
    ids = build(graph)
    proc.add_tap(ids[0], 0)
    proc.set_blocks_per_launch(16)
    out = proc.render(N_BLOCKS)
    taps = proc.read_taps()
    assert np.isfinite(out).all() and np.isfinite(taps).all()
    seen += proc.info()["kernels"]
    print(name, proc.info()["kernels"], float(np.abs(out).max()))
print("SANITIZE_RENDERS_OK", sorted(set(seen)))

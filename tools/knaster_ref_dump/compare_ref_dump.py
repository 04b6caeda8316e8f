#!/usr/bin/env python3
"""Compare a `knaster_ref_dump` output (real knaster) with the oracle's -- and, when a GPU is present,
the engine's -- render of the same manifest.  A pass turns "parity unpinned" into "pinned" for the
configuration: commit the .f32 as a fixture under tests/golden/ together with its manifest.

    python tools/knaster_ref_dump/compare_ref_dump.py m.txt ref.f32 [--gpu]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import export_manifest as X  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("manifest")
    ap.add_argument("ref")
    ap.add_argument("--gpu", action="store_true")
    a = ap.parse_args()
    head = dict(l.split(None, 1) for l in open(a.manifest).read().splitlines()[1:3])
    config = head["config"].strip()
    t = head["sr"].split()
    n_blocks, voices = int(t[4]), int(t[6])
    seconds = n_blocks * X.BLOCK / X.SR
    ref = np.fromfile(a.ref, dtype="<f4").reshape(n_blocks, 2, X.BLOCK)
    from oracle.oracle import OracleProcessor

    g, ev = X.build(config, voices, seconds)
    g.pending_event_arrays = [ev.copy()] if len(ev) else []
    orc, _ = OracleProcessor(g, ring_buffer_size=1 << 22).render(n_blocks)
    print(f"oracle vs knaster: max abs diff {np.abs(orc - ref).max():.3e} (peak {np.abs(ref).max():.4f}), "
          f"bit-identical: {np.array_equal(orc, ref)}")
    if a.gpu:
        from knaster_b200.processor import AudioProcessor, AudioProcessorOptions

        graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions())
        _, ev2 = X.build(config, voices, seconds, graph)   # same builder, into the engine's graph
        graph.pending_event_arrays = [ev2.copy()] if len(ev2) else []
        gpu = proc.render(n_blocks)
        print(f"engine vs knaster: max abs diff {np.abs(gpu - ref).max():.3e}")


if __name__ == "__main__":
    main()

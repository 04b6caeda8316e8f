#!/usr/bin/env python3
"""The voice shape of knaster/examples/many_sines.rs:51-60 -- 600 x ((EnvAr * SinWt.wr_mul) >> Pan2) -- with
every envelope re-triggered at random times, rendered in 1 s pieces and streamed to a WAVE file, plus the first
four voices as stems.

    python examples/many_sines.py out.wav
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knaster_b200 as kn
from knaster_b200 import sinks
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions

SR, BLOCK = 48000, 64


def main(path="many_sines.wav", seconds=4):
    rng = np.random.default_rng(1)
    graph, processor = AudioProcessor.new(0, 2, AudioProcessorOptions(block_size=BLOCK, sample_rate=SR))
    restarts, pans = [], []
    with graph.edit() as g:
        for _ in range(600):
            env = g.push(kn.EnvAr(0.01, 0.1))
            sine = g.push(kn.SinWt(float(rng.uniform(3000.0, 10000.0))).wr_mul(float(rng.uniform(0.01, 0.015))))
            pan = g.push(kn.Pan2(float(rng.uniform(-1.0, 1.0))))
            ((env * sine) >> pan).to_graph_out()
            restarts.append(env.param("t_restart"))
            pans.append(pan)
    for i in range(4):
        processor.add_tap(pans[i].id(), 0)
    for p in restarts:                               # schedule ahead, sample-accurately, like Parameter::trig_at
        for t in np.sort(rng.uniform(0.0, seconds, 6)):
            p.trig_at(kn.Seconds.from_samples(int(t * SR), SR))
    stems = []
    with sinks.WavStream(path, 2, SR) as wav:
        for _ in range(int(seconds)):
            wav.write(processor.render(SR // BLOCK))
            stems.append(processor.read_taps())
    sinks.save_stems(np.concatenate(stems, axis=1), os.path.splitext(path)[0] + "_voice{:02d}.wav", SR)
    print(f"{path}: {seconds} s of 600 voices, kernels {sorted(set(processor.info()['kernels']))}")


if __name__ == "__main__":
    main(*sys.argv[1:2])

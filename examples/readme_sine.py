#!/usr/bin/env python3
"""knaster's README example (README.md:35-47 of the reference) through the B200 engine: a 440 Hz SinWt times
0.2 to both channels, rendered non-realtime and written as a 16-bit WAVE file.

    python examples/readme_sine.py out.wav            # needs a CUDA device: there is no CPU fallback
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import knaster_b200 as kn
from knaster_b200 import sinks
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions


def main(path="readme_sine.wav", seconds=2.0):
    graph, processor = AudioProcessor.new(0, 2, AudioProcessorOptions(block_size=64, sample_rate=48000))
    with graph.edit() as g:
        sine = g.push(kn.SinWt(440.0))          # let sine = g.push(SinWt::new(440.0));
        sig = sine * 0.2                        # let sig = sine * 0.2;
        sig.out([0, 0]).to_graph_out()          # sig.out([0, 0]).to_graph_out();
    audio = processor.render(int(seconds * 48000) // 64)   # [n_blocks][2][64] f32: the batched run_without_inputs() loop
    sinks.save_to_disk(audio, path, 48000)
    print(f"{path}: {audio.shape[0] * 64} frames, peak {abs(audio).max():.3f}, kernel {processor.info()['kernels']}")


if __name__ == "__main__":
    main(*sys.argv[1:2])

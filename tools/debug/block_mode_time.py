"""How fast is the reference's own calling pattern -- AudioProcessor::run_without_inputs() once per block,
then output_block() -- through the engine?  (bench.py measures the batched kgpu_render instead.)"""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from knaster_b200 import banks
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions
for voices in (16384, 256):
    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions())
    banks.subtractive_bank(graph, voices, 2.0)
    for _ in range(20): proc.run_without_inputs()
    t0 = time.perf_counter()
    n = 750
    for _ in range(n):
        proc.run_without_inputs()
        blk = proc.output_block()
    dt = time.perf_counter() - t0
    print(f"{voices} voices: {1e6*dt/n:.1f} us per block (a block is 1333 us of audio) = {voices*n*64/dt:.3e} voice-samples/s, peak {float(np.abs(blk).max()):.4f}")

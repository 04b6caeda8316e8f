import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from knaster_b200 import banks
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions
for force in (False, True):
    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(force_interpreter=force))
    banks.readme_sine(graph)
    proc.render(64)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); out = proc.render(7500); ts.append(time.perf_counter() - t0)
    print("README example, 10 s:", proc.info()["kernels"], "%.2f ms per render" % (1e3 * min(ts)), "peak", float(np.abs(out).max()))

import sys, numpy as np
sys.path.insert(0, '/root/repo')
import knaster_b200 as kn
from knaster_b200 import banks
from knaster_b200.processor import AudioProcessor, AudioProcessorOptions
SR = 48000

def run(force, nv=4, nb=750, notes=8):
    graph, proc = AudioProcessor.new(0, 2, AudioProcessorOptions(sample_rate=SR, force_interpreter=force))
    ids = banks.subtractive_bank(graph, nv, 1.0, envelope="segments", n_notes=notes)
    for i in ids: proc.add_tap(i, 0)
    out = proc.render(nb)
    return out, proc.read_taps(), proc.info()["kernels"]

fo, ft, fk = run(False)
io, it, ik = run(True)
print(fk, ik)
print("fused taps max", np.abs(ft).max(axis=1), "bus", np.abs(fo).max())
print("interp taps max", np.abs(it).max(axis=1), "bus", np.abs(io).max())
for v in range(ft.shape[0]):
    d = np.nonzero(ft[v] != it[v])[0]
    print(v, "first diff", d[:3], "n", len(d), "first nonzero interp", np.nonzero(it[v])[0][:1], "fused", np.nonzero(ft[v])[0][:1])
    if len(d):
        k = d[0]; print("   ", ft[v][k:k+4], it[v][k:k+4])

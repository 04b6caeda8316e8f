import sys, re, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import knaster_b200 as kn
import test_gpu_parity as T
src = open('/root/repo/tests/test_gpu_parity.py').read()
m = re.search(r"def test_envelope_choreography_in_the_fused_voice_shape\(\):.*?\n    def build\(graph\):\n(.*?)\n        return ids\n", src, re.S)
code = "def build(graph):\n" + m.group(1) + "\n        return ids\n"
SR = 48000
exec(code)
gpu, ref, gt, rt, proc = T.both(build, 200)
_, _, gi, _, _ = T.both(build, 200, force_interpreter=True)
print(proc.info()["kernels"])
for v in range(gt.shape[0]):
    df = np.abs(gt[v] - rt[v]); di = np.abs(gi[v] - rt[v]); dfi = np.nonzero(gt[v] != gi[v])[0]
    if df.max() > 0 or di.max() > 0 or len(dfi):
        k = int(np.nonzero(df)[0][0]) if df.max() > 0 else -1
        print(f"voice {v}: fused-oracle max {df.max():.3e} first {k}; interp-oracle max {di.max():.3e}; fused!=interp at {dfi[:3]} n={len(dfi)}")
        if k >= 0: print("    fused", gt[v][k:k+3], "oracle", rt[v][k:k+3], "interp", gi[v][k:k+3])
print("bus", np.abs(gpu - ref).max())

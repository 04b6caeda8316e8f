"""One-off: run tests/test_gpu_fuzz.py's random-voice parity check over many more seeds than the suite does."""
import sys, traceback
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import test_gpu_fuzz as T
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = []
for seed in range(lo, hi):
    try:
        T.test_random_voice_shapes_match_the_oracle(seed)
    except Exception as e:
        bad.append(seed)
        print("seed", seed, "FAILED:", str(e).splitlines()[0][:200])
print(f"{hi - lo - len(bad)} of {hi - lo} seeds passed; failing seeds: {bad}")

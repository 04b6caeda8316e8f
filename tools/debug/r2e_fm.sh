#!/bin/bash
# Round-2: render_fm2 with the lean sine -- parity tests, then interleaved A/B of FM_SUB 16 (_build) and 8 (_build_F8)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2m; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
#python -m pytest tests -m gpu -x -q -k "fm or FM or fuzz or golden or pan2 or noise" > $O/pytest_fm.log 2>&1; echo "rc=$?" >> $O/pytest_fm.log
#python -m pytest tests/test_gpu_full_size.py -x -q -k "fm" > $O/pytest_full_fm.log 2>&1; echo "rc=$?" >> $O/pytest_full_fm.log
for rep in 1 2; do for v in _build _build_F32 _build_ST; do
  KNASTER_GPU_LIB=knaster_b200/csrc/$v/libknaster_gpu.so python bench.py --workload fm $Q --steps 5 > $O/bench_fm${v}_$rep.json 2>$O/bench_fm${v}_$rep.err
done; done
tail -3 $O/pytest_fm.log $O/pytest_full_fm.log; cat $O/bench_fm*.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"

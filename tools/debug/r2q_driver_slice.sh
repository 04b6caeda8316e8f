#!/bin/bash
# Round-2: the calling thread's own slice of the control simulation (HostPlan::driver_slice_through), A/B at 3 and 7 workers per GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2q; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
line() { python -c "
import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',round(d['ms_per_step'],3),round(d['e2e']['ms_per_step'],3))"; }
for rep in 1 2; do
  for ht in 3 7; do for ds in 0 1; do
    KGPU_DRIVER_SLICE=$ds timeout 200 python bench.py $Q --steps 10 --host-threads $ht > $O/ht${ht}_ds${ds}_$rep.json 2>/dev/null; line $O/ht${ht}_ds${ds}_$rep.json "threads=$ht slice=$ds rep=$rep"
  done; done
  timeout 200 python bench.py $Q --steps 10 > $O/default_$rep.json 2>/dev/null; line $O/default_$rep.json "default rep=$rep"
done
KGPU_THREADS=3 timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3

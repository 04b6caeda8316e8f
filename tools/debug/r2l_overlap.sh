#!/bin/bash
# Round-2: mix-bus reduction on its own stream beside the next launch's rendering -- tests, A/B against KGPU_NO_REDUCE_OVERLAP=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2l; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
for rep in 1 2; do for v in 1 0; do
  if [ $v = 1 ]; then export KGPU_NO_REDUCE_OVERLAP=1; else unset KGPU_NO_REDUCE_OVERLAP; fi
  timeout 200 python bench.py $Q --steps 10 > $O/ab_no$v_$rep.json 2>/dev/null
  python -c "
import json;d=json.loads(open('$O/ab_no$v_$rep.json').read().strip().splitlines()[-1]);print('no_overlap=$v',$rep,round(d['ms_per_step'],3),round(d['e2e']['ms_per_step'],3),round(d['roofline']['frac'],4),d['roofline']['avg_launch_ms'],d['roofline']['kernel_share_of_step'])"
done; done
unset KGPU_NO_REDUCE_OVERLAP
KGPU_TIMING=1 timeout 300 python bench.py $Q --steps 2 > $O/bench_timing.json 2> $O/bench_timing.err; grep "device span\|host loop" $O/bench_timing.err | tail -4
for w in subtractive_seg additive fm chain; do timeout 200 python bench.py $Q --workload $w --steps 5 > $O/w_$w.json 2>/dev/null; python -c "
import json;d=json.loads(open('$O/w_$w.json').read().strip().splitlines()[-1]);print('$w',round(d['ms_per_step'],3),round(d['e2e']['ms_per_step'],3))"; done

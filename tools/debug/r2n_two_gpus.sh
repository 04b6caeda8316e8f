#!/bin/bash
# Round-2: the 2-GPU path as the driver runs it -- multi-GPU tests, default bench line (peer bus, parity, bus_check), the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2n; mkdir -p $O
( time timeout 600 python -m pytest tests/test_multi_gpu_gpu.py -x -q -s ) > $O/multi_gpu_test.log 2>&1; echo "rc=$?" >> $O/multi_gpu_test.log
tail -5 $O/multi_gpu_test.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 5 --warmup 3 ) > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
tail -6 $O/bench_2gpu.err
python -c "
import json
d=json.loads([l for l in open('$O/bench_2gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','bus','bus_check','parity','host') if k in d})"

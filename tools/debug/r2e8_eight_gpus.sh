#!/bin/bash
# Round-2: 8 GPUs as the driver runs them (default worker count: 3 per GPU on a 32-core box, the calling thread's slice active)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2e8; mkdir -p $O
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus 8 --steps 10 --warmup 3 --no-other-workloads --no-cpu-baseline ) > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "rc=$?" >> $O/bench_8gpu.err
tail -3 $O/bench_8gpu.err
python -c "
import json
d=json.loads([l for l in open('$O/bench_8gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','bus_check','host') if k in d}); print(d.get('parity'))"

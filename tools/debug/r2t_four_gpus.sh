#!/bin/bash
# Round-2: 4 GPUs as the driver runs them, with 3 event-pipeline workers per GPU (one GPU's share of a 32-core box with 8 GPUs:
# the calling thread's own simulation slice is active) -- default bench line with parity and bus_check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2t; mkdir -p $O
nproc > $O/nproc.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 4 --steps 10 --warmup 3 --host-threads 3 --no-other-workloads --no-cpu-baseline ) > $O/bench_4gpu_ht3.json 2> $O/bench_4gpu_ht3.err; echo "rc=$?" >> $O/bench_4gpu_ht3.err
tail -4 $O/bench_4gpu_ht3.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29575 bench.py --gpus 4 --steps 10 --warmup 3 --no-parity --no-other-workloads --no-cpu-baseline ) > $O/bench_4gpu.json 2> $O/bench_4gpu.err; echo "rc=$?" >> $O/bench_4gpu.err
for f in bench_4gpu_ht3 bench_4gpu; do python -c "
import json
d=json.loads([l for l in open('$O/$f.json').read().strip().splitlines() if l.startswith('{')][-1])
print('$f', {k:d[k] for k in ('value','ms_per_step','e2e','bus','bus_check','parity','host') if k in d})"; done

#!/bin/bash
# Round-2: 4 GPUs as the driver runs them, with 3 event-pipeline workers per GPU (one GPU's share of a 32-core box with 8 GPUs:
# the calling thread's own simulation slice is active) -- bench line with parity and bus_check, then the slice weight A/B and 7 workers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2t; mkdir -p $O
nproc > $O/nproc.txt
run() { # name, env..., -- args
  local name=$1; shift
  ( time timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 4 --steps 10 --warmup 3 --no-other-workloads --no-cpu-baseline $ARGS ) > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err
  python -c "
import json
d=json.loads([l for l in open('$O/$name.json').read().strip().splitlines() if l.startswith('{')][-1])
print('$name', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d.get('bus_check',{}).get('max_abs_peer_minus_nccl'), d.get('parity',{}).get('taps_bit_identical'), d['host'])"
}
ARGS="--host-threads 3" run bench_4gpu_ht3 KGPU_DRIVER_WEIGHT=50
ARGS="--host-threads 3 --no-parity" run bench_4gpu_ht3_w100 KGPU_DRIVER_WEIGHT=100
ARGS="--host-threads 3 --no-parity" run bench_4gpu_ht3_w50 KGPU_DRIVER_WEIGHT=50
ARGS="--host-threads 3 --no-parity" run bench_4gpu_ht3_noslice KGPU_DRIVER_SLICE=0
ARGS="--no-parity" run bench_4gpu KGPU_DRIVER_WEIGHT=50

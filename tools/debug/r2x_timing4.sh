#!/bin/bash
# Round-2: where a 4-GPU end-to-end step spends its host time (KGPU_TIMING log of every rank, 3 workers per GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2x; mkdir -p $O
KGPU_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 4 --steps 3 --warmup 3 --no-other-workloads --no-cpu-baseline --no-parity --host-threads 3 > $O/t.json 2> $O/t.err
grep -c timing $O/t.err; tail -150 $O/t.err | grep "push 5\|stream_begin\|launch [0-9]*:\|worker 0 finished\|render" | tail -60

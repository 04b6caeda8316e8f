#!/bin/bash
# Round-2: compute-sanitizer over one small render per kernel (1 GPU) and over the peer-memory mix bus (2 GPUs).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2s; mkdir -p $O
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/debug/sanitize.py > $O/sanitize_$tool.log 2>&1; echo "rc=$?" >> $O/sanitize_$tool.log
done
KGPU_SUB_TWO_WARPS=1 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/debug/sanitize.py > $O/sanitize_racecheck_two_warps.log 2>&1; echo "rc=$?" >> $O/sanitize_racecheck_two_warps.log
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/helpers/peer_bus_check.py > $O/peer_bus_$tool.log 2>&1; echo "rc=$?" >> $O/peer_bus_$tool.log
  done
  python -m pytest tests/test_multi_gpu_gpu.py -x -q -s > $O/multi_gpu_test.log 2>&1; echo "rc=$?" >> $O/multi_gpu_test.log
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
fi
tail -3 $O/*.log

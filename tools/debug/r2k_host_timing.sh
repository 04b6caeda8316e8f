#!/bin/bash
# Round-2: phase timing of the end-to-end call (15 and 3 event-pipeline threads) + one compute-sanitizer attempt
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2k; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
for t in 15 3; do KGPU_TIMING=1 timeout 300 python bench.py $Q --steps 3 --host-threads $t > $O/bench_ht$t.json 2> $O/bench_ht$t.err; done
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/debug/sanitize.py > $O/sanitize_memcheck.log 2>&1; echo "rc=$?" >> $O/sanitize_memcheck.log
tail -5 $O/sanitize_memcheck.log
grep -c . $O/bench_ht15.err; tail -60 $O/bench_ht15.err

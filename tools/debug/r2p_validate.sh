#!/bin/bash
# Round-2 validation of HEAD on a fresh box: the -m gpu suite, the default bench line and the reference arm, the launch list.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/${TAG:-r2p}; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
SUB="python bench.py --steps 2 --warmup 3 $Q"
timeout 300 $SUB > $O/sub_plain0.json 2> $O/sub_plain0.err && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG:-r2p}_launches_bench_subtractive_10s.csv $SUB > $O/sub_ncu1.log 2>&1
tail -3 $O/pytest_gpu.log; head -c 400 $O/bench_default.json

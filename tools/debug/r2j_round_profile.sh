#!/bin/bash
# Round-2 capture of the kernels the round ends on: GPU test suite, default bench line (+ reference arm), launch list,
# one `--set full` capture each of render_sub_asr, render_fm2, render_sub_scan (256 voices), render_jit (chain), add_wt_render.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2j; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
cap() { # name kernel-regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 "$@" > $O/${name}_plain.json 2> $O/${name}_plain.err && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/r2j_${name}_full "$@" > $O/${name}_ncu.log 2>&1
}
SUB="python bench.py --steps 2 --warmup 3 $Q"
timeout 300 $SUB > $O/sub_plain0.json 2> $O/sub_plain0.err && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2j_launches_bench_subtractive_10s.csv $SUB > $O/sub_ncu1.log 2>&1
cap render_sub_asr render_sub_asr 9 $SUB
cap render_fm2 render_fm2 9 python bench.py --workload fm --steps 2 --warmup 3 $Q
cap render_sub_scan render_sub_scan 9 python bench.py --voices 256 --steps 2 --warmup 3 $Q
cap render_jit render_jit 9 python bench.py --workload chain --steps 2 --warmup 3 $Q
cap add_wt_render add_wt_render 9 python bench.py --workload additive --steps 2 --warmup 3 $Q
ls -la $O
tail -3 $O/pytest_gpu.log; head -c 600 $O/bench_default.json

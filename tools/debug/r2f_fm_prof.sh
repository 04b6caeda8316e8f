#!/bin/bash
# Round-2: FP64 microbenchmark + ncu capture of render_fm2 (lean sine)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2o; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
true
FM="python bench.py --workload fm --steps 2 --warmup 3 $Q"
$FM > $O/fm_plain.json 2> $O/fm_plain.err && ncu --set full --clock-control none --import-source on -k regex:render_fm2 -s 9 -c 1 -f -o $O/r2d_render_fm2_full $FM > $O/fm_ncu.log 2>&1
cat $O/fm_plain.json | head -c 300

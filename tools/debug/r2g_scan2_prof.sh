#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2u; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
SC="env KGPU_SCAN_WARPS=2 timeout 120 python bench.py --voices 256 --steps 2 --warmup 3 $Q"
$SC > $O/scan_plain.json 2> $O/scan_plain.err && timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_sub_scan -s 9 -c 1 -f -o $O/r2d_render_sub_scan_n_full $SC > $O/scan_ncu.log 2>&1
head -c 400 $O/scan_plain.json

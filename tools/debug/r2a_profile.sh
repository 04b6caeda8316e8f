#!/bin/bash
# Round-2 first capture: render_sub_asr with 8-frame groups, render_sub_scan (256 voices), e2e timing at 15 / 3 host threads.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2d; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
python -m pytest tests/test_gpu_scan.py -x -q -s > $O/scan.log 2>&1; echo "rc=$?" >> $O/scan.log
KGPU_SCAN_PIPE=0 python -m pytest tests/test_gpu_scan.py -x -q -s -k faster > $O/scan_nopipe.log 2>&1
for t in 15 3; do KGPU_TIMING=1 python bench.py $Q --steps 3 --host-threads $t > $O/bench_ht$t.json 2> $O/bench_ht$t.err; done
SUB="python bench.py --steps 2 --warmup 3 $Q"
$SUB > $O/sub_plain.json 2> $O/sub_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches_bench_subtractive_10s.csv $SUB > $O/sub_ncu1.log 2>&1
$SUB > $O/sub_plain2.json 2> $O/sub_plain2.err && ncu --set full --clock-control none --import-source on -k regex:render_sub_asr -s 9 -c 1 -f -o $O/r2a_render_sub_asr_full $SUB > $O/sub_ncu2.log 2>&1
SC="python bench.py --voices 256 --steps 2 --warmup 3 $Q"
$SC > $O/scan_plain.json 2> $O/scan_plain.err && ncu --set full --clock-control none --import-source on -k regex:render_sub_scan -s 9 -c 1 -f -o $O/r2a_render_sub_scan_full $SC > $O/scan_ncu2.log 2>&1
for v in 64 256 1024 2048 4096; do python bench.py --voices $v --steps 3 $Q > $O/scan_v$v.json 2>/dev/null; KGPU_SUB_SCAN=0 python bench.py --voices $v --steps 3 $Q > $O/lane_v$v.json 2>/dev/null; done
ls -la $O

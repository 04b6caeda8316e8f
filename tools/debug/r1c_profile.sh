#!/bin/bash
# Round-1 third capture: tests, bench lines, then ncu launch lists + full captures of the three fused
# bank kernels that had no committed capture yet (render_sub_seg, render_fm2, add_wt_render).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --no-cpu-baseline > $O/bench_s3_b.json 2> $O/bench_s3_b.err
python bench.py --workload subtractive_seg --no-cpu-baseline > $O/bench_seg_b.json 2> $O/bench_seg_b.err
python bench.py --workload fm --voices 8192 --no-cpu-baseline > $O/bench_fm_w.json 2> $O/bench_fm_w.err
python bench.py --workload additive --voices 4096 --no-cpu-baseline > $O/bench_add_w.json 2> $O/bench_add_w.err
SEG="python bench.py --workload subtractive_seg --steps 2 --warmup 3 --no-cpu-baseline"
FM="python bench.py --workload fm --voices 8192 --steps 2 --warmup 3 --no-cpu-baseline"
ADD="python bench.py --workload additive --voices 4096 --steps 2 --warmup 3 --no-cpu-baseline"
$SEG > $O/r1c_seg_plain.json 2> $O/r1c_seg_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1c_seg_launches.csv $SEG > $O/r1c_seg_ncu1.log 2>&1
$SEG > $O/r1c_seg_plain2.json 2> $O/r1c_seg_plain2.err && ncu --set full --clock-control none --import-source on -k regex:render_sub_seg -s 6 -c 1 -f -o $O/r1c_render_sub_seg_full $SEG > $O/r1c_seg_ncu2.log 2>&1
$FM > $O/r1c_fm_plain.json 2> $O/r1c_fm_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1c_fm_launches.csv $FM > $O/r1c_fm_ncu1.log 2>&1
$FM > $O/r1c_fm_plain2.json 2> $O/r1c_fm_plain2.err && ncu --set full --clock-control none --import-source on -k regex:render_fm2 -s 6 -c 1 -f -o $O/r1c_render_fm2_full $FM > $O/r1c_fm_ncu2.log 2>&1
$ADD > $O/r1c_add_plain.json 2> $O/r1c_add_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r1c_add_launches.csv $ADD > $O/r1c_add_ncu1.log 2>&1
$ADD > $O/r1c_add_plain2.json 2> $O/r1c_add_plain2.err && ncu --set full --clock-control none --import-source on -k regex:add_wt_render -s 6 -c 1 -f -o $O/r1c_add_wt_render_full $ADD > $O/r1c_add_ncu2.log 2>&1
ls -la $O | grep r1c

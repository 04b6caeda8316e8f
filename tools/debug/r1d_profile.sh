#!/bin/bash
# Round-1 fourth capture: the kernels the round ends on (render_sub_asr v9, two-lane render_fm2).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
SUB="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
FM="python bench.py --workload fm --voices 8192 --steps 2 --warmup 3 --no-cpu-baseline"
$SUB > $O/r1d_sub_plain.json 2> $O/r1d_sub_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1d_sub_launches.csv $SUB > $O/r1d_sub_ncu1.log 2>&1
$SUB > $O/r1d_sub_plain2.json 2> $O/r1d_sub_plain2.err && ncu --set full --clock-control none --import-source on -k regex:render_sub_asr -s 6 -c 1 -f -o $O/r1d_render_sub_asr_full $SUB > $O/r1d_sub_ncu2.log 2>&1
$FM > $O/r1d_fm_plain.json 2> $O/r1d_fm_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1d_fm_launches.csv $FM > $O/r1d_fm_ncu1.log 2>&1
$FM > $O/r1d_fm_plain2.json 2> $O/r1d_fm_plain2.err && ncu --set full --clock-control none --import-source on -k regex:render_fm2 -s 6 -c 1 -f -o $O/r1d_render_fm2_full $FM > $O/r1d_fm_ncu2.log 2>&1
python bench.py > $O/bench_r1d_default.json 2> $O/bench_r1d_default.err
python bench.py --impl reference > $O/bench_r1d_ref.json 2> $O/bench_r1d_ref.err
ls -la $O | grep r1d

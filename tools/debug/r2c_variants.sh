#!/bin/bash
# Round-2: repeated, interleaved A/B of render_sub_asr variants (A: no rotation / predicated event load, B: unconditional
# event load, C: rotation, _build = both), scan-kernel tests + ncu capture of the speculative pre-pass version.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2f; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
python -m pytest tests/test_gpu_scan.py -q -s > $O/scan.log 2>&1; echo "rc=$?" >> $O/scan.log
for rep in 1 2 3; do for v in _build_A _build_B _build_C _build; do
  KNASTER_GPU_LIB=knaster_b200/csrc/$v/libknaster_gpu.so python bench.py $Q --steps 10 > $O/bench${v}_$rep.json 2>/dev/null
done; done
SC="python bench.py --voices 256 --steps 2 --warmup 3 $Q"
$SC > $O/scan_plain.json 2> $O/scan_plain.err && ncu --set full --clock-control none --import-source on -k regex:render_sub_scan -s 9 -c 1 -f -o $O/r2b_render_sub_scan_full $SC > $O/scan_ncu2.log 2>&1
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
ls $O

#!/bin/bash
# Round-2: A/B of the render_sub_asr main-loop changes (rotation, compact limit path) + the speculative scan pre-pass.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2e; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
python -m pytest tests/test_gpu_scan.py tests/test_gpu_properties.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -s > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
for v in _build _build_r0c0 _build_r1c0 _build_r0c1; do
  KNASTER_GPU_LIB=knaster_b200/csrc/$v/libknaster_gpu.so python bench.py $Q --steps 5 > $O/bench$v.json 2>/dev/null
done
KNASTER_GPU_LIB=knaster_b200/csrc/_build/libknaster_gpu.so python bench.py $Q --steps 5 --workload subtractive_seg > $O/bench_seg.json 2>/dev/null
for v in 64 256 1024 2048 4096; do python bench.py --voices $v --steps 3 $Q > $O/scan_v$v.json 2>/dev/null; done
for t in 15 3; do KGPU_TIMING=1 python bench.py $Q --steps 3 --host-threads $t > $O/bench_ht$t.json 2> $O/bench_ht$t.err; done
ls $O

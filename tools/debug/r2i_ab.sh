#!/bin/bash
# Round-2: interleaved A/B of two builds on the headline bench (variants: $VARIANTS, default "_build_A _build")
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2w; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
V=${VARIANTS:-"_build_A _build"}
W=${WORKLOAD:-subtractive}
for rep in 1 2 3; do for v in $V; do
  KNASTER_GPU_LIB=knaster_b200/csrc/$v/libknaster_gpu.so timeout 200 python bench.py --workload $W $Q --steps 10 ${HT:+--host-threads $HT} > $O/ab${v}_$rep.json 2>/dev/null
  python -c "
import json;d=json.loads(open('$O/ab${v}_$rep.json').read().strip().splitlines()[-1]);print('$v',$rep,round(d['ms_per_step'],3),round(d['e2e']['ms_per_step'],3))"
done; done
if [ -n "$TESTS" ]; then timeout 600 python -m pytest $TESTS -x -q 2>&1 | tail -3; fi
if [ -n "$TESTS2" ]; then KNASTER_GPU_LIB=knaster_b200/csrc/${TESTLIB2:-_build_A}/libknaster_gpu.so timeout 600 python -m pytest $TESTS2 -x -q 2>&1 | tail -3; fi

#!/bin/bash
# Round-2: end-to-end call after the start-up work (parallel context layout / voice totals) and the overlapped reduction
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2m; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
for rep in 1 2 3; do for t in 15 3; do
  timeout 200 python bench.py $Q --steps 10 --host-threads $t > $O/e2e_ht${t}_$rep.json 2>/dev/null
  python -c "
import json;d=json.loads(open('$O/e2e_ht${t}_$rep.json').read().strip().splitlines()[-1]);print('threads=$t',$rep,round(d['ms_per_step'],3),round(d['e2e']['ms_per_step'],3))"
done; done
for t in 15 3; do KGPU_TIMING=1 timeout 300 python bench.py $Q --steps 2 --host-threads $t > $O/timing_ht$t.json 2> $O/timing_ht$t.err; echo "== $t threads"; grep -v "launch [0-9]*:\|stream_launch\|worker" $O/timing_ht$t.err | tail -6; done
[ -n "$TESTS" ] && timeout 900 python -m pytest $TESTS -x -q 2>&1 | tail -3

#!/bin/bash
# Round-2: scan kernel (two warps per voice) against the lane-per-voice kernel over bank sizes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2v; mkdir -p $O
Q="--no-parity --no-other-workloads --no-cpu-baseline"
for v in 64 256 512 1024 1536 2048; do
  KGPU_SUB_SCAN=1 timeout 100 python bench.py --voices $v --steps 3 $Q > $O/scan_v$v.json 2>/dev/null
  KGPU_SUB_SCAN=0 timeout 100 python bench.py --voices $v --steps 3 $Q > $O/lane_v$v.json 2>/dev/null
  python - <<PY
import json
a=json.loads(open("$O/scan_v$v.json").read().strip().splitlines()[-1]); b=json.loads(open("$O/lane_v$v.json").read().strip().splitlines()[-1])
print($v, "scan", round(a["ms_per_step"],2), a["kernels"], "lane", round(b["ms_per_step"],2), b["kernels"])
PY
done

#!/bin/bash
# cycle breakdown of the host control simulation on the bench configuration (built with -DKGPU_PROFILE_HOST)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
cd knaster_b200/csrc
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -I/usr/local/cuda/include -DKGPU_PROFILE_HOST -c -o /tmp/plan_prof.o plan.cpp && \
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /tmp/libknaster_gpu_prof.so /tmp/plan_prof.o _build/capi.o _build/jit.o _build/kernels.o _build/fused.o _build/fused_wt.o -lcudart -ldl
cd ../..
KGPU_TIMING=1 python - <<'PY' 2>&1 | grep "prof\|compile_events\|push 5"
import knaster_b200._ffi as F
F.LIB_PATH = '/tmp/libknaster_gpu_prof.so'
from knaster_b200 import banks
from knaster_b200.graph import Graph
g = Graph(0,2,64,48000)
banks.subtractive_bank(g, 16384, 10.0)
ev = g.take_events()
for i in range(3):
    F.debug_simulate(g, ev, 7500)
PY

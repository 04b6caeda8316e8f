// launch_args.h -- the argument block of the voice-bank kernels (fused.cu, fused_wt.cu, fused_scan.cuh and the kernels
// jit.cpp generates).  Plain data, no CUDA headers: NVRTC compiles this file too.
#pragma once
#include <stdint.h>

#include "dev.h"

namespace kgpu {

struct FusedArgs {
    const DevProgram *prog;
    uint32_t *regs;
    uint32_t n_voices;
    const DevEvent *events;
    const uint32_t *ev_off;
    uint32_t n_frames;
    float *partials;
    uint32_t row0;
    const DevTap *taps;
    uint32_t n_taps;
    float *tap_out;
    uint64_t tap_stride;
    uint64_t tap_frame0;
    const float *sine_table;
    const DevProgram *host_prog; // host copy of *prog (launch-time decisions)
    uint32_t block_size;
    void *scratch;               // fused_scratch_bytes() bytes of device memory, private to the launch
    uint32_t *regs_out;          // recipe-internal: where a kernel leaves the launch-end registers (default: regs)
    const float *ext;            // internal signals [signal][ext_stride] of this launch (plan.hpp HostPlan::signal_level); NULL if none
    uint32_t ext_stride;
};

} // namespace kgpu

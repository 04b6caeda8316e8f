// jit.hpp -- kernels generated per voice template ("render_jit").
//
// The hand-written recipes of fused.cu match four voice shapes; every other supported chain used to run on the plan
// interpreter (render_interp: node buffers in shared memory, a warp-uniform switch per node visit), 7-30x slower.
// Here the plan compiler emits, for ANY voice template, the CUDA source of a kernel in which the whole voice lives in
// registers -- one lane per voice, the nodes of the template inlined in topological order from the per-sample bodies of
// nodes.cuh, frames in straight-line groups of 8 (what graph_gen.rs:196-200's task loop does per node per block, turned
// into one basic block per 8 frames) -- compiles it with NVRTC for sm_100a (same numeric flags as the library:
// -fmad=false, IEEE division and square root, denormals kept) and caches the cubin by the hash of its source.
// Bit-identical to render_interp by construction: same node functions, same order, same event semantics.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "dev.h"

namespace kgpu {

// The CUDA source of the kernel for one voice template.  tapped: value slots whose per-frame values the kernel must be
// able to record (kgpu_plan_add_tap); returns "" if the template holds something the generator does not cover.
std::string jit_source(const DevProgram &p, const std::vector<uint16_t> &tapped_slots);

// Source -> cubin for sm_100a: from the cache directory if present, else through NVRTC (dlopen'ed; needs no GPU).
// false + `err` if NVRTC is unavailable or the compilation fails.
bool jit_cubin(const std::string &source, std::vector<char> &cubin, std::string &err, bool *from_cache = nullptr);

// A loaded kernel (cudaLibrary_t + cudaKernel_t), owned by the plan.
struct JitKernel {
    void *library = nullptr;
    void *kernel = nullptr;
};
bool jit_load(const std::vector<char> &cubin, JitKernel &k, std::string &err);
void jit_unload(JitKernel &k);

} // namespace kgpu

// fused.cu -- register-resident bank kernels for known voice shapes.
//
// A voice template that matches one of the recipes below is rendered by a kernel in which the
// whole voice (every node of the template) lives in registers for the entire launch: no node
// buffer is ever materialised, per-voice state is read from HBM once at launch start and
// written once at launch end, and the only per-frame memory traffic is the mix-bus staging in
// shared memory.  One lane = one voice, evaluated strictly sequentially in the reference's
// rounding order (the f32 phase / envelope recurrences cannot be re-associated).
//
//   recipe 0  "render_sub_asr"   PolyBlep(saw) -> SvfFilter -> (* EnvAsr.wr_mul) [MathUGen<Mul>]   (configs[2], [4])
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev.h"
#include "kernels.h"
#include "nodes.cuh"
#include "plan.hpp"

namespace kgpu {

namespace {

constexpr int SUB_TILE = 32;          // frames between mix-bus reductions
constexpr int SUB_PAD = 33;           // smem row stride (bank-conflict-free transpose)

// register file of the subtractive voice (absolute register indices inside the voice, fixed by
// compile_template's allocation order: node 0 PolyBlep, node 1 Svf, node 2 EnvAsr + WrMul)
enum : uint32_t {
    R_T = 0, R_DT = 1, R_USESIN = 2, R_PW = 3, R_WF = 4,
    R_IC1 = 5, R_IC2 = 6, R_A1 = 7, R_A2 = 8, R_A3 = 9, R_M0 = 10, R_M1 = 11, R_M2 = 12,
    R_EST = 13, R_ET = 14, R_AR = 15, R_RR = 16, R_SC = 17, R_GAIN = 18, SUB_NREGS = 19,
};

struct SubVoice {
    float t, dt;
    uint32_t use_sin;
    float ic1, ic2, a1, a2, a3, m0, m1, m2;
    uint32_t est;
    float et, ar, rr, sc, gain;
    KN_DEV void set(uint32_t reg, uint32_t bits) {
        const float f = __uint_as_float(bits);
        switch (reg) {
        case R_T: t = f; break;
        case R_DT: dt = f; break;
        case R_USESIN: use_sin = bits; break;
        case R_IC1: ic1 = f; break;
        case R_IC2: ic2 = f; break;
        case R_A1: a1 = f; break;
        case R_A2: a2 = f; break;
        case R_A3: a3 = f; break;
        case R_M0: m0 = f; break;
        case R_M1: m1 = f; break;
        case R_M2: m2 = f; break;
        case R_EST: est = bits; break;
        case R_ET: et = f; break;
        case R_AR: ar = f; break;
        case R_RR: rr = f; break;
        case R_SC: sc = f; break;
        case R_GAIN: gain = f; break;
        default: break; // pulse_width / waveform: not read by the sawtooth path
        }
    }
    // reference-order evaluation, any parameter values (used on tiles with events / odd dt)
    KN_DEV float tick() {
        const float saw = polyblep_saw_tick(t, dt, use_sin);
        const float y = svf_tick(saw, ic1, ic2, a1, a2, a3, m0, m1, m2);
        const float e = envasr_tick(est, et, ar, rr, sc) * gain; // WrMul, wrappers_core/math.rs:63-67
        return y * e;                                            // MathUGen<Mul>, math.rs:45-47
    }
};

// x - trunc(x) for x in [0, 2): trunc(x) is 0 or 1, and x - 1 is exact for x in [1, 2)
KN_DEV float wrap01(float x) { return x >= 1.0f ? x - 1.0f : x; }

// EnvAsr::next_sample (envelopes.rs:52-81) as straight-line selects: same values, no divergence
KN_DEV float envasr_tick_sel(uint32_t &st, float &t, float ar, float rr, float sc) {
    const bool att = st == ASR_ATTACKING, rel = st == ASR_RELEASING;
    const float cube = ((t * t) * t) * sc;
    const float out = att ? t : (st == ASR_SUSTAINING ? 1.0f : (rel ? cube : 0.0f));
    float tn = att ? t + ar : (rel ? t - rr : t);
    const bool to_sus = att && tn >= 1.0f;
    const bool to_stop = rel && tn <= 0.0f;
    st = to_sus ? (uint32_t)ASR_SUSTAINING : (to_stop ? (uint32_t)ASR_STOPPED : st);
    t = to_stop ? 0.0f : tn;
    return out;
}

constexpr int SUB_SUB = 8; // frames per straight-line group

// 8 frames, no events, 0 <= dt < 1 and !use_sin for every lane of the warp: the three
// recurrences (phase, filter, envelope) are independent chains that ptxas interleaves.
KN_DEV void sub_group_fast(SubVoice &s, float omd, float *out8) {
    float saw[SUB_SUB], env[SUB_SUB];
#pragma unroll
    for (int k = 0; k < SUB_SUB; k++) {
        // PolyBlep::saw (polyblep.rs:490-498) with t in [0,1): _t = frac(t + 0.5)
        const float _t = wrap01(s.t + 0.5f);
        float y = 2.0f * _t - 1.0f;
        if (_t < s.dt || _t > omd) y = y - blep(_t, s.dt); // 2 samples per period
        saw[k] = y;
        s.t = wrap01(s.t + s.dt); // inc(), polyblep.rs:232-235
    }
#pragma unroll
    for (int k = 0; k < SUB_SUB; k++) env[k] = envasr_tick_sel(s.est, s.et, s.ar, s.rr, s.sc) * s.gain;
#pragma unroll
    for (int k = 0; k < SUB_SUB; k++) out8[k] = svf_tick(saw[k], s.ic1, s.ic2, s.a1, s.a2, s.a3, s.m0, s.m1, s.m2) * env[k];
}

template <bool TAPS>
__global__ void __launch_bounds__(32, 8) render_sub_asr(FusedArgs a) {
    __shared__ float st[SUB_TILE * SUB_PAD];
    const uint32_t lane = threadIdx.x;
    const uint32_t gwarp = blockIdx.x;
    const uint32_t v = gwarp * 32 + lane;
    const uint32_t V = a.n_voices;
    const bool active = v < V;

    SubVoice s;
    {
        uint32_t r[SUB_NREGS];
#pragma unroll
        for (int i = 0; i < SUB_NREGS; i++) r[i] = active ? a.regs[(size_t)i * V + v] : 0u;
#pragma unroll
        for (int i = 0; i < SUB_NREGS; i++) s.set(i, r[i]);
    }
    uint32_t cur = 0, end = 0, next_frame = 0xFFFFFFFFu;
    if (a.events && active) {
        cur = a.ev_off[v];
        end = a.ev_off[v + 1];
        if (cur < end) next_frame = a.events[cur].frame;
    }
    int tap_row = -1;
    if (TAPS)
        for (uint32_t i = 0; i < a.n_taps; i++)
            if (a.taps[i].voice == v) tap_row = (int)a.taps[i].tap;

    float *prow = a.partials + (size_t)(a.row0 + gwarp) * a.n_frames;
    // the straight-line group needs t in [0,1), 0 <= dt < 1 and the sawtooth branch of next_sample
    bool fast_ok = __all_sync(0xFFFFFFFFu, !active || (s.dt >= 0.0f && s.dt < 1.0f && s.t >= 0.0f && s.t < 1.0f && !s.use_sin));
    float omd = 1.0f - s.dt;
    for (uint32_t f0 = 0; f0 < a.n_frames; f0 += SUB_TILE) {
        const uint32_t nf = min((uint32_t)SUB_TILE, a.n_frames - f0);
#pragma unroll 1
        for (uint32_t g0 = 0; g0 < SUB_TILE; g0 += SUB_SUB) {
            const uint32_t gf = f0 + g0;
            const bool ev_group = __any_sync(0xFFFFFFFFu, next_frame < gf + SUB_SUB);
            if (fast_ok && !ev_group && g0 + SUB_SUB <= nf) {
                float o[SUB_SUB];
                sub_group_fast(s, omd, o);
#pragma unroll
                for (int k = 0; k < SUB_SUB; k++) {
                    st[(g0 + k) * SUB_PAD + lane] = active ? o[k] : 0.f;
                    if (TAPS && tap_row >= 0) a.tap_out[(size_t)tap_row * a.tap_stride + a.tap_frame0 + gf + k] = o[k];
                }
            } else {
                for (uint32_t k = 0; k < SUB_SUB; k++) {
                    float o = 0.f;
                    if (g0 + k < nf) {
                        while (next_frame <= gf + k) { // events are sorted by (frame, node, arrival)
                            const DevEvent e = a.events[cur];
                            if (e.op == OP_SET) s.set(e.reg, e.value);
                            else if (e.op == OP_ASR_RELEASE) envasr_release(s.est, s.et, s.sc);
                            cur++;
                            next_frame = cur < end ? a.events[cur].frame : 0xFFFFFFFFu;
                        }
                        o = s.tick();
                        if (TAPS && tap_row >= 0) a.tap_out[(size_t)tap_row * a.tap_stride + a.tap_frame0 + gf + k] = o;
                    }
                    st[(g0 + k) * SUB_PAD + lane] = active ? o : 0.f;
                }
                if (ev_group) {
                    omd = 1.0f - s.dt;
                    fast_ok = __all_sync(0xFFFFFFFFu, !active || (s.dt >= 0.0f && s.dt < 1.0f && s.t >= 0.0f && s.t < 1.0f && !s.use_sin));
                }
            }
        }
        __syncwarp();
        // lane l sums frame l over the warp's 32 voices (fixed order => deterministic)
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j++) acc = acc + st[lane * SUB_PAD + j];
        if (lane < nf) prow[f0 + lane] = acc;
        __syncwarp();
    }
    if (active) {
        a.regs[(size_t)R_T * V + v] = __float_as_uint(s.t);
        a.regs[(size_t)R_IC1 * V + v] = __float_as_uint(s.ic1);
        a.regs[(size_t)R_IC2 * V + v] = __float_as_uint(s.ic2);
        a.regs[(size_t)R_EST * V + v] = s.est;
        a.regs[(size_t)R_ET * V + v] = __float_as_uint(s.et);
        a.regs[(size_t)R_SC * V + v] = __float_as_uint(s.sc);
        // parameter registers change only through events: write them back as well
        a.regs[(size_t)R_DT * V + v] = __float_as_uint(s.dt);
        a.regs[(size_t)R_USESIN * V + v] = s.use_sin;
        a.regs[(size_t)R_A1 * V + v] = __float_as_uint(s.a1);
        a.regs[(size_t)R_A2 * V + v] = __float_as_uint(s.a2);
        a.regs[(size_t)R_A3 * V + v] = __float_as_uint(s.a3);
        a.regs[(size_t)R_M0 * V + v] = __float_as_uint(s.m0);
        a.regs[(size_t)R_M1 * V + v] = __float_as_uint(s.m1);
        a.regs[(size_t)R_M2 * V + v] = __float_as_uint(s.m2);
        a.regs[(size_t)R_AR * V + v] = __float_as_uint(s.ar);
        a.regs[(size_t)R_RR * V + v] = __float_as_uint(s.rr);
        a.regs[(size_t)R_GAIN * V + v] = __float_as_uint(s.gain);
    }
}

bool match_sub_asr(const DevProgram &p) {
    if (p.n_nodes != 4 || p.n_regs != SUB_NREGS || p.n_ubus != 1) return false;
    const DevNode &saw = p.nodes[0], &svf = p.nodes[1], &env = p.nodes[2], &mul = p.nodes[3];
    if (saw.kind != DK_POLYBLEP || saw.mode != 0 || saw.n_post || saw.n_ar || saw.reg != R_T) return false;
    if (svf.kind != DK_SVF || svf.n_post || svf.n_ar || svf.reg != R_IC1 || svf.in_slot[0] != (int)saw.out_slot[0]) return false;
    if (env.kind != DK_ENVASR || env.n_post != 1 || env.post_op[0] != PO_MUL || env.post_reg[0] != R_GAIN || env.n_ar || env.reg != R_EST) return false;
    if (mul.kind != DK_MATH || mul.mode != 2 || mul.n_out != 1 || mul.n_post || mul.n_ar) return false;
    if (mul.in_slot[0] != (int)svf.out_slot[0] || mul.in_slot[1] != (int)env.out_slot[0]) return false;
    return p.ubus_slot[0] == mul.out_slot[0];
}

} // namespace

int match_fused_recipe(const DevProgram &p) {
    if (match_sub_asr(p)) return 0;
    return -1;
}
const char *fused_recipe_name(int recipe) {
    switch (recipe) {
    case 0: return "render_sub_asr";
    default: return "render_interp";
    }
}
uint32_t fused_rows(int recipe, uint32_t n_voices, uint32_t n_ubus) {
    (void)recipe;
    return ((n_voices + 31) / 32) * n_ubus; // one partial row per warp
}
cudaError_t launch_fused(int recipe, const FusedArgs &a, cudaStream_t stream) {
    if (recipe != 0) return cudaErrorNotSupported;
    // one warp per CTA: 512 warps spread over all 148 SMs (3-4 per SM, one per SM sub-partition)
    const uint32_t n_warps = (a.n_voices + 31) / 32;
    if (a.n_taps) render_sub_asr<true><<<n_warps, 32, 0, stream>>>(a);
    else render_sub_asr<false><<<n_warps, 32, 0, stream>>>(a);
    return cudaGetLastError();
}

} // namespace kgpu

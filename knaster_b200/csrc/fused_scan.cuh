// fused_scan.cuh -- "render_sub_scan": the subtractive recipe for SMALL banks, time-parallel.
// Included by fused.cu (inside its anonymous namespace, after SubVoice / AsrEnv / EvCursor).
//
// render_sub_asr gives a voice one lane for the whole render: a launch takes as long as ONE voice
// whatever the bank size (14.5 ms per 10 s), so a bank of a few hundred voices leaves > 95 % of the
// machine idle.  Here a voice owns a whole WARP and the lanes are the 32 frames of a chunk:
//
//   pre-pass (sequential, reference rounding order -- SURVEY F5: the f32 phase and envelope ramps
//       drift 4.4e-3 over 10 s if re-associated): t_{k+1} = wrap01(t_k + dt), et_{k+1} = et_k + delta
//       for the chunk's 32 frames, all lanes in step, lane 0 leaves (t_k, et_k) in shared memory;
//   frame-parallel: lane k evaluates saw + blep at t_k (polyblep.rs:490-498) and the envelope at et_k
//       (envelopes.rs:52-81) exactly as the reference does;
//   SvfFilter (svf.rs:245-280) as a chunked linear-recurrence scan: the two states obey
//       s_{k+1} = A s_k + b x_k,  A = [[2 a1 - 1, -2 a2], [2 a2, 1 - 2 a3]],  b = [2 a2, 2 a3]  (SURVEY App. A.4),
//       so the states before every frame of the chunk are s_k = A^k s_0 + sum_{j<k} A^(k-1-j) b x_j: a
//       Hillis-Steele scan of 2-vectors over the warp with shuffle carries (5 rounds; the powers A^(2^r)
//       and A^lane only change when a coefficient does and are kept in registers), then every lane forms its
//       frame's v1, v2 and output from its own pre-update state.  This re-associates the filter's sums:
//       <= 1e-4 against the reference (BASELINE.json's budget for scan-reordered IIR filters; measured
//       ~1e-6), where render_sub_asr is bit-identical.
//
// Measured (256 voices x 10 s, one warp per scheduler): 6.7 ms against 13.0 ms for render_sub_asr.  The chunk is latency-bound:
// the frame-parallel half is a ~320-cycle dependent path (saw, 5 shuffle rounds, state fix-up) and the pre-pass a ~530-cycle
// one (32 x add / compare / subtract); software-pipelining them against each other (PIPE) took 8.1 -> 6.7 ms.  Tried and
// reverted, both parity-clean: running the phase chain without its wrap and redoing only the chunks (later: only the
// frames) behind a wrap -- 7.1 ms, the redo's extra branches and barriers cost what the shorter chain saved; two chunks
// per iteration with independent scans -- 10.0 ms.
//
// Chunks that hold a parameter event, may move the envelope's state machine, or lie outside the
// straight-line domain (other waveforms, dt >= 1/4, ...) run the reference-order per-frame code on
// all lanes instead (a few per note), so events and envelope transitions stay sample-exact.
// One warp per CTA: a bank of V voices is V CTAs spread over all SMs.

constexpr int SCAN_CHUNK = 32;

struct ScanK {
    float a[4];        // A, row-major
    float b[2];
    float pw[5][4];    // A^(2^r), r = 0..4
    float p32[4];      // A^32
    float pk[4];       // A^lane
    KN_DEV static void mul(const double *x, const double *y, double *z) { // z = x y (2x2, row-major)
        const double z0 = x[0] * y[0] + x[1] * y[2], z1 = x[0] * y[1] + x[1] * y[3];
        const double z2 = x[2] * y[0] + x[3] * y[2], z3 = x[2] * y[1] + x[3] * y[3];
        z[0] = z0; z[1] = z1; z[2] = z2; z[3] = z3;
    }
    // powers in f64 (a handful of operations per coefficient change), rounded once
    KN_DEV void build(float a1, float a2, float a3, uint32_t lane) {
        double A[4] = {2.0 * (double)a1 - 1.0, -2.0 * (double)a2, 2.0 * (double)a2, 1.0 - 2.0 * (double)a3};
        b[0] = 2.0f * a2;
        b[1] = 2.0f * a3;
        double P[4] = {A[0], A[1], A[2], A[3]};
        double K[4] = {1.0, 0.0, 0.0, 1.0};
#pragma unroll
        for (int r = 0; r < 5; r++) {
#pragma unroll
            for (int i = 0; i < 4; i++) pw[r][i] = (float)P[i];
            if ((lane >> r) & 1u) mul(P, K, K);
            mul(P, P, P);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            a[i] = (float)A[i];
            p32[i] = (float)P[i];
            pk[i] = (float)K[i];
        }
    }
};

template <bool TAPS, bool PIPE>
__global__ void __launch_bounds__(32, 16) render_sub_scan(FusedArgs a) {
    __shared__ float2 pre[2][SCAN_CHUNK];
    const uint32_t lane = threadIdx.x;
    const uint32_t v = blockIdx.x;
    const uint32_t V = a.n_voices;
    constexpr uint32_t FULL = 0xFFFFFFFFu;

    // the voice's registers, identical on every lane
    SubVoice<AsrEnv> s;
    {
        uint32_t r[R_EST];
#pragma unroll
        for (int i = 0; i < R_EST; i++) r[i] = a.regs[(size_t)i * V + v];
#pragma unroll
        for (int i = 0; i < R_EST; i++) s.set_core(i, r[i]);
        s.e.load(a, v);
    }
    uint32_t cur = 0, end = 0, next_frame = 0xFFFFFFFFu;
    if (a.events) {
        cur = a.ev_off[v];
        end = a.ev_off[v + 1];
        if (cur < end) next_frame = __ldg(&a.events[cur].frame);
    }
    float *tap = nullptr;
    if (TAPS)
        for (uint32_t i = 0; i < a.n_taps; i++)
            if (a.taps[i].voice == v) tap = a.tap_out + (size_t)a.taps[i].tap * a.tap_stride + a.tap_frame0;
    float *prow = a.partials + (size_t)(a.row0 + v) * a.n_frames;

    ScanK K;
    K.build(s.a1, s.a2, s.a3, lane);
    float omd = 1.0f - s.dt, rc = div_prep(s.dt);
    AsrEnv::D d;
    s.e.derive(d);
    const uint32_t NF = a.n_frames;
    const bool writer = lane == 0;

    // the pre-pass of one chunk: the two f32 recurrences, sequentially, in the reference's rounding order.  Every lane
    // runs the same chain (the stores are predicated); lane 0 leaves (t_k, et_k) in `dst`.
    auto prepass = [&](float2 *dst) {
        float t = s.t, et = s.e.et;
#pragma unroll
        for (int k = 0; k < SCAN_CHUNK; k++) {
            if (writer) dst[k] = make_float2(t, et);
            t = wrap01(t + s.dt);      // inc(), polyblep.rs:232-235
            et = et + d.delta;         // envelopes.rs:58-66 with the state fixed over the chunk
        }
        s.t = t;
        s.e.et = (d.att || d.rel) ? et : s.e.et;
    };
    // whether the chunk starting at f0 can take the scan path, given the state at its first frame
    auto chunk_fast = [&](uint32_t f0) {
        return f0 + SCAN_CHUNK <= NF && next_frame >= f0 + SCAN_CHUNK && sub_lane_fast(s) && s.e.safe_frames() >= (uint32_t)SCAN_CHUNK;
    };
    // the frame-parallel half of a chunk: lane k renders frame f0 + k from (t_k, et_k)
    auto parallel = [&](const float2 pk, uint32_t f0) {
        const bool ramp = d.att || d.rel;
        const float x = saw_eval(pk.x, s.dt, omd, rc);             // saw + blep, polyblep.rs:490-498
        const float tl = ramp ? pk.y : d.cval;
        const float u = d.rel ? pk.y : 1.0f;
        const float env = (((tl * u) * u) * d.sc2) * s.e.gain;     // EnvAsr::next_sample, then WrMul (wrappers_core/math.rs:63-67)
        // SvfFilter: inclusive scan of c_k = sum_{j<=k} A^(k-j) b x_j
        float c1 = K.b[0] * x, c2 = K.b[1] * x;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const float u1 = __shfl_up_sync(FULL, c1, 1u << r), u2 = __shfl_up_sync(FULL, c2, 1u << r);
            if (lane >= (1u << r)) {
                c1 = __fmaf_rn(K.pw[r][0], u1, __fmaf_rn(K.pw[r][1], u2, c1));
                c2 = __fmaf_rn(K.pw[r][2], u1, __fmaf_rn(K.pw[r][3], u2, c2));
            }
        }
        // the state before this lane's frame: s_k = A^k s_0 + c_(k-1)
        float e1 = __shfl_up_sync(FULL, c1, 1), e2 = __shfl_up_sync(FULL, c2, 1);
        if (lane == 0) e1 = e2 = 0.0f;
        const float ic1 = __fmaf_rn(K.pk[0], s.ic1, __fmaf_rn(K.pk[1], s.ic2, e1));
        const float ic2 = __fmaf_rn(K.pk[2], s.ic1, __fmaf_rn(K.pk[3], s.ic2, e2));
        // svf.rs:272-278 from the pre-update state
        const float v3 = x - ic2;
        const float v1 = s.a1 * ic1 + s.a2 * v3;
        const float v2 = (ic2 + s.a2 * ic1) + s.a3 * v3;
        const float y = (s.m0 * x + s.m1 * v1) + s.m2 * v2;
        const float out = y * env;                                  // MathUGen<Mul>, math.rs:45-47
        // the state after the chunk: s_32 = A^32 s_0 + c_31
        const float l1 = __shfl_sync(FULL, c1, 31), l2 = __shfl_sync(FULL, c2, 31);
        const float n1 = __fmaf_rn(K.p32[0], s.ic1, __fmaf_rn(K.p32[1], s.ic2, l1));
        const float n2 = __fmaf_rn(K.p32[2], s.ic1, __fmaf_rn(K.p32[3], s.ic2, l2));
        s.ic1 = n1;
        s.ic2 = n2;
        prow[f0 + lane] = out;
        if (TAPS && tap) tap[f0 + lane] = out;
    };

    bool have = false;   // the pre-pass of the chunk at f0 is already in pre[buf] (PIPE)
    uint32_t buf = 0;
#pragma unroll 1
    for (uint32_t f0 = 0; f0 < NF; f0 += SCAN_CHUNK) {
        if (have || chunk_fast(f0)) {
            if (!have) {
                prepass(pre[buf]);
                __syncwarp();
            }
            const float2 pk = pre[buf][lane];
            // PIPE: the NEXT chunk's pre-pass (a latency-bound chain of 32 x 3 dependent operations) is issued in the same
            // basic block as this chunk's frame-parallel half (shuffle latencies): each fills the other's stalls.  Its
            // inputs are all known here -- the state at the next chunk's first frame is where this chunk's pre-pass ended.
            if (PIPE && chunk_fast(f0 + SCAN_CHUNK)) {
                prepass(pre[buf ^ 1]);
                parallel(pk, f0);
                have = true;
            } else {
                parallel(pk, f0);
                have = false;
            }
            __syncwarp();
            buf ^= 1;
        } else {
            // ---- exact chunk: events applied at their frames, every frame in reference order, all lanes in step
            float out = 0.0f;
            bool touched = false;
            const uint32_t nf = min((uint32_t)SCAN_CHUNK, NF - f0);
#pragma unroll 1
            for (uint32_t k = 0; k < nf; k++) {
                while (next_frame <= f0 + k) { // sorted by (frame, node, arrival)
                    const DevEvent e = ldg_event(a.events + cur);
                    if (e.op == OP_SET) s.set(e.reg, e.value);
                    else s.e.op(e);
                    cur++;
                    next_frame = cur < end ? __ldg(&a.events[cur].frame) : 0xFFFFFFFFu;
                    touched = true;
                }
                const float o = s.tick();
                if (lane == k) out = o;
            }
            if (touched) {
                K.build(s.a1, s.a2, s.a3, lane);
                omd = 1.0f - s.dt;
                rc = div_prep(s.dt);
            }
            s.e.derive(d);
            if (f0 + lane < NF) {
                prow[f0 + lane] = out;
                if (TAPS && tap) tap[f0 + lane] = out;
            }
        }
    }
    if (lane == 0) {
        const uint32_t regs_out[R_EST] = {__float_as_uint(s.t), __float_as_uint(s.dt), s.use_sin, __float_as_uint(s.pw), s.wf,
                                           __float_as_uint(s.ic1), __float_as_uint(s.ic2), __float_as_uint(s.a1), __float_as_uint(s.a2),
                                           __float_as_uint(s.a3), __float_as_uint(s.m0), __float_as_uint(s.m1), __float_as_uint(s.m2)};
#pragma unroll
        for (int i = 0; i < R_EST; i++) a.regs[(size_t)i * V + v] = regs_out[i];
        s.e.store(a, v);
    }
}

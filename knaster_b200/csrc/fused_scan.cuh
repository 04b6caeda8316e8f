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

// The phase recurrence t' = wrap01(t + dt) is the longest dependent chain of a chunk: FADD -> compare -> FADD.  Two changes, both
// with identical values:
//   * rounding is monotone, so "fl(t + dt) >= 1" is the same predicate as "t >= T" for T = the smallest float whose sum with dt
//     rounds to 1 or more -- a compare on the OLD phase, which issues beside the addition instead of behind it.  T is found by
//     stepping from fl(1 - dt) one ulp at a time (at most a few steps); valid for dt in (0, 1), t in [0, 1), then T in [0.5, 1];
//   * the 0/1 flag of that compare is made on the FMA pipe (4 cycles) instead of by FSET on the ALU pipe (~8): for t and T in
//     [0.5, 1) both are multiples of 2^-24, so (t - T) * 2^24 + 1 is an integer, >= 1 exactly when t >= T and <= 0 otherwise; a
//     smaller t only makes it more negative.  One FFMA.SAT (wrap_flag); the constant 1 - T * 2^24 is exact (|.| < 2^24).
// Together: 8 cycles per frame instead of 16.
KN_DEV float wrap_flag_const(float T) { return 1.0f - T * 16777216.0f; }
KN_DEV float wrap_flag(float t, float c) { return __saturatef(__fmaf_rn(t, 16777216.0f, c)); }
KN_DEV float wrap_threshold(float dt) {
    float c = 1.0f - dt;
    for (int i = 0; i < 4 && __fadd_rn(c, dt) >= 1.0f; i++) c = __uint_as_float(__float_as_uint(c) - 1u);
    for (int i = 0; i < 8 && __fadd_rn(c, dt) < 1.0f; i++) c = __uint_as_float(__float_as_uint(c) + 1u);
    return c;
}

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
    float omd = 1.0f - s.dt, rc = div_prep(s.dt), tstar = wrap_flag_const(wrap_threshold(s.dt));
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
            t = (t + s.dt) - wrap_flag(t, tstar);        // inc(), polyblep.rs:232-235: wrap01(t + dt), see wrap_threshold
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
                tstar = wrap_flag_const(wrap_threshold(s.dt));
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

// ------------------------------------------------------------------------------------------------------------------------------
// "render_sub_scan2": the same chunks on TWO warps per voice.  In render_sub_scan one warp carries both halves of a chunk and
// its time is the sum of two latency-bound paths that the compiler overlaps only in part.  Here the halves are roles:
//   warp 0 (phase warp)  runs the sequential f32 recurrences and nothing else -- 8 cycles per frame (wrap_threshold) -- and
//       leaves (t_k, et_k) of every frame in a ring of chunks in shared memory; it also owns the decision "scan or exact";
//   warp 1 (filter warp) takes the chunks from the ring: saw + blep, envelope, the SvfFilter scan, output -- two chunks per
//       trip when both are there, so that the shuffle rounds of one scan fill the stalls of the other.
// Warps of a CTA sit on different SM sub-partitions, so the two chains run side by side: a bank of V voices is V CTAs of 64.
// Hand-over: named barriers, the producer / consumer pattern of the PTX manual (st.shared; bar.arrive -- bar.sync; ld.shared): per ring
// slot one barrier "full" (phase warp arrives, filter warp waits) and one "empty" (the reverse), plus one for the return of an
// exact chunk.  (A first version used sequence counters behind __threadfence_block(): the fence also waits for the filter warp's
// global stores, 7.9 ms per step against 6.1 for the one-warp kernel.)  An EXACT chunk (parameter event, envelope transition,
// outside the straight-line domain) is rendered by the filter warp in reference order, as in render_sub_scan; the phase warp
// hands it (t, et), waits, and takes the voice's registers back.
constexpr int SCAN2_RING = 4;     // chunks between the two warps
constexpr int SCAN2_WORDS = 19;   // the voice's registers: 13 core + 6 envelope

KN_DEV void scan2_pack(const SubVoice<AsrEnv> &s, uint32_t *w) {
    w[0] = __float_as_uint(s.t); w[1] = __float_as_uint(s.dt); w[2] = s.use_sin; w[3] = __float_as_uint(s.pw); w[4] = s.wf;
    w[5] = __float_as_uint(s.ic1); w[6] = __float_as_uint(s.ic2); w[7] = __float_as_uint(s.a1); w[8] = __float_as_uint(s.a2);
    w[9] = __float_as_uint(s.a3); w[10] = __float_as_uint(s.m0); w[11] = __float_as_uint(s.m1); w[12] = __float_as_uint(s.m2);
    w[13] = s.e.est; w[14] = __float_as_uint(s.e.et); w[15] = __float_as_uint(s.e.ar); w[16] = __float_as_uint(s.e.rr);
    w[17] = __float_as_uint(s.e.sc); w[18] = __float_as_uint(s.e.gain);
}
KN_DEV void scan2_unpack(SubVoice<AsrEnv> &s, const uint32_t *w) {
#pragma unroll
    for (int i = 0; i < R_EST; i++) s.set_core(i, w[i]);
#pragma unroll
    for (int i = 0; i < 6; i++) s.e.set(R_EST + i, w[R_EST + i]);
}
// barrier ids (0 is __syncthreads'): full[slot] = 1 + slot, empty[slot] = 1 + RING + slot, returned = 1 + 2 * RING; 64 threads each.
// Immediate ids: a barrier named by a register costs a warp-synchronising prologue of ~100 cycles per use (measured, ncu r2c).
template <int ID> KN_DEV void scan2_sync() { asm volatile("bar.sync %0, 64;" ::"n"(ID) : "memory"); }
template <int ID> KN_DEV void scan2_arrive() { asm volatile("bar.arrive %0, 64;" ::"n"(ID) : "memory"); }
constexpr int SCAN2_FULL = 1, SCAN2_EMPTY = 1 + SCAN2_RING, SCAN2_RETURNED = 1 + 2 * SCAN2_RING;
// ... and ONE copy of everything else (the kernel has to stay inside the instruction cache: two warps of a CTA run different code):
// the slot is a run-time value, only the barrier instruction itself is picked by a (warp-uniform) switch
template <int BASE> KN_DEV void scan2_sync_slot(uint32_t slot) {
    static_assert(SCAN2_RING == 4, "one case per ring slot");
    switch (slot) {
    case 0: scan2_sync<BASE + 0>(); break;
    case 1: scan2_sync<BASE + 1>(); break;
    case 2: scan2_sync<BASE + 2>(); break;
    default: scan2_sync<BASE + 3>(); break;
    }
}
template <int BASE> KN_DEV void scan2_arrive_slot(uint32_t slot) {
    switch (slot) {
    case 0: scan2_arrive<BASE + 0>(); break;
    case 1: scan2_arrive<BASE + 1>(); break;
    case 2: scan2_arrive<BASE + 2>(); break;
    default: scan2_arrive<BASE + 3>(); break;
    }
}

template <bool TAPS>
__global__ void __launch_bounds__(64, 8) render_sub_scan2(FusedArgs a) {
    __shared__ __align__(16) float ring_t[SCAN2_RING][SCAN_CHUNK];   // the phase before every frame of the chunk ([0]: at an exact chunk's start)
    __shared__ __align__(16) float ring_et[SCAN2_RING][SCAN_CHUNK];  // the envelope ramp likewise (written while the envelope ramps)
    __shared__ uint32_t ring_exact[SCAN2_RING];
    __shared__ uint32_t hand[SCAN2_WORDS];
    const uint32_t lane = threadIdx.x & 31u;
    const bool phase_warp = threadIdx.x < 32u;
    const uint32_t v = blockIdx.x;
    const uint32_t V = a.n_voices;
    constexpr uint32_t FULL = 0xFFFFFFFFu;

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
    const uint32_t NF = a.n_frames, NC = (NF + SCAN_CHUNK - 1) / SCAN_CHUNK;
    AsrEnv::D d;
    s.e.derive(d);

    if (phase_warp) {
        float tstar = wrap_flag_const(wrap_threshold(s.dt));
        // Whether a chunk takes the scan path is a function of the state at ITS first frame only (so that a render does not depend on
        // how it is cut into launches): no event inside, the straight-line domain, no envelope transition possible inside.
        bool lane_fast = sub_lane_fast(s);                 // the voice's parameters: they only change in exact chunks
#pragma unroll 1
        for (uint32_t c = 0; c < NC; c++) {
            const uint32_t SLOT = c % SCAN2_RING;
            const uint32_t f0 = c * SCAN_CHUNK;
            if (c >= SCAN2_RING) scan2_sync_slot<SCAN2_EMPTY>(SLOT);              // the slot is free again
            const bool fast = lane_fast && f0 + SCAN_CHUNK <= min(next_frame, NF) &&
                              (!(d.att || d.rel) || s.e.safe_frames() >= (uint32_t)SCAN_CHUNK);
            if (fast) {
                // the f32 recurrences of the chunk, sequentially, in the reference's rounding order (every lane the same chain;
                // lane 0 leaves the values before each frame in the ring)
                float t = s.t;
                float4 *dt4 = reinterpret_cast<float4 *>(ring_t[SLOT]);
                if (d.att || d.rel) {
                    float et = s.e.et;
                    float4 *de4 = reinterpret_cast<float4 *>(ring_et[SLOT]);
#pragma unroll
                    for (int k = 0; k < SCAN_CHUNK; k += 4) {
                        float tt[4], ee[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            tt[j] = t;
                            ee[j] = et;
                            t = (t + s.dt) - wrap_flag(t, tstar);                 // inc(), polyblep.rs:232-235 (wrap_threshold)
                            et = et + d.delta;                                    // envelopes.rs:58-66, state fixed over the chunk
                        }
                        if (lane == 0) {
                            dt4[k / 4] = make_float4(tt[0], tt[1], tt[2], tt[3]);
                            de4[k / 4] = make_float4(ee[0], ee[1], ee[2], ee[3]);
                        }
                    }
                    s.e.et = et;
                } else {
#pragma unroll
                    for (int k = 0; k < SCAN_CHUNK; k += 4) {
                        float tt[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            tt[j] = t;
                            t = (t + s.dt) - wrap_flag(t, tstar);
                        }
                        if (lane == 0) dt4[k / 4] = make_float4(tt[0], tt[1], tt[2], tt[3]);
                    }
                }
                s.t = t;
                if (lane == 0) ring_exact[SLOT] = 0u;
                scan2_arrive_slot<SCAN2_FULL>(SLOT);
            } else {
                if (lane == 0) {
                    ring_t[SLOT][0] = s.t;
                    ring_et[SLOT][0] = s.e.et;
                    ring_exact[SLOT] = 1u;
                }
                scan2_arrive_slot<SCAN2_FULL>(SLOT);
                // the filter warp renders this chunk in reference order and returns the voice
                scan2_sync<SCAN2_RETURNED>();
                uint32_t w[SCAN2_WORDS];
#pragma unroll
                for (int i = 0; i < SCAN2_WORDS; i++) w[i] = hand[i];
                scan2_unpack(s, w);
                const uint32_t fe = min(f0 + SCAN_CHUNK, NF);
                while (next_frame < fe) { // the events of the chunk are applied (by the filter warp): step over them
                    cur++;
                    next_frame = cur < end ? __ldg(&a.events[cur].frame) : 0xFFFFFFFFu;
                }
                s.e.derive(d);
                tstar = wrap_flag_const(wrap_threshold(s.dt));
                lane_fast = sub_lane_fast(s);
            }
        }
        // the phase warp's share of the state: the phase and the envelope
        if (lane == 0) {
            a.regs[(size_t)0 * V + v] = __float_as_uint(s.t);
            s.e.store(a, v);
        }
        return;
    }

    // ---- filter warp
    float *tap = nullptr;
    if (TAPS)
        for (uint32_t i = 0; i < a.n_taps; i++)
            if (a.taps[i].voice == v) tap = a.tap_out + (size_t)a.taps[i].tap * a.tap_stride + a.tap_frame0;
    float *prow = a.partials + (size_t)(a.row0 + v) * a.n_frames;
    ScanK K;
    K.build(s.a1, s.a2, s.a3, lane);
    float omd = 1.0f - s.dt, rc = div_prep(s.dt);

    // the scan of one chunk up to the inclusive prefix (c1, c2) and the input x: no dependence on the filter's state
    auto scan = [&](float t, float et, float &x, float &env, float &c1, float &c2) {
        const bool ramp = d.att || d.rel;
        x = saw_eval(t, s.dt, omd, rc);                            // saw + blep, polyblep.rs:490-498
        const float tl = ramp ? et : d.cval;
        const float u = d.rel ? et : 1.0f;
        env = (((tl * u) * u) * d.sc2) * s.e.gain;                 // EnvAsr::next_sample, then WrMul
        c1 = K.b[0] * x, c2 = K.b[1] * x;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const float u1 = __shfl_up_sync(FULL, c1, 1u << r), u2 = __shfl_up_sync(FULL, c2, 1u << r);
            if (lane >= (1u << r)) {
                c1 = __fmaf_rn(K.pw[r][0], u1, __fmaf_rn(K.pw[r][1], u2, c1));
                c2 = __fmaf_rn(K.pw[r][2], u1, __fmaf_rn(K.pw[r][3], u2, c2));
            }
        }
    };
    // the state-dependent rest: every lane's pre-update state, its frame's output, the state after the chunk
    auto finish = [&](float x, float env, float c1, float c2, uint32_t f0) {
        float e1 = __shfl_up_sync(FULL, c1, 1), e2 = __shfl_up_sync(FULL, c2, 1);
        if (lane == 0) e1 = e2 = 0.0f;
        const float l1 = __shfl_sync(FULL, c1, 31), l2 = __shfl_sync(FULL, c2, 31);
        const float ic1 = __fmaf_rn(K.pk[0], s.ic1, __fmaf_rn(K.pk[1], s.ic2, e1));
        const float ic2 = __fmaf_rn(K.pk[2], s.ic1, __fmaf_rn(K.pk[3], s.ic2, e2));
        const float n1 = __fmaf_rn(K.p32[0], s.ic1, __fmaf_rn(K.p32[1], s.ic2, l1));
        const float n2 = __fmaf_rn(K.p32[2], s.ic1, __fmaf_rn(K.p32[3], s.ic2, l2));
        s.ic1 = n1;
        s.ic2 = n2;
        const float v3 = x - ic2;                                   // svf.rs:272-278 from the pre-update state
        const float v1 = s.a1 * ic1 + s.a2 * v3;
        const float v2 = (ic2 + s.a2 * ic1) + s.a3 * v3;
        const float y = (s.m0 * x + s.m1 * v1) + s.m2 * v2;
        const float out = y * env;                                  // MathUGen<Mul>, math.rs:45-47
        prow[f0 + lane] = out;
        if (TAPS && tap) tap[f0 + lane] = out;
    };
    // a slot is handed back as soon as it has been read -- except the last RING chunks, whose "empty" nobody would wait for
    auto release = [&](uint32_t c) {
        if (c + SCAN2_RING < NC) scan2_arrive_slot<SCAN2_EMPTY>(c % SCAN2_RING);
    };
    // Two chunks per trip when both are scan chunks: the two scans run side by side.  (The second chunk's barrier is only waited on
    // when the first is a scan chunk: behind an exact chunk the phase warp is itself waiting for this warp.)
    uint32_t c = 0;
    bool synced = false; // chunk c's "full" barrier has been passed already
#pragma unroll 1
    while (c < NC) {
        const uint32_t S0 = c % SCAN2_RING, S1 = (c + 1) % SCAN2_RING, f0 = c * SCAN_CHUNK;
        if (!synced) scan2_sync_slot<SCAN2_FULL>(S0);
        synced = false;
        if (__shfl_sync(FULL, ring_exact[S0], 0) != 0u) { // (one lane's reading: the branches hold barriers)
            // ---- exact chunk: events applied at their frames, every frame in reference order, all lanes in step
            s.t = ring_t[S0][0];
            s.e.et = ring_et[S0][0];
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
            if (lane == 0) {
                uint32_t w[SCAN2_WORDS];
                scan2_pack(s, w);
#pragma unroll
                for (int i = 0; i < SCAN2_WORDS; i++) hand[i] = w[i];
            }
            release(c);
            scan2_arrive<SCAN2_RETURNED>();
            c += 1;
            continue;
        }
        bool two = false;
        if (c + 1 < NC) {
            scan2_sync_slot<SCAN2_FULL>(S1);
            two = __shfl_sync(FULL, ring_exact[S1], 0) == 0u;
            synced = !two;
        }
        if (two) {
            const float ta = ring_t[S0][lane], ea = ring_et[S0][lane], tb = ring_t[S1][lane], eb = ring_et[S1][lane];
            release(c);
            release(c + 1);
            float xa, va, a1, a2, xb, vb, b1, b2;
            scan(ta, ea, xa, va, a1, a2);
            scan(tb, eb, xb, vb, b1, b2);
            finish(xa, va, a1, a2, f0);
            finish(xb, vb, b1, b2, f0 + SCAN_CHUNK);
            c += 2;
        } else {
            const float ta = ring_t[S0][lane], ea = ring_et[S0][lane];
            release(c);
            float xa, va, a1, a2;
            scan(ta, ea, xa, va, a1, a2);
            finish(xa, va, a1, a2, f0);
            c += 1;
        }
    }
    // the filter warp's share of the state: everything but the phase and the envelope
    if (lane == 0) {
        uint32_t w[SCAN2_WORDS];
        scan2_pack(s, w);
#pragma unroll
        for (int i = 1; i < R_EST; i++) a.regs[(size_t)i * V + v] = w[i];
    }
}

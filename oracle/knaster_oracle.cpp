// knaster_oracle.cpp -- CPU oracle: a C++ restatement of knaster's f32 CPU render path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE (see knaster_oracle.h).
//
// Architecture is deliberately knaster's own (so that it doubles as the "C++
// restatement of knaster's CPU path" baseline): one output buffer per node,
// block-at-a-time processing, one virtual dispatch per node per block, wrappers
// nested around UGens, left-fold Add chains, scheduling-event ring + waiting
// queue.  All file:line citations are relative to /root/reference.
//
// Build: -O2 -ffp-contract=off -fno-fast-math (rustc/LLVM never contracts a*b+c
// and never reassociates).  f32 transcendental functions are glibc's sinf/cosf/
// tanf/expf/powf, the same libm knaster's `std` feature binds through
// num-traits (knaster_primitives/Cargo.toml:12,19).
//
// Parity pinning: see the header.  "parity unpinned" for the DSP bodies that the
// reference's tests do not pin.

#include "knaster_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_last_error;

using F = float;
using PFloat = double; // knaster_primitives/src/parameters.rs:6

constexpr F F_PI = 3.14159265358979323846264338327950288f;  // core::f32::consts::PI
constexpr F F_TAU = 6.28318530717958647692528676655900577f; // core::f32::consts::TAU

// Rust `as u32` from f64: truncating, saturating, NaN -> 0.
inline uint32_t sat_u32(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}
// Rust `as usize` from f64
inline uint64_t sat_usize(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)v;
}

// ---------------------------------------------------------------- Seconds
// knaster_primitives/src/time.rs:10,25-90,113-135
constexpr uint32_t TESIMALS = 282240000u;
struct Seconds {
    uint32_t seconds = 0;
    uint32_t sub = 0;
    static Seconds from_samples(uint64_t samples, uint64_t sr) { // time.rs:76-84
        Seconds s;
        s.seconds = (uint32_t)(samples / sr);
        s.sub = (uint32_t)((samples % sr) * (uint64_t)TESIMALS / sr);
        return s;
    }
    uint64_t to_samples(uint64_t sr) const { // time.rs:86-90
        return (uint64_t)seconds * sr + ((uint64_t)sub * sr) / (uint64_t)TESIMALS;
    }
    static Seconds from_secs_f64(double v) { // time.rs:58-63
        Seconds s;
        s.seconds = sat_u32(std::floor(v));
        double fract = v - std::trunc(v);
        s.sub = sat_u32(fract * (double)TESIMALS);
        return s;
    }
    uint64_t to_tesimals() const { return (uint64_t)seconds * TESIMALS + sub; } // time.rs:54-56
    static Seconds from_tesimals(uint64_t t) {                                   // time.rs:47-52
        Seconds s;
        s.seconds = (uint32_t)(t / TESIMALS);
        s.sub = (uint32_t)(t - (uint64_t)s.seconds * TESIMALS);
        return s;
    }
    bool operator==(const Seconds &o) const { return seconds == o.seconds && sub == o.sub; }
    bool le(const Seconds &o) const {
        return seconds < o.seconds || (seconds == o.seconds && sub <= o.sub);
    }
    Seconds saturating_sub(const Seconds &rhs) const { // time.rs:120-135
        if (le(rhs)) return Seconds{};
        Seconds r;
        if (sub >= rhs.sub) {
            r.seconds = seconds - rhs.seconds;
            r.sub = sub - rhs.sub;
        } else {
            r.seconds = seconds - rhs.seconds - 1;
            r.sub = TESIMALS - (rhs.sub - sub);
        }
        return r;
    }
};

// ---------------------------------------------------------------- parameters
// knaster_core/src/parameters/types.rs:25-37,108-124
enum class PK { Float, Trigger, Integer, Bool, Smoothing };
struct ParamValue {
    PK kind = PK::Float;
    PFloat f = 0.0;
    int64_t i = 0;
    bool b = false;
    int smoothing = 0; // 0 None, 1 Linear
    float smooth_seconds = 0.f;
    int rate = 0; // 0 BlockRate, 1 AudioRate
    static ParamValue Float(PFloat v) {
        ParamValue p;
        p.kind = PK::Float;
        p.f = v;
        return p;
    }
};

// knaster_core/src/ugen.rs:57-112
struct BlockMeta {
    size_t block_start_offset = 0;
    size_t frames_to_process = 0;
    uint64_t frame_clock = 0;
    BlockMeta make_partial(size_t start, size_t len) const { // ugen.rs:87-93
        BlockMeta m;
        m.block_start_offset = block_start_offset + start;
        m.frames_to_process = len;
        m.frame_clock = frame_clock + start;
        return m;
    }
};
// knaster_core/src/ugen.rs:8-49
struct Ctx {
    uint32_t sample_rate = 48000;
    size_t block_size = 64;
    BlockMeta block;
    uint64_t log_count = 0; // stands in for rt_log! (knaster_core/src/log.rs)
};

constexpr int MAX_CH = 16;

// knaster_core/src/ugen.rs:232-369 (trait UGen) / knaster_graph/src/dynugen.rs:23-63
struct UGen {
    virtual ~UGen() {}
    virtual int inputs() const = 0;
    virtual int outputs() const = 0;
    virtual int parameters() const = 0;
    virtual void init(uint32_t, size_t) {}
    // one frame: in[inputs], out[outputs]
    virtual void process(Ctx &ctx, const F *in, F *out) = 0;
    // block: channel pointers already offset to the (partial) block start;
    // ctx.block.frames_to_process frames.  Default = per-frame loop, ugen.rs:263-284.
    virtual void process_block(Ctx &ctx, const F *const *in, F *const *out) {
        const int ni = inputs(), no = outputs();
        F fin[MAX_CH], fout[MAX_CH];
        for (size_t fr = 0; fr < ctx.block.frames_to_process; fr++) {
            for (int i = 0; i < ni; i++) fin[i] = in[i][fr];
            process(ctx, fin, fout);
            for (int i = 0; i < no; i++) out[i][fr] = fout[i];
        }
    }
    virtual void param_apply(Ctx &ctx, size_t index, const ParamValue &v) = 0;
    virtual void set_ar_param_buffer(Ctx &ctx, size_t, const F *) { ctx.log_count++; } // ugen.rs:322-329
    virtual void set_delay_within_block_for_param(Ctx &ctx, size_t, uint16_t) {         // ugen.rs:339-341
        ctx.log_count++;
    }
};

// CRTP helper: Rust monomorphises the default process_block around an inlined
// `process`; do the same so the CPU baseline is not penalised by a virtual
// call per frame.
template <class D, int NI, int NO, int NP> struct UGenT : UGen {
    int inputs() const override { return NI; }
    int outputs() const override { return NO; }
    int parameters() const override { return NP; }
    void process(Ctx &ctx, const F *in, F *out) override { static_cast<D *>(this)->tick(ctx, in, out); }
    void process_block(Ctx &ctx, const F *const *in, F *const *out) override {
        D *d = static_cast<D *>(this);
        F fin[NI > 0 ? NI : 1], fout[NO > 0 ? NO : 1];
        for (size_t fr = 0; fr < ctx.block.frames_to_process; fr++) {
            for (int i = 0; i < NI; i++) fin[i] = in[i][fr];
            d->tick(ctx, fin, fout);
            for (int i = 0; i < NO; i++) out[i][fr] = fout[i];
        }
    }
};

// ---------------------------------------------------------------- SinWt
// knaster_core_dsp/src/dsp/wavetable.rs:8-15,130-139 ; osc.rs:90-91
constexpr uint32_t TABLE_SIZE = 16384;
constexpr uint32_t TABLE_HIGH_MASK = TABLE_SIZE - 1;
constexpr uint32_t FRACTIONAL_PART = 65536;
const std::vector<float> &sine_table() {
    static std::vector<float> t = [] {
        std::vector<float> v(TABLE_SIZE);
        for (uint32_t i = 0; i < TABLE_SIZE; i++)
            v[i] = (float)std::sin(((double)i / (double)TABLE_SIZE) * M_PI * 2.0); // wavetable.rs:134-136
        return v;
    }();
    return t;
}

// knaster_core_dsp/src/ugens/osc.rs:97-168
struct SinWt : UGenT<SinWt, 0, 1, 3> {
    uint32_t phase = 0, phase_offset = 0, phase_increment = 0;
    double freq_to_phase_inc = 0.0;
    F freq;
    const float *table;
    explicit SinWt(F f) : freq(f), table(sine_table().data()) {} // osc.rs:110-123
    void set_freq(PFloat f) {                                    // osc.rs:127-130
        freq = (F)f;
        phase_increment = sat_u32((double)freq * freq_to_phase_inc);
    }
    void init(uint32_t sr, size_t) override { // osc.rs:142-147
        phase = 0;
        freq_to_phase_inc = (double)TABLE_SIZE * (double)FRACTIONAL_PART * (1.0 / (double)sr);
        set_freq((double)freq);
    }
    inline F next_sample() { // osc.rs:151-156 ; wavetable.rs:27-32,50-52,322-324
        uint32_t p = phase + phase_offset;
        F s = table[(p >> 16) & TABLE_HIGH_MASK];
        phase += phase_increment;
        return s;
    }
    inline void tick(Ctx &, const F *, F *out) { out[0] = next_sample(); }
    void process_block(Ctx &ctx, const F *const *, F *const *out) override { // osc.rs:162-167
        for (size_t i = 0; i < ctx.block.frames_to_process; i++) out[0][i] = next_sample();
    }
    void param_apply(Ctx &, size_t index, const ParamValue &v) override {
        switch (index) {
        case 0: if (v.kind == PK::Float) set_freq(v.f); break;
        case 1: if (v.kind == PK::Float) phase_offset = sat_u32(v.f * (double)FRACTIONAL_PART); break; // osc.rs:133-135
        case 2: phase = 0; break; // osc.rs:138-140
        default: break;
        }
    }
};

// knaster_core_dsp/src/ugens/osc.rs:222-271
struct SinNumeric : UGenT<SinNumeric, 0, 1, 3> {
    F phase, phase_offset = 0.f, phase_increment = 0.f;
    explicit SinNumeric(F f) : phase(f) {} // osc.rs:231-237 (freq stashed in phase)
    void init(uint32_t sr, size_t) override { // osc.rs:253-261
        if (phase_increment == 0.f) phase_increment = phase / (F)(float)sr;
        phase = 0.f;
    }
    inline void tick(Ctx &, const F *, F *out) { // osc.rs:263-270
        F o = sinf((phase + phase_offset) * F_TAU);
        phase += phase_increment;
        if (phase > 1.0f) phase -= 1.0f;
        out[0] = o;
    }
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override {
        switch (index) {
        case 0: if (v.kind == PK::Float) phase_increment = (F)v.f / (F)(float)ctx.sample_rate; break; // osc.rs:240-242
        case 1: if (v.kind == PK::Float) phase_offset = (F)v.f; break;
        case 2: phase = 0.f; break;
        default: break;
        }
    }
};

// ---------------------------------------------------------------- PolyBlep
// knaster_core_dsp/src/ugens/polyblep.rs
inline F sq(F x) { return x * x; } // polyblep.rs:39-41
inline F blep(F t, F dt) {         // polyblep.rs:47-55
    if (t < dt) return -sq(t / dt - 1.0f);
    else if (t > 1.0f - dt) return sq((t - 1.0f) / dt + 1.0f);
    else return 0.0f;
}
inline F blamp(F t, F dt) { // polyblep.rs:58-68
    if (t < dt) {
        t = t / dt - 1.0f;
        return -1.0f / 3.0f * sq(t) * t;
    } else if (t > 1.0f - dt) {
        t = (t - 1.0f) / dt + 1.0f;
        return 1.0f / 3.0f * sq(t) * t;
    } else return 0.0f;
}
inline F boz(F t) { return truncf(t); } // polyblep.rs:70-72

struct PolyBlep : UGenT<PolyBlep, 0, 1, 3> {
    // polyblep.rs:90-120 enum Waveform
    enum { Sawtooth = 0, Sine, Cosine, Triangle, Square, Rectangle, Ramp, ModifiedTriangle, ModifiedSquare,
           HalfWaveRectifiedSine, FullWaveRectifiedSine, TriangularPulse, TrapezoidFixed, TrapezoidVariable };
    int waveform;
    F sample_rate = 0.f, freq_in_hz, dt = 0.f, pulse_width = 0.5f, t = 0.f; // polyblep.rs:139-148
    PolyBlep(int wf, F f) : waveform(wf), freq_in_hz(f) {}
    void set_freq(F f) { // polyblep.rs:181-184
        freq_in_hz = f;
        dt = f / sample_rate;
    }
    void init(uint32_t sr, size_t) override { // polyblep.rs:150-155
        sample_rate = (F)sr;
        if (dt == 0.f && freq_in_hz != 0.f) set_freq(freq_in_hz);
    }
    inline F get_freq_in_hz() const { return dt * sample_rate; } // polyblep.rs:194-196
    inline F w_sin() { return sinf(t * F_TAU); }                 // polyblep.rs:243-245
    inline F w_cos() { return cosf(t * F_TAU); }                 // polyblep.rs:247-249
    inline F w_saw() {                                           // polyblep.rs:490-498
        F _t = t + 0.5f;
        _t -= boz(_t);
        F y = 2.0f * _t - 1.0f;
        y -= blep(_t, dt);
        return y;
    }
    inline F w_ramp() { // polyblep.rs:500-508
        F _t = t;
        _t -= boz(_t);
        F y = 1.0f - 2.0f * _t;
        y += blep(_t, dt);
        return y;
    }
    inline F w_sqr() { // polyblep.rs:434-447
        F t2 = t + 0.5f;
        t2 -= boz(t2);
        F y = t < 0.5f ? 1.0f : -1.0f;
        y += blep(t, dt) - blep(t2, dt);
        return y;
    }
    inline F w_rect() { // polyblep.rs:475-488
        F t2 = t + 1.0f - pulse_width;
        t2 -= boz(t2);
        F y = -2.0f * pulse_width;
        if (t < pulse_width) y += 2.0f;
        y += blep(t, dt) - blep(t2, dt);
        return y;
    }
    inline F w_tri() { // polyblep.rs:278-299
        F t1 = t + 0.25f;
        t1 -= boz(t1);
        F t2 = t + 0.75f;
        t2 -= boz(t2);
        F y = t * 4.0f;
        if (y >= 3.0f) y -= 4.0f;
        else if (y > 1.0f) y = 2.0f - y;
        y += 4.0f * dt * (blamp(t1, dt) - blamp(t2, dt));
        return y;
    }
    inline F fold4() { // the 4t fold shared by tri / trap / trap2, polyblep.rs:286-292
        F y = t * 4.0f;
        if (y >= 3.0f) y -= 4.0f;
        else if (y > 1.0f) y = 2.0f - y;
        return y;
    }
    static inline F clamp1(F x) { return x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x); } // f32::clamp(-1, 1)
    inline F w_half() { // polyblep.rs:251-266
        F t2 = t + 0.5f;
        t2 -= boz(t2);
        F y = t < 0.5f ? 2.0f * sinf(t * F_TAU) - 2.0f / F_PI : -2.0f / F_PI;
        y += F_TAU * dt * (blamp(t, dt) + blamp(t2, dt));
        return y;
    }
    inline F w_full() { // polyblep.rs:268-276
        F _t = t + 0.25f;
        _t -= boz(_t);
        F y = 2.0f * sinf(_t * F_PI) - 4.0f / F_PI;
        y += F_TAU * dt * blamp(_t, dt);
        return y;
    }
    inline F w_tri2() { // polyblep.rs:301-324
        F pw = pulse_width;
        if (pw > 0.9999f) pw = 0.9999f; // .min(0.9999).max(0.0001)
        if (pw < 0.0001f) pw = 0.0001f;
        F t1 = t + 0.5f * pw;
        t1 -= boz(t1);
        F t2 = t + 1.0f - 0.5f * pw;
        t2 -= boz(t2);
        F y = t * 2.0f;
        if (y >= 2.0f - pw) y = (y - 2.0f) / pw;
        else if (y >= pw) y = 1.0f - (y - pw) / (1.0f - pw);
        else y /= pw;
        y += dt / (pw - pw * pw) * (blamp(t1, dt) - blamp(t2, dt));
        return y;
    }
    inline F w_trip() { // polyblep.rs:326-353
        const F pw = pulse_width;
        F t1 = t + 0.75f + 0.5f * pw;
        t1 -= boz(t1);
        F y;
        if (t1 >= pw) {
            y = -pw;
        } else {
            y = 4.0f * t1;
            y = y >= 2.0f * pw ? 4.0f - y / pw - pw : y / pw - pw;
        }
        if (pw > 0.0f) {
            F t2 = t1 + 1.0f - 0.5f * pw;
            t2 -= boz(t2);
            F t3 = t1 + 1.0f - pw;
            t3 -= boz(t3);
            y += 2.0f * dt / pw * (blamp(t1, dt) - 2.0f * blamp(t2, dt) + blamp(t3, dt));
        }
        return y;
    }
    inline F w_trap() { // polyblep.rs:355-388
        F y = clamp1(2.0f * fold4());
        F t1 = t + 0.125f;
        t1 -= boz(t1);
        F t2 = t1 + 0.5f;
        t2 -= boz(t2);
        y += 4.0f * dt * (blamp(t1, dt) - blamp(t2, dt)); // triangle #1
        t1 = t + 0.375f;
        t1 -= boz(t1);
        t2 = t1 + 0.5f;
        t2 -= boz(t2);
        y += 4.0f * dt * (blamp(t1, dt) - blamp(t2, dt)); // triangle #2
        return y;
    }
    inline F w_trap2() { // polyblep.rs:390-432
        F pw = pulse_width < 0.9999f ? pulse_width : 0.9999f;
        const F scale = 1.0f / (1.0f - pw);
        F y = clamp1(scale * fold4());
        F t1 = t + 0.25f - 0.25f * pw;
        t1 -= boz(t1);
        F t2 = t1 + 0.5f;
        t2 -= boz(t2);
        y += scale * 2.0f * dt * (blamp(t1, dt) - blamp(t2, dt));
        t1 = t + 0.25f + 0.25f * pw;
        t1 -= boz(t1);
        t2 = t1 + 0.5f;
        t2 -= boz(t2);
        y += scale * 2.0f * dt * (blamp(t1, dt) - blamp(t2, dt));
        return y;
    }
    inline F w_sqr2() { // polyblep.rs:449-473
        F t1 = t + 0.875f + 0.25f * (pulse_width - 0.5f);
        t1 -= boz(t1);
        F t2 = t + 0.375f + 0.25f * (pulse_width - 0.5f);
        t2 -= boz(t2);
        F y = t1 < 0.5f ? 1.0f : -1.0f; // square #1
        y += blep(t1, dt) - blep(t2, dt);
        t1 += 0.5f * (1.0f - pulse_width);
        t1 -= boz(t1);
        t2 += 0.5f * (1.0f - pulse_width);
        t2 -= boz(t2);
        y += t1 < 0.5f ? 1.0f : -1.0f; // square #2
        y += blep(t1, dt) - blep(t2, dt);
        return 0.5f * y;
    }
    inline F next_sample() { // polyblep.rs:209-230
        if (get_freq_in_hz() >= sample_rate / 4.0f) return w_sin();
        switch (waveform) {
        case Sine: return w_sin();
        case Cosine: return w_cos();
        case Triangle: return w_tri();
        case Square: return w_sqr();
        case Rectangle: return w_rect();
        case Ramp: return w_ramp();
        case ModifiedTriangle: return w_tri2();
        case ModifiedSquare: return w_sqr2();
        case HalfWaveRectifiedSine: return w_half();
        case FullWaveRectifiedSine: return w_full();
        case TriangularPulse: return w_trip();
        case TrapezoidFixed: return w_trap();
        case TrapezoidVariable: return w_trap2();
        default: return w_saw();
        }
    }
    inline void tick(Ctx &, const F *, F *out) { // polyblep.rs:157-159,232-241
        F s = next_sample();
        t += dt;
        t -= boz(t);
        out[0] = s;
    }
    void param_apply(Ctx &, size_t index, const ParamValue &v) override {
        switch (index) {
        case 0: if (v.kind == PK::Float) set_freq((F)v.f); break;          // polyblep.rs:162-165
        case 1: if (v.kind == PK::Float) pulse_width = (F)v.f; break;      // polyblep.rs:167-170
        case 2: if (v.kind == PK::Integer) waveform = (int)v.i; break;     // polyblep.rs:172-175
        default: break;
        }
    }
};

// ---------------------------------------------------------------- SvfFilter
// knaster_core_dsp/src/ugens/svf.rs
struct SvfFilter : UGenT<SvfFilter, 1, 1, 5> {
    enum { Low = 0, High, Band, Notch, Peak, All, Bell, LowShelf, HighShelf }; // svf.rs:19-39
    int ty;
    F cutoff, q, gain_db;
    F ic1eq = 0.f, ic2eq = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f;
    SvfFilter(int ty_, F c, F q_, F g) : ty(ty_), cutoff(c), q(q_), gain_db(g) {} // svf.rs:64-79
    void set_coeffs(F cutoff_, F q_, F gain, F sr) { // svf.rs:146-242
        F g, k;
        switch (ty) {
        case Bell: { // svf.rs:208-218
            F amp = powf(10.0f, gain / 40.0f);
            g = tanf((F_PI * cutoff_) / sr) / sqrtf(amp);
            k = 1.0f / (q_ * amp);
            a1 = 1.0f / (1.0f + g * (g + k)); a2 = g * a1; a3 = g * a2;
            m0 = 1.0f; m1 = k * (amp * amp - 1.0f); m2 = 0.0f;
            return;
        }
        case LowShelf: { // svf.rs:219-229
            F amp = powf(10.0f, gain / 40.0f);
            g = tanf((F_PI * cutoff_) / sr) / sqrtf(amp);
            k = 1.0f / q_;
            a1 = 1.0f / (1.0f + g * (g + k)); a2 = g * a1; a3 = g * a2;
            m0 = 1.0f; m1 = k * (amp - 1.0f); m2 = amp * amp - 1.0f;
            return;
        }
        case HighShelf: { // svf.rs:230-240
            F amp = powf(10.0f, gain / 40.0f);
            g = tanf((F_PI * cutoff_) / sr) * sqrtf(amp);
            k = 1.0f / q_;
            a1 = 1.0f / (1.0f + g * (g + k)); a2 = g * a1; a3 = g * a2;
            m0 = amp * amp; m1 = k * (1.0f - amp) * amp; m2 = 1.0f - amp * amp;
            return;
        }
        default: break;
        }
        g = tanf((F_PI * cutoff_) / sr); // svf.rs:149 etc.
        k = 1.0f / q_;
        a1 = 1.0f / (1.0f + g * (g + k));
        a2 = g * a1;
        a3 = g * a2;
        switch (ty) {
        case Low: m0 = 0.f; m1 = 0.f; m2 = 1.f; break;          // svf.rs:154-156
        case Band: m0 = 0.f; m1 = 1.f; m2 = 0.f; break;         // svf.rs:164-166
        case High: m0 = 1.f; m1 = -k; m2 = -1.f; break;         // svf.rs:174-176
        case Notch: m0 = 1.f; m1 = -k; m2 = 0.f; break;         // svf.rs:184-186
        case Peak: m0 = 1.f; m1 = -k; m2 = -2.0f; break;        // svf.rs:194-196
        case All: m0 = 1.f; m1 = -2.0f * k; m2 = 0.f; break;    // svf.rs:204-206
        default: break;
        }
    }
    void init(uint32_t sr, size_t) override { set_coeffs(cutoff, q, gain_db, (F)(float)sr); } // svf.rs:134-141
    inline void tick(Ctx &, const F *in, F *out) { // svf.rs:245-280
        F v0 = in[0];
        F v3 = v0 - ic2eq;
        F v1 = a1 * ic1eq + a2 * v3;
        F v2 = ic2eq + a2 * ic1eq + a3 * v3;
        ic1eq = 2.0f * v1 - ic1eq;
        ic2eq = 2.0f * v2 - ic2eq;
        out[0] = m0 * v0 + m1 * v1 + m2 * v2;
    }
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override { // svf.rs:81-133
        F sr = (F)(float)ctx.sample_rate;
        switch (index) {
        case 0: if (v.kind != PK::Float) return; cutoff = (F)v.f; break;
        case 1: if (v.kind != PK::Float) return; q = (F)v.f; break;
        case 2: if (v.kind != PK::Float) return; gain_db = (F)v.f; break;
        case 3: if (v.kind != PK::Integer) return; ty = (int)v.i; break;
        case 4: break;
        default: return;
        }
        set_coeffs(cutoff, q, gain_db, sr);
    }
};

// ---------------------------------------------------------------- OnePole
// knaster_core_dsp/src/ugens/onepole.rs:13-92
struct OnePole {
    F last_output = 0.f, a0 = 1.f, b1 = 0.f;
    void set_freq_lowpass(F freq, F sr) { // onepole.rs:35-46
        F f = freq / sr;
        F b_tmp = expf(-2.0f * F_PI * f);
        b1 = b_tmp;
        a0 = 1.0f - b1;
    }
    void set_freq_highpass(F freq, F sr) { set_freq_lowpass(freq, sr); } // onepole.rs:50-60
    inline F process_lp(F x) {                                           // onepole.rs:64-77
        last_output = x * a0 + last_output * b1;
        return last_output;
    }
    inline F process_hp(F x) { // onepole.rs:80-92
        last_output = x * a0 + last_output * b1;
        return x - last_output;
    }
};
struct OnePoleLpf : UGenT<OnePoleLpf, 1, 1, 1> { // onepole.rs:111-140
    OnePole op;
    explicit OnePoleLpf(F cutoff) { op.b1 = cutoff; }
    void init(uint32_t sr, size_t) override {
        if (op.a0 == 1.0f) op.set_freq_lowpass(op.b1, (F)(float)sr);
    }
    inline void tick(Ctx &, const F *in, F *out) { out[0] = op.process_lp(in[0]); }
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override {
        if (index == 0 && v.kind == PK::Float) op.set_freq_lowpass((F)v.f, (F)ctx.sample_rate);
    }
};
struct OnePoleHpf : UGenT<OnePoleHpf, 1, 1, 1> { // onepole.rs:144-177
    OnePole op;
    OnePoleHpf() {}
    void init(uint32_t sr, size_t) override {
        if (op.a0 == 1.0f) op.set_freq_highpass(op.b1, (F)(float)sr);
    }
    inline void tick(Ctx &, const F *in, F *out) { out[0] = op.process_hp(in[0]); }
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override {
        if (index == 0 && v.kind == PK::Float) op.set_freq_highpass((F)v.f, (F)ctx.sample_rate);
    }
};

// ---------------------------------------------------------------- EnvAsr / EnvAr
// knaster_core_dsp/src/ugens/envelopes.rs:19-163
struct EnvAsr : UGenT<EnvAsr, 0, 1, 4> {
    enum { Stopped, Attacking, Sustaining, Releasing };
    int state = Stopped;
    F t = 0.f, attack_seconds, attack_rate = 1.f, release_seconds, release_rate = 1.f, release_scale = 1.f;
    EnvAsr(F a, F r) : attack_seconds(a), release_seconds(r) {} // envelopes.rs:33-43
    void init(uint32_t sr, size_t) override {                   // envelopes.rs:135-151
        if (attack_rate == 1.0f) {
            if (attack_seconds == 0.f) attack_rate = 1.0f;
            else attack_rate = 1.0f / (attack_seconds * (F)sr);
        }
        if (release_rate == 1.0f) {
            if (release_seconds == 0.f) release_rate = 1.0f;
            else release_rate = 1.0f / (release_seconds * (F)sr);
        }
    }
    inline F next_sample() { // envelopes.rs:52-81
        F out;
        switch (state) {
        case Attacking:
            out = t;
            t += attack_rate;
            if (t >= 1.0f) state = Sustaining;
            break;
        case Sustaining: out = 1.0f; break;
        case Releasing:
            out = (t * t * t) * release_scale; // powi(3): (t*t)*t, see SURVEY 8c
            t -= release_rate;
            if (t <= 0.f) {
                state = Stopped;
                t = 0.f;
            }
            break;
        default: out = 0.f; break;
        }
        return out;
    }
    inline void tick(Ctx &, const F *, F *out) { out[0] = next_sample(); }
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override {
        switch (index) {
        case 0: { // envelopes.rs:84-96
            if (v.kind != PK::Float) return;
            F atk = (F)v.f;
            if (attack_seconds != atk) {
                attack_seconds = atk;
                if (atk == 0.f) attack_rate = 1.0f;
                else attack_rate = 1.0f / (attack_seconds * (F)ctx.sample_rate);
            }
            break;
        }
        case 1: { // envelopes.rs:98-110
            if (v.kind != PK::Float) return;
            F rel = (F)v.f;
            if (release_seconds != rel) {
                release_seconds = rel;
                if (rel == 0.f) release_rate = 1.0f;
                else release_rate = 1.0f / (release_seconds * (F)ctx.sample_rate);
            }
            break;
        }
        case 2: // t_release envelopes.rs:112-128
            if (state == Attacking) {
                release_scale = t;
                state = Releasing;
                t = 1.0f;
            } else if (state == Sustaining) {
                release_scale = 1.0f;
                state = Releasing;
                t = 1.0f;
            }
            break;
        case 3: state = Attacking; break; // t_restart envelopes.rs:130-133,47-49
        default: break;
        }
    }
};
// envelopes.rs:174-303
struct EnvAr : UGenT<EnvAr, 0, 1, 3> {
    enum { Stopped, Attacking, Releasing };
    int state = Stopped;
    F t = 0.f, attack_seconds, attack_rate = 1.f, release_seconds, release_rate = 1.f, release_scale = 1.f;
    EnvAr(F a, F r) : attack_seconds(a), release_seconds(r) {}
    void init(uint32_t sr, size_t) override { // envelopes.rs:268-284
        if (attack_rate == 1.0f) {
            if (attack_seconds == 0.f) attack_rate = 1.0f;
            else attack_rate = 1.0f / (attack_seconds * (F)sr);
        }
        if (release_rate == 1.0f) {
            if (release_seconds == 0.f) release_rate = 1.0f;
            else release_rate = 1.0f / (release_seconds * (F)sr);
        }
    }
    inline void tick(Ctx &, const F *, F *out) { // envelopes.rs:205-233
        F o;
        switch (state) {
        case Attacking:
            o = t;
            t += attack_rate;
            if (t >= 1.0f) {
                release_scale = 1.0f;
                state = Releasing;
                t = 1.0f;
            }
            break;
        case Releasing:
            o = (t * t * t) * release_scale;
            t -= release_rate;
            if (t <= 0.f) {
                state = Stopped;
                t = 0.f;
            }
            break;
        default: o = 0.f; break;
        }
        out[0] = o;
    }
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override { // envelopes.rs:234-266
        switch (index) {
        case 0: {
            if (v.kind != PK::Float) return;
            F atk = (F)v.f;
            if (attack_seconds != atk) {
                attack_seconds = atk;
                if (atk == 0.f) attack_rate = 1.0f;
                else attack_rate = 1.0f / (attack_seconds * (F)ctx.sample_rate);
            }
            break;
        }
        case 1: {
            if (v.kind != PK::Float) return;
            F rel = (F)v.f;
            if (release_seconds != rel) {
                release_seconds = rel;
                if (rel == 0.f) release_rate = 1.0f;
                else release_rate = 1.0f / (release_seconds * (F)ctx.sample_rate);
            }
            break;
        }
        case 2: state = Attacking; break;
        default: break;
        }
    }
};

// ---------------------------------------------------------------- Envelope
// knaster_core_dsp/src/ugens/envelopes.rs:322-527
struct EnvelopeSegment {
    double reciprocal_duration, duration, value;
    EnvelopeSegment(double d, double v) : reciprocal_duration(1.0 / d), duration(d), value(v) {} // :329-335
};
struct Envelope : UGenT<Envelope, 0, 1, 4> {
    bool running = false;
    size_t current_segment = 0;
    double current_time = 0.0;
    std::vector<EnvelopeSegment> segments;
    double start_value, from_value, time_scale = 1.0, base_scale = 0.0;
    bool looping = false;
    // new() :372-384, then the builders time_scale() :386-389 and looping() :392-395
    Envelope(double start, std::vector<EnvelopeSegment> segs, bool loop, double ts)
        : segments(std::move(segs)), start_value(start), from_value(start), time_scale(ts), looping(loop) {}
    void init(uint32_t sr, size_t) override { base_scale = 1.0 / (double)sr; } // :403-405
    inline void tick(Ctx &, const F *, F *out) {                               // :407-463
        F o;
        if (!running) {
            o = (F)from_value;
        } else {
            double t = current_time;
            size_t cs = current_segment;
            if (t < segments[cs].duration) {
                const EnvelopeSegment &s = segments[cs];
                o = (F)(from_value + (t * s.reciprocal_duration) * (s.value - from_value));
                current_time = t + (time_scale * base_scale);
            } else if (cs + 1 < segments.size()) {
                from_value = segments[cs].value;
                const EnvelopeSegment &s = segments[cs];
                o = (F)(from_value + (t * s.reciprocal_duration) * (s.value - from_value));
                current_segment = cs + 1;
                current_time = current_time - s.duration + (time_scale * base_scale);
            } else {
                from_value = segments[cs].value;
                o = (F)from_value;
                if (looping) {
                    current_segment = 0;
                    current_time = 0.0;
                } else {
                    running = false;
                }
            }
        }
        out[0] = o;
    }
    void param_apply(Ctx &, size_t index, const ParamValue &v) override { // :476-526
        switch (index) {
        case 0:
            if (v.kind != PK::Float) return;
            time_scale = (double)(F)v.f; // F::new(value).to_f64()
            break;
        case 1: {
            if (v.kind != PK::Integer) return;
            size_t j = (size_t)v.i;
            if (j >= segments.size()) j = segments.size() - 1;
            running = true;
            current_segment = j;
            current_time = 0.0;
            break;
        }
        case 2:
            running = true;
            current_segment = 0;
            current_time = 0.0;
            from_value = start_value;
            break;
        case 3:
            if (running) {
                const EnvelopeSegment &s = segments[current_segment];
                from_value = from_value + (current_time * s.reciprocal_duration) * (s.value - from_value);
            }
            running = false;
            break;
        default: break;
        }
    }
};

// ---------------------------------------------------------------- Math / Constant / fixtures
// knaster_core_dsp/src/ugens/math.rs:17-165
struct MathUGen : UGen {
    int n, op;
    MathUGen(int channels, int op_) : n(channels), op(op_) {}
    int inputs() const override { return 2 * n; }
    int outputs() const override { return n; }
    int parameters() const override { return 0; }
    static inline F apply1(int op, F a, F b) {
        switch (op) {
        case KO_OP_ADD: return a + b;
        case KO_OP_SUB: return a - b;
        case KO_OP_MUL: return a * b;
        case KO_OP_DIV: return a / b;
        default: return powf(a, b);
        }
    }
    void process(Ctx &, const F *in, F *out) override { // math.rs:112-137
        for (int c = 0; c < n; c++) out[c] = apply1(op, in[c], in[c + n]);
    }
    void process_block(Ctx &ctx, const F *const *in, F *const *out) override { // math.rs:138-156
        const size_t fr = ctx.block.frames_to_process;
        for (int c = 0; c < n; c++) {
            const F *a = in[c], *b = in[c + n];
            F *o = out[c];
            switch (op) { // separate loops so each auto-vectorises like the Rust version (math.rs:28-30)
            case KO_OP_ADD: for (size_t i = 0; i < fr; i++) o[i] = a[i] + b[i]; break;
            case KO_OP_SUB: for (size_t i = 0; i < fr; i++) o[i] = a[i] - b[i]; break;
            case KO_OP_MUL: for (size_t i = 0; i < fr; i++) o[i] = a[i] * b[i]; break;
            case KO_OP_DIV: for (size_t i = 0; i < fr; i++) o[i] = a[i] / b[i]; break;
            default: for (size_t i = 0; i < fr; i++) o[i] = powf(a[i], b[i]); break;
            }
        }
    }
    void param_apply(Ctx &, size_t, const ParamValue &) override {}
};
// knaster_core_dsp/src/ugens/util.rs:37-64
// Math1UGen: math.rs:167-305 (Operation1 = Ceil / Sqrt / Floor / Trunc / Fract / Exp)
struct Math1UGen : UGenT<Math1UGen, 1, 1, 0> {
    int op;
    explicit Math1UGen(int op_) : op(op_) {}
    inline void tick(Ctx &, const F *in, F *out) {
        const F a = in[0];
        switch (op) {
        case 0: out[0] = ceilf(a); break;        // math.rs:172-183
        case 1: out[0] = sqrtf(a); break;        // :184-195
        case 2: out[0] = floorf(a); break;       // :196-207
        case 3: out[0] = truncf(a); break;       // :208-219
        case 4: out[0] = a - truncf(a); break;   // :220-231  f32::fract = self - self.trunc()
        default: out[0] = expf(a); break;        // :232-243
        }
    }
    void param_apply(Ctx &, size_t, const ParamValue &) override {}
};

// Phasor: osc.rs:170-213.  All f64; the output is F::new(phase).
struct Phasor : UGenT<Phasor, 0, 1, 1> {
    double phase = 0.0, step, mult = 0.0;
    explicit Phasor(double freq) : step(freq) {}
    void set_freq(double f) { step = mult == 0.0 ? f : f * mult; } // osc.rs:191-197
    void init(uint32_t sr, size_t) override {                      // osc.rs:199-202
        mult = 1.0 / (double)sr;
        set_freq(step);
    }
    inline void tick(Ctx &, const F *, F *out) {                   // osc.rs:204-212
        out[0] = (F)phase;
        phase += step;
        while (phase >= 1.0) phase -= 1.0;
    }
    void param_apply(Ctx &, size_t index, const ParamValue &v) override {
        if (index == 0 && v.kind == PK::Float) set_freq(v.f);
    }
};

// ---------------------------------------------------------------- noise
// knaster_core_dsp/src/ugens/noise.rs.  The generator is the `fastrand` crate, pinned at 2.3.0
// (Cargo.lock:937-938; knaster_core_dsp/Cargo.toml:34), which is NOT vendored under the reference:
// PARITY UNPINNED for the random stream -- wyrand restated from fastrand's published source:
//   Rng::with_seed(seed) = Rng(seed);
//   gen_u64: s = state + 0x2d358dccaa6c78a5 (wrapping), state = s,
//            t = u128(s) * u128(s ^ 0x8bb84b93962eacc9), return lo(t) ^ hi(t);
//   u32(..) = gen_u64() as u32;  f32() = from_bits(0x3F800000 + (u32(..) >> 9)) - 1.0.
// Everything downstream of rng.f32() follows noise.rs line by line and is pinned like the other UGens.
struct FastRand {
    uint64_t state;
    explicit FastRand(uint64_t seed) : state(seed) {}
    uint64_t gen_u64() {
        state += 0x2d358dccaa6c78a5ull;
        const unsigned __int128 t = (unsigned __int128)state * (unsigned __int128)(state ^ 0x8bb84b93962eacc9ull);
        return (uint64_t)t ^ (uint64_t)(t >> 64);
    }
    float f32() {
        uint32_t bits = 0x3F800000u + ((uint32_t)gen_u64() >> 9);
        float f;
        std::memcpy(&f, &bits, 4);
        return f - 1.0f;
    }
};
// noise.rs:26-46.  The seed is the value next_randomness_seed() returned at construction (:11-22).
struct WhiteNoise : UGenT<WhiteNoise, 0, 1, 0> {
    FastRand rng;
    explicit WhiteNoise(uint64_t seed) : rng(seed) {}
    inline void tick(Ctx &, const F *, F *out) { out[0] = (F)(rng.f32() * 2.0f - 1.0f); } // :40-42
    void param_apply(Ctx &, size_t, const ParamValue &) override {}
};
// noise.rs:53-115 (Voss-McCartney)
struct PinkNoise : UGenT<PinkNoise, 0, 1, 0> {
    static constexpr uint32_t OCTAVES = 9;
    F white_noises[OCTAVES] = {0};
    F always_on_white_noise = 0;
    uint32_t counter = 1, mask = 256; // 2^(OCTAVES - 1)
    F pink = 0;
    FastRand rng;
    explicit PinkNoise(uint64_t seed) : rng(seed) {}
    inline void tick(Ctx &, const F *, F *out) {                  // :94-114
        const size_t index = (size_t)__builtin_ctz(counter);      // counter.trailing_zeros()
        pink -= white_noises[index];
        white_noises[index] = (F)(rng.f32() * 2.0f - 1.0f);
        pink += white_noises[index];
        pink -= always_on_white_noise;
        always_on_white_noise = (F)(rng.f32() * 2.0f - 1.0f);
        pink += always_on_white_noise;
        counter &= mask - 1;                                      // increment_counter :85-91
        counter += 1;
        out[0] = pink / ((F)OCTAVES + (F)1);
    }
    void param_apply(Ctx &, size_t, const ParamValue &) override {}
};
// noise.rs:122-153
struct BrownNoise : UGenT<BrownNoise, 0, 1, 0> {
    FastRand rng;
    F last_output = 0;
    explicit BrownNoise(uint64_t seed) : rng(seed) {}
    inline void tick(Ctx &, const F *, F *out) {                  // :141-149
        const F white = (F)(rng.f32() * 2.0f - 1.0f);
        last_output += white * (F)0.1;
        last_output = std::fmin(std::fmax(last_output, (F)-1), (F)1); // f32::clamp
        out[0] = last_output;
    }
    void param_apply(Ctx &, size_t, const ParamValue &) override {}
};
// noise.rs:156-217
struct RandomLin : UGenT<RandomLin, 0, 1, 1> {
    FastRand rng;
    F current_value, current_change_width = 0, phase = 0, phase_step, freq_to_phase_inc = 0;
    RandomLin(F freq, uint64_t seed) : rng(seed), phase_step(freq) { current_value = (F)rng.f32(); } // :172-182
    void new_value() {                                             // :185-191
        const F old_target = current_value + current_change_width;
        const F nv = (F)rng.f32();
        current_value = old_target;
        current_change_width = nv - old_target;
        phase = (F)0.0;
    }
    void init(uint32_t sr, size_t) override {                      // :192-197
        freq_to_phase_inc = (F)1 / (F)sr;
        phase_step *= freq_to_phase_inc;
        new_value();
    }
    inline void tick(Ctx &, const F *, F *out) {                   // :198-206
        out[0] = current_value + phase * current_change_width;
        phase += phase_step;
        if (phase >= (F)1) new_value();
    }
    void param_apply(Ctx &, size_t index, const ParamValue &v) override { // :208-216
        if (index == 0 && v.kind == PK::Float) {
            if (freq_to_phase_inc == (F)0) phase_step = (F)v.f;
            else phase_step = (F)v.f * freq_to_phase_inc;
        }
    }
};

// ---------------------------------------------------------------- Pan2
// knaster_core_dsp/src/ugens/pan.rs:12-38.  The gains come from the `fastapprox` crate, pinned at 0.3.1
// (knaster_core_dsp/Cargo.toml:35) and NOT vendored under the reference: PARITY UNPINNED for fast::sin /
// fast::cos, restated here from the crate's published source (a port of Paul Mineiro's fastsin / fastcos).
namespace fastapprox_fast {
inline uint32_t to_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float from_bits(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float sin(float x) {
    const float FOUROVERPI = 1.2732395447351627f, FOUROVERPISQ = 0.40528473456935109f, Q = 0.78444488374548933f;
    uint32_t p = to_bits(0.20363937680730309f), r = to_bits(0.015124940802184233f), s = to_bits(-0.0032225901625579573f);
    uint32_t v = to_bits(x);
    const uint32_t sign = v & 0x80000000u;
    v &= 0x7FFFFFFFu;
    const float qpprox = FOUROVERPI * x - FOUROVERPISQ * x * from_bits(v);
    const float qpproxsq = qpprox * qpprox;
    p |= sign;
    r |= sign;
    s ^= sign;
    return Q * qpprox + qpproxsq * (from_bits(p) + qpproxsq * (from_bits(r) + qpproxsq * from_bits(s)));
}
inline float cos(float x) {
    const float HALFPI = 1.5707963267948966f, HALFPIMINUSTWOPI = -4.7123889803846899f;
    const float offset = x > HALFPI ? HALFPIMINUSTWOPI : HALFPI;
    return sin(x + offset);
}
} // namespace fastapprox_fast
struct Pan2 : UGenT<Pan2, 1, 2, 1> {
    float pan;
    explicit Pan2(float p) : pan(p * 0.5f + 0.5f) {}                          // :19-24
    inline void tick(Ctx &, const F *in, F *out) {                            // :31-36
        const F signal = in[0];
        const float pan_pos_radians = pan * 1.57079632679489661923f;          // core::f32::consts::FRAC_PI_2
        const F left_gain = (F)fastapprox_fast::cos(pan_pos_radians);
        const F right_gain = (F)fastapprox_fast::sin(pan_pos_radians);
        out[0] = signal * left_gain;
        out[1] = signal * right_gain;
    }
    void param_apply(Ctx &, size_t index, const ParamValue &v) override {    // :26-29
        if (index == 0 && v.kind == PK::Float) pan = (float)v.f * 0.5f + 0.5f;
    }
};

struct Constant : UGenT<Constant, 0, 1, 1> {
    F value;
    explicit Constant(F v) : value(v) {}
    inline void tick(Ctx &, const F *, F *out) { out[0] = value; }
    void process_block(Ctx &ctx, const F *const *, F *const *out) override {
        std::fill(out[0], out[0] + ctx.block.frames_to_process, value);
    }
    void param_apply(Ctx &, size_t index, const ParamValue &v) override {
        if (index == 0 && v.kind == PK::Float) value = (F)v.f;
    }
};
// knaster_core_dsp/src/test_utils.rs:8-41 ; knaster_graph/src/tests/utils.rs:4-17
struct TestNumUGen : UGenT<TestNumUGen, 0, 1, 0> {
    F number;
    explicit TestNumUGen(F n) : number(n) {}
    inline void tick(Ctx &, const F *, F *out) { out[0] = number; }
    void param_apply(Ctx &, size_t, const ParamValue &) override {}
};
// knaster_core_dsp/src/test_utils.rs:44-86 ; knaster_graph/src/tests/utils.rs:20-67
struct TestInPlusParam : UGenT<TestInPlusParam, 1, 1, 1> {
    F number = 0.f;
    inline void tick(Ctx &, const F *in, F *out) { out[0] = number + in[0]; }
    void param_apply(Ctx &, size_t index, const ParamValue &v) override {
        if (index == 0 && v.kind == PK::Float) number = (F)v.f;
    }
};

// ---------------------------------------------------------------- arithmetic wrappers
// knaster_core_dsp/src/wrappers_core/math.rs:15-661
struct WrMath : UGen {
    std::unique_ptr<UGen> ugen;
    int kind;
    F value;
    int ivalue;
    WrMath(std::unique_ptr<UGen> u, int k, double v) : ugen(std::move(u)), kind(k), value((F)v), ivalue((int)v) {}
    int inputs() const override { return ugen->inputs(); }
    int outputs() const override { return ugen->outputs(); }
    // only WrMul adds a parameter ("wr_mul" at index T::Parameters), math.rs:70,78
    int parameters() const override { return ugen->parameters() + (kind == KO_WR_MUL ? 1 : 0); }
    void init(uint32_t sr, size_t bs) override { ugen->init(sr, bs); }
    inline F op(F s) const {
        switch (kind) {
        case KO_WR_MUL: return s * value;   // math.rs:48,65
        case KO_WR_ADD: return s + value;   // math.rs:142,159
        case KO_WR_SUB: return s - value;   // math.rs:220,237
        case KO_WR_VSUB: return value - s;  // math.rs:298,316
        case KO_WR_DIV: return s / value;   // math.rs:377,394
        case KO_WR_VDIV: return value / s;  // math.rs:455,473
        case KO_WR_POWF: return powf(s, value); // math.rs:534,552
        default: { // powi math.rs:613,630 (compiler-rt __powisf2: square-and-multiply)
            int n = ivalue;
            bool recip = n < 0;
            F r = 1.0f, b = s;
            unsigned un = recip ? (unsigned)(-(long long)n) : (unsigned)n;
            while (true) {
                if (un & 1) r *= b;
                un >>= 1;
                if (un == 0) break;
                b *= b;
            }
            return recip ? 1.0f / r : r;
        }
        }
    }
    void process(Ctx &ctx, const F *in, F *out) override {
        ugen->process(ctx, in, out);
        for (int c = 0; c < outputs(); c++) out[c] = op(out[c]);
    }
    void process_block(Ctx &ctx, const F *const *in, F *const *out) override {
        ugen->process_block(ctx, in, out);
        const size_t fr = ctx.block.frames_to_process;
        const int no = outputs();
        for (int c = 0; c < no; c++) {
            F *o = out[c];
            if (kind == KO_WR_MUL) for (size_t i = 0; i < fr; i++) o[i] *= value;
            else for (size_t i = 0; i < fr; i++) o[i] = op(o[i]);
        }
    }
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override {
        if (kind == KO_WR_MUL && index == (size_t)ugen->parameters()) { // math.rs:92-98
            if (v.kind == PK::Float) value = (F)v.f;
        } else ugen->param_apply(ctx, index, v);
    }
    void set_ar_param_buffer(Ctx &ctx, size_t i, const F *b) override { ugen->set_ar_param_buffer(ctx, i, b); }
    void set_delay_within_block_for_param(Ctx &ctx, size_t i, uint16_t d) override {
        ugen->set_delay_within_block_for_param(ctx, i, d);
    }
};

// ---------------------------------------------------------------- WrSmoothParams
// knaster_core_dsp/src/wrappers_core/smooth_params.rs
struct SmoothState { // smooth_params.rs:249-261
    bool linear = false;
    PFloat current_value = 0.0; // None{current_value}, Default = 0. (:312-316)
    PFloat start_value = 0.0, end_value = 0.0;
    uint64_t duration_frames = 0, frames_elapsed = 0;
    int rate = 0;
    bool done = true;
    bool next_value(size_t block_size, size_t frame_in_block, PFloat *out) { // :263-300
        if (!linear) return false;
        if (rate == 0 && frame_in_block != 0) return false;
        if (done) return false;
        PFloat mix = (PFloat)frames_elapsed / (PFloat)duration_frames;
        PFloat cur = (end_value - start_value) * mix + start_value;
        if (frames_elapsed == duration_frames) done = true;
        else if (rate == 0) frames_elapsed = std::min<uint64_t>(frames_elapsed + block_size, duration_frames);
        else frames_elapsed += 1;
        *out = cur;
        return true;
    }
};
struct WrSmoothParams : UGen {
    std::unique_ptr<UGen> ugen;
    std::vector<SmoothState> st;
    explicit WrSmoothParams(std::unique_ptr<UGen> u) : ugen(std::move(u)), st(ugen->parameters()) {}
    int inputs() const override { return ugen->inputs(); }
    int outputs() const override { return ugen->outputs(); }
    int parameters() const override { return ugen->parameters(); }
    void init(uint32_t sr, size_t bs) override { ugen->init(sr, bs); }
    void set_smoothing(size_t index, int smoothing, float seconds, int new_rate, double sr) { // :31-102
        SmoothState &s = st[index];
        if (smoothing == 0) {
            if (s.linear) {
                PFloat mix = (PFloat)s.frames_elapsed / (PFloat)s.duration_frames;
                PFloat cur = (s.end_value - s.start_value) * mix + s.start_value;
                s = SmoothState{};
                s.current_value = cur;
            }
        } else {
            uint64_t dur = sat_usize((double)seconds * sr);
            if (!s.linear) {
                PFloat cur = s.current_value;
                s.linear = true;
                s.start_value = cur; s.end_value = cur;
                s.duration_frames = dur; s.frames_elapsed = 0; s.rate = new_rate; s.done = true;
            } else if (s.done) {
                s.start_value = s.end_value;
                s.duration_frames = dur; s.frames_elapsed = 0; s.rate = new_rate; s.done = true;
            } else {
                PFloat mix = (PFloat)s.frames_elapsed / (PFloat)s.duration_frames;
                PFloat cur = (s.end_value - s.start_value) * mix + s.start_value;
                s.start_value = cur; // end_value, frames_elapsed kept (:89-96)
                s.duration_frames = dur; s.rate = new_rate; s.done = true;
            }
        }
    }
    void process(Ctx &ctx, const F *in, F *out) override { // :115-129
        for (size_t j = 0; j < st.size(); j++) {
            PFloat v;
            if (st[j].next_value(1, 0, &v)) ugen->param_apply(ctx, j, ParamValue::Float(v));
        }
        ugen->process(ctx, in, out);
    }
    void process_block(Ctx &ctx, const F *const *in, F *const *out) override { // :130-189
        // `parameters` (Rate per param) is never set to AudioRate (TODO at :142), so only
        // the block-rate branch :179-188 is reachable.
        for (size_t j = 0; j < st.size(); j++) {
            PFloat v;
            if (st[j].next_value(ctx.block_size, 0, &v)) ugen->param_apply(ctx, j, ParamValue::Float(v));
        }
        ugen->process_block(ctx, in, out);
    }
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override { // :200-244
        if (index >= st.size()) return;
        switch (v.kind) {
        case PK::Integer: case PK::Trigger: case PK::Bool: ugen->param_apply(ctx, index, v); break;
        case PK::Float: {
            SmoothState &s = st[index];
            if (!s.linear) ugen->param_apply(ctx, index, v);
            else {
                if (s.done) s.start_value = s.end_value;
                else {
                    PFloat mix = (PFloat)s.frames_elapsed / (PFloat)s.duration_frames;
                    s.start_value = (s.end_value - s.start_value) * mix + s.start_value;
                }
                s.end_value = v.f;
                s.done = false;
                s.frames_elapsed = 0;
            }
            break;
        }
        case PK::Smoothing: set_smoothing(index, v.smoothing, v.smooth_seconds, v.rate, (double)ctx.sample_rate); break;
        }
    }
    // set_ar_param_buffer / set_delay_within_block_for_param are NOT forwarded (trait defaults)
};

// ---------------------------------------------------------------- WrPreciseTiming
// knaster_core_dsp/src/wrappers_core/precise_timing.rs
struct WrPreciseTiming : UGen {
    std::unique_ptr<UGen> ugen;
    struct Change { bool some = false; uint16_t delay = 0; size_t index = 0; ParamValue value; };
    std::vector<Change> waiting;
    std::vector<uint16_t> next_delay;
    size_t next_delay_i = 0;
    WrPreciseTiming(std::unique_ptr<UGen> u, size_t cap)
        : ugen(std::move(u)), waiting(cap), next_delay(ugen->parameters(), 0) {}
    int inputs() const override { return ugen->inputs(); }
    int outputs() const override { return ugen->outputs(); }
    int parameters() const override { return ugen->parameters(); }
    void init(uint32_t sr, size_t bs) override { ugen->init(sr, bs); }
    void process(Ctx &ctx, const F *in, F *out) override { // :51-64
        for (auto &w : waiting)
            if (w.some) {
                w.some = false;
                ugen->param_apply(ctx, w.index, w.value);
            }
        ugen->process(ctx, in, out);
    }
    void process_block(Ctx &ctx, const F *const *in, F *const *out) override { // :65-114
        size_t block_i = 0, change_i = 0;
        const BlockMeta org_block = ctx.block;
        const size_t num_changes = next_delay_i;
        const int ni = inputs(), no = outputs();
        while (true) {
            size_t local = ctx.block.frames_to_process - block_i;
            while (change_i < num_changes) {
                Change &w = waiting[change_i];
                if (w.some) {
                    if ((size_t)w.delay <= block_i + ctx.block.block_start_offset) {
                        ugen->param_apply(ctx, w.index, w.value);
                        w.some = false;
                    } else {
                        local = std::min(local, (size_t)w.delay - ctx.block.block_start_offset - block_i);
                        break;
                    }
                }
                change_i++;
            }
            if (block_i >= ctx.block.frames_to_process) break;
            if (local == ctx.block.frames_to_process) {
                ugen->process_block(ctx, in, out);
            } else {
                const F *pin[MAX_CH];
                F *pout[MAX_CH];
                for (int i = 0; i < ni; i++) pin[i] = in[i] + block_i;
                for (int i = 0; i < no; i++) pout[i] = out[i] + block_i;
                ctx.block = org_block.make_partial(block_i, local);
                ugen->process_block(ctx, pin, pout);
                ctx.block = org_block;
            }
            block_i += local;
        }
        ctx.block = org_block;
        next_delay_i = 0; // next_delay[] is NOT reset (:111-113)
    }
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override { // :126-135
        if (next_delay[index] == 0) ugen->param_apply(ctx, index, v);
        else if (next_delay_i < waiting.size()) {
            Change &w = waiting[next_delay_i];
            w.some = true; w.delay = next_delay[index]; w.index = index; w.value = v;
            next_delay_i++;
        } else ctx.log_count++;
    }
    void set_ar_param_buffer(Ctx &ctx, size_t i, const F *b) override { ugen->set_ar_param_buffer(ctx, i, b); }
    void set_delay_within_block_for_param(Ctx &, size_t i, uint16_t d) override { next_delay[i] = d; } // :146-148
};

// ---------------------------------------------------------------- WrArParams
// knaster_core_dsp/src/wrappers_core/audio_rate.rs:11-85
struct WrArParams : UGen {
    std::unique_ptr<UGen> ugen;
    std::vector<const F *> buffers;
    size_t block_index = 0;
    explicit WrArParams(std::unique_ptr<UGen> u) : ugen(std::move(u)), buffers(ugen->parameters(), nullptr) {}
    int inputs() const override { return ugen->inputs(); }
    int outputs() const override { return ugen->outputs(); }
    int parameters() const override { return ugen->parameters(); }
    void init(uint32_t sr, size_t bs) override { ugen->init(sr, bs); }
    void process(Ctx &ctx, const F *in, F *out) override { // :42-57
        for (size_t p = 0; p < buffers.size(); p++)
            if (buffers[p]) {
                PFloat value = (PFloat)(double)buffers[p][block_index];
                ugen->param_apply(ctx, p, ParamValue::Float(value));
            }
        block_index = (block_index + 1) % ctx.block_size;
        ugen->process(ctx, in, out);
    }
    // no process_block: the trait default per-frame loop (ugen.rs:263-284) is used
    void param_apply(Ctx &ctx, size_t index, const ParamValue &v) override { // :70-74
        if (!buffers[index]) ugen->param_apply(ctx, index, v);
    }
    void set_ar_param_buffer(Ctx &, size_t i, const F *b) override { buffers[i] = b; } // :76-84
    // set_delay_within_block_for_param not forwarded (trait default)
};

// ---------------------------------------------------------------- factory
std::unique_ptr<UGen> make_ugen(const ko_node_desc &d) {
    std::unique_ptr<UGen> u;
    switch (d.kind) {
    case KO_SIN_WT: u.reset(new SinWt((F)d.args[0])); break;
    case KO_SIN_NUMERIC: u.reset(new SinNumeric((F)d.args[0])); break;
    case KO_POLYBLEP:
        if (d.mode > PolyBlep::TrapezoidVariable) { g_last_error = "oracle: bad PolyBlep waveform"; return nullptr; }
        u.reset(new PolyBlep((int)d.mode, (F)d.args[0]));
        break;
    case KO_SVF: u.reset(new SvfFilter((int)d.mode, (F)d.args[0], (F)d.args[1], (F)d.args[2])); break;
    case KO_ONEPOLE_LPF: u.reset(new OnePoleLpf((F)d.args[0])); break;
    case KO_ONEPOLE_HPF: u.reset(new OnePoleHpf()); break;
    case KO_ENV_ASR: u.reset(new EnvAsr((F)d.args[0], (F)d.args[1])); break;
    case KO_ENV_AR: u.reset(new EnvAr((F)d.args[0], (F)d.args[1])); break;
    case KO_ENVELOPE: {
        std::vector<EnvelopeSegment> segs;
        for (uint32_t i = 0; i < d.n_segments; i++) segs.emplace_back(d.segments[2 * i], d.segments[2 * i + 1]);
        if (segs.empty()) { g_last_error = "oracle: Envelope needs >=1 segment"; return nullptr; }
        u.reset(new Envelope(d.args[0], std::move(segs), (d.flags & 1) != 0, d.args[1]));
        break;
    }
    case KO_MATH:
        if (d.channels < 1 || 2 * d.channels > MAX_CH) { g_last_error = "oracle: bad MathUGen channel count"; return nullptr; }
        u.reset(new MathUGen((int)d.channels, (int)d.mode));
        break;
    case KO_CONSTANT: u.reset(new Constant((F)d.args[0])); break;
    case KO_TEST_NUM: u.reset(new TestNumUGen((F)d.args[0])); break;
    case KO_TEST_IN_PLUS_PARAM: u.reset(new TestInPlusParam()); break;
    case KO_MATH1:
        if (d.mode > 5) { g_last_error = "oracle: bad Math1UGen operation"; return nullptr; }
        u.reset(new Math1UGen((int)d.mode));
        break;
    case KO_PHASOR: u.reset(new Phasor(d.args[0])); break;
    case KO_WHITE_NOISE: u.reset(new WhiteNoise((uint64_t)d.args[0])); break;
    case KO_PINK_NOISE: u.reset(new PinkNoise((uint64_t)d.args[0])); break;
    case KO_BROWN_NOISE: u.reset(new BrownNoise((uint64_t)d.args[0])); break;
    case KO_RANDOM_LIN: u.reset(new RandomLin((F)d.args[0], (uint64_t)d.args[1])); break;
    case KO_PAN2: u.reset(new Pan2((float)d.args[0])); break;
    default: g_last_error = "oracle: unknown ugen kind"; return nullptr;
    }
    for (uint32_t i = 0; i < d.n_wrappers; i++) {
        const ko_wrapper_desc &w = d.wrappers[i];
        switch (w.kind) {
        case KO_WR_MUL: case KO_WR_ADD: case KO_WR_SUB: case KO_WR_VSUB: case KO_WR_DIV:
        case KO_WR_VDIV: case KO_WR_POWF: case KO_WR_POWI:
            u.reset(new WrMath(std::move(u), (int)w.kind, w.value));
            break;
        case KO_WR_SMOOTH_PARAMS: u.reset(new WrSmoothParams(std::move(u))); break;
        case KO_WR_PRECISE_TIMING: u.reset(new WrPreciseTiming(std::move(u), w.capacity)); break;
        case KO_WR_AR_PARAMS: u.reset(new WrArParams(std::move(u))); break;
        default: g_last_error = "oracle: unknown wrapper kind"; return nullptr;
        }
    }
    return u;
}

ParamValue event_value(const ko_event &e) {
    ParamValue v;
    switch (e.value_kind) {
    case 1: v.kind = PK::Float; v.f = e.value; break;
    case 2: v.kind = PK::Trigger; break;
    case 3: v.kind = PK::Integer; v.i = (int64_t)e.value; break;
    case 4: v.kind = PK::Bool; v.b = e.value != 0.0; break;
    default: break;
    }
    return v;
}
ParamValue event_smoothing(const ko_event &e) { // types.rs:115-119 (Rate::BlockRate default)
    ParamValue v;
    v.kind = PK::Smoothing;
    v.smoothing = e.smoothing_kind == 2 ? 1 : 0;
    v.smooth_seconds = e.smooth_seconds;
    v.rate = (int)e.smooth_rate;
    return v;
}

// ---------------------------------------------------------------- graph + processor
// knaster_graph/src/scheduling.rs:29-36,73-121
struct SchedEvent {
    int node;
    size_t parameter;
    bool has_value = false, has_smoothing = false;
    ParamValue value, smoothing;
    bool has_time = false, absolute = false;
    Seconds seconds;
};
struct Edge { int source = -1; uint32_t channel = 0; }; // -1 none, -2 graph input
struct ParamEdge { uint32_t param; int source; uint32_t channel; };

struct Node {
    std::unique_ptr<UGen> ugen;
    std::vector<Edge> in_edges;
    std::vector<ParamEdge> param_edges;
    std::vector<F> out_buf; // [outputs][block]   (one buffer per node: no reuse plan, results identical)
};

} // namespace

struct ko_graph {
    uint32_t sample_rate, block_size, n_inputs, n_outputs;
    size_t ring_capacity;
    Ctx ctx;
    std::vector<Node> nodes;
    std::vector<Edge> output_edges;
    std::vector<int> order;                    // node_task_order
    std::vector<std::vector<const F *>> in_ptrs; // per task
    std::vector<F> zero;                       // buffer_allocator.rs: offset 0 = permanent zero channel
    std::vector<F> output;                     // processor.rs:53-55
    std::deque<SchedEvent> ring;               // rtrb ring (graph.rs:225-230)
    std::deque<std::pair<SchedEvent, uint32_t>> waiting; // graph_gen.rs:49
    uint32_t blocks_to_keep = 0;               // graph_gen.rs:74
    uint64_t frame_clock = 0;                  // processor.rs:57
    const float *const *cur_inputs = nullptr;
    struct Tap { int node; uint32_t channel; };
    std::vector<Tap> taps;
    bool committed = false;
};

namespace {

// knaster_graph/src/scheduling.rs:95-121
uint64_t to_samples_until_due(SchedEvent &e, uint64_t block_size, uint64_t sr, uint64_t frame_clock) {
    if (e.absolute) {
        uint64_t t = e.seconds.to_samples(sr);
        return t > frame_clock ? t - frame_clock : 0;
    }
    if (e.seconds == Seconds{}) return 0;
    uint64_t samples = e.seconds.to_samples(sr);
    e.seconds = e.seconds.saturating_sub(Seconds::from_samples(block_size, sr));
    return samples;
}

// knaster_graph/src/graph_gen.rs:269-305. Returns true if applied.
bool apply_parameter_change(ko_graph *g, SchedEvent &ev) {
    bool ready = true; // tokens unsupported (SchedulingToken::activate is todo!(), scheduling.rs:175-178)
    uint64_t delay = 0;
    if (ev.has_time) {
        delay = to_samples_until_due(ev, g->block_size, g->sample_rate, g->ctx.block.frame_clock);
        ready = ready && (delay < g->block_size);
    }
    if (ready) {
        // linear search of node_task_order (graph_gen.rs:288-289); index lookup is equivalent
        if (ev.node >= 0 && (size_t)ev.node < g->nodes.size()) {
            UGen *u = g->nodes[ev.node].ugen.get();
            if (delay > 0) u->set_delay_within_block_for_param(g->ctx, ev.parameter, (uint16_t)delay);
            if (ev.has_smoothing) u->param_apply(g->ctx, ev.parameter, ev.smoothing);
            if (ev.has_value) u->param_apply(g->ctx, ev.parameter, ev.value);
            return true;
        }
    }
    return false;
}

// knaster_graph/src/graph_gen.rs:77-239
void graph_process_block(ko_graph *g) {
    // retry waiting events :111-140
    if (!g->waiting.empty()) {
        size_t n = g->waiting.size();
        for (size_t i = 0; i < n; i++) {
            auto item = g->waiting.front();
            g->waiting.pop_front();
            if (item.second > g->blocks_to_keep) continue; // dropped :123-126
            if (!apply_parameter_change(g, item.first)) g->waiting.push_back({item.first, item.second + 1});
        }
    }
    // drain ring :143-166
    while (!g->ring.empty()) {
        SchedEvent ev = g->ring.front();
        g->ring.pop_front();
        if (!apply_parameter_change(g, ev)) {
            if (g->waiting.size() < g->ring_capacity) g->waiting.push_back({ev, 0});
        }
    }
    // graph inputs -> node inputs :187-194
    for (size_t ti = 0; ti < g->order.size(); ti++) {
        Node &n = g->nodes[g->order[ti]];
        for (size_t c = 0; c < n.in_edges.size(); c++)
            if (n.in_edges[c].source == -2) g->in_ptrs[ti][c] = g->cur_inputs[n.in_edges[c].channel];
    }
    // run tasks :196-200 ; task.rs:25-31
    for (size_t ti = 0; ti < g->order.size(); ti++) {
        Node &n = g->nodes[g->order[ti]];
        F *outp[MAX_CH];
        const int no = n.ugen->outputs();
        for (int c = 0; c < no; c++) outp[c] = n.out_buf.data() + (size_t)c * g->block_size;
        n.ugen->process_block(g->ctx, g->in_ptrs[ti].data(), outp);
    }
    // outputs :205-224
    for (uint32_t c = 0; c < g->n_outputs; c++) {
        F *dst = g->output.data() + (size_t)c * g->block_size;
        const Edge &e = g->output_edges[c];
        if (e.source >= 0) {
            const F *src = g->nodes[e.source].out_buf.data() + (size_t)e.channel * g->block_size;
            std::memcpy(dst, src, sizeof(F) * g->block_size);
        } else if (e.source == -2) {
            std::memcpy(dst, g->cur_inputs[e.channel], sizeof(F) * g->block_size);
        } else {
            std::fill(dst, dst + g->block_size, 0.f);
        }
    }
}

// knaster_graph/src/processor.rs:119-179
void processor_run(ko_graph *g, const float *const *inputs) {
    g->cur_inputs = inputs;
    g->ctx.block = BlockMeta{0, g->block_size, g->frame_clock}; // set_frame_clock :165
    graph_process_block(g);
    g->frame_clock += g->block_size; // :173
}

SchedEvent to_sched(const ko_event &e) {
    SchedEvent s;
    s.node = (int)e.node;
    s.parameter = e.param;
    if (e.value_kind != 0) { s.has_value = true; s.value = event_value(e); }
    if (e.smoothing_kind != 0) { s.has_smoothing = true; s.smoothing = event_smoothing(e); }
    if (e.time_kind != 0) {
        s.has_time = true;
        s.absolute = e.time_kind == 1;
        s.seconds.seconds = e.seconds;
        s.seconds.sub = e.subsec;
    }
    return s;
}

} // namespace

extern "C" {

const char *ko_last_error(void) { return g_last_error.c_str(); }

ko_graph *ko_graph_create(uint32_t sample_rate, uint32_t block_size, uint32_t n_inputs, uint32_t n_outputs,
                          uint32_t ring_buffer_size) {
    if (block_size == 0 || sample_rate == 0 || n_outputs == 0) { // processor.rs:75 assert, Outputs: NonZero
        g_last_error = "oracle: block_size, sample_rate and n_outputs must be non-zero";
        return nullptr;
    }
    ko_graph *g = new ko_graph();
    g->sample_rate = sample_rate;
    g->block_size = block_size;
    g->n_inputs = n_inputs;
    g->n_outputs = n_outputs;
    g->ring_capacity = ring_buffer_size;
    g->ctx.sample_rate = sample_rate;
    g->ctx.block_size = block_size;
    g->ctx.block = BlockMeta{0, block_size, 0};
    g->output_edges.assign(n_outputs, Edge{});
    g->zero.assign(block_size, 0.f);
    g->output.assign((size_t)block_size * n_outputs, 0.f);
    g->blocks_to_keep = sample_rate / block_size; // graph_gen.rs:73-75
    return g;
}
void ko_graph_destroy(ko_graph *g) { delete g; }

int ko_push(ko_graph *g, const ko_node_desc *desc) {
    std::unique_ptr<UGen> u = make_ugen(*desc);
    if (!u) return -1;
    u->init(g->sample_rate, g->block_size); // graph.rs:462-475
    Node n;
    n.in_edges.assign(u->inputs(), Edge{});
    n.out_buf.assign((size_t)u->outputs() * g->block_size, 0.f);
    n.ugen = std::move(u);
    g->nodes.push_back(std::move(n));
    g->committed = false;
    return (int)g->nodes.size() - 1;
}
int ko_node_inputs(ko_graph *g, int node) { return g->nodes.at(node).ugen->inputs(); }
int ko_node_outputs(ko_graph *g, int node) { return g->nodes.at(node).ugen->outputs(); }
int ko_node_parameters(ko_graph *g, int node) { return g->nodes.at(node).ugen->parameters(); }

static int check_source(ko_graph *g, int source_node, uint32_t source_channel) {
    if (source_node == -1) return 0;
    if (source_node == -2) {
        if (source_channel >= g->n_inputs) { g_last_error = "oracle: graph input out of bounds"; return -1; }
        return 0;
    }
    if (source_node < 0 || (size_t)source_node >= g->nodes.size()) { g_last_error = "oracle: source node not found"; return -1; }
    if ((int)source_channel >= g->nodes[source_node].ugen->outputs()) { g_last_error = "oracle: output out of bounds"; return -1; }
    return 0;
}
int ko_set_input_edge(ko_graph *g, int sink_node, uint32_t sink_channel, int source_node, uint32_t source_channel) {
    if (sink_node < 0 || (size_t)sink_node >= g->nodes.size()) { g_last_error = "oracle: sink node not found"; return -1; }
    Node &n = g->nodes[sink_node];
    if (sink_channel >= n.in_edges.size()) { g_last_error = "oracle: input out of bounds"; return -1; }
    if (check_source(g, source_node, source_channel)) return -1;
    n.in_edges[sink_channel] = Edge{source_node, source_channel};
    g->committed = false;
    return 0;
}
int ko_set_output_edge(ko_graph *g, uint32_t out_channel, int source_node, uint32_t source_channel) {
    if (out_channel >= g->n_outputs) { g_last_error = "oracle: graph output out of bounds"; return -1; }
    if (check_source(g, source_node, source_channel)) return -1;
    g->output_edges[out_channel] = Edge{source_node, source_channel};
    g->committed = false;
    return 0;
}
int ko_set_param_edge(ko_graph *g, int sink_node, uint32_t param_index, int source_node, uint32_t source_channel) {
    if (sink_node < 0 || (size_t)sink_node >= g->nodes.size()) { g_last_error = "oracle: sink node not found"; return -1; }
    if (source_node < 0 || check_source(g, source_node, source_channel)) { g_last_error = "oracle: bad parameter source"; return -1; }
    Node &n = g->nodes[sink_node];
    if ((int)param_index >= n.ugen->parameters()) { g_last_error = "oracle: parameter index out of bounds"; return -1; }
    for (auto &pe : n.param_edges)
        if (pe.param == param_index) { pe.source = source_node; pe.channel = source_channel; g->committed = false; return 0; }
    n.param_edges.push_back(ParamEdge{param_index, source_node, source_channel});
    g->committed = false;
    return 0;
}

// Graph::commit_changes graph.rs:1707-1726: node order (post-order DFS from the output
// edges, parameter edges count as dependencies, graph.rs:1938-2067), then every remaining
// (disconnected) node; tasks + AR parameter buffers (graph.rs:1532-1562, task.rs:113-120).
// Any valid topological order renders identical samples for a feedback-free graph.
int ko_commit(ko_graph *g) {
    const size_t N = g->nodes.size();
    std::vector<char> visited(N, 0), on_stack(N, 0);
    g->order.clear();
    std::vector<std::pair<int, size_t>> stack;
    auto dfs = [&](int root) -> bool {
        if (visited[root]) return true;
        visited[root] = 1;
        stack.push_back({root, 0});
        on_stack[root] = 1;
        while (!stack.empty()) {
            int nk = stack.back().first;
            size_t &cursor = stack.back().second;
            Node &n = g->nodes[nk];
            const size_t ne = n.in_edges.size(), np = n.param_edges.size();
            bool pushed = false;
            while (cursor < ne + np) {
                int src = cursor < ne ? n.in_edges[cursor].source : n.param_edges[cursor - ne].source;
                cursor++;
                if (src >= 0) {
                    if (on_stack[src]) { g_last_error = "oracle: cycle in graph (feedback edges unsupported)"; return false; }
                    if (!visited[src]) {
                        visited[src] = 1;
                        on_stack[src] = 1;
                        stack.push_back({src, 0});
                        pushed = true;
                        break;
                    }
                }
            }
            if (!pushed) {
                on_stack[nk] = 0;
                g->order.push_back(nk);
                stack.pop_back();
            }
        }
        return true;
    };
    for (auto &e : g->output_edges)
        if (e.source >= 0 && !dfs(e.source)) return -1;
    for (size_t i = 0; i < N; i++)
        if (!visited[i] && !dfs((int)i)) return -1;
    g->in_ptrs.assign(g->order.size(), {});
    for (size_t ti = 0; ti < g->order.size(); ti++) {
        Node &n = g->nodes[g->order[ti]];
        g->in_ptrs[ti].resize(n.in_edges.size());
        for (size_t c = 0; c < n.in_edges.size(); c++) {
            const Edge &e = n.in_edges[c];
            if (e.source >= 0) g->in_ptrs[ti][c] = g->nodes[e.source].out_buf.data() + (size_t)e.channel * g->block_size;
            else g->in_ptrs[ti][c] = g->zero.data();
        }
        for (auto &pe : n.param_edges)
            n.ugen->set_ar_param_buffer(g->ctx, pe.param, g->nodes[pe.source].out_buf.data() + (size_t)pe.channel * g->block_size);
    }
    g->committed = true;
    return 0;
}

int ko_send_event(ko_graph *g, const ko_event *ev) {
    if (ev->node >= g->nodes.size()) { g_last_error = "oracle: event node not found"; return -1; }
    if ((int)ev->param >= g->nodes[ev->node].ugen->parameters()) { g_last_error = "oracle: event parameter out of bounds"; return -1; }
    if (g->ring.size() >= g->ring_capacity) { g_last_error = "oracle: scheduling ring full"; return -2; } // handle.rs:38-73
    g->ring.push_back(to_sched(*ev));
    return 0;
}

int ko_run_block(ko_graph *g, const float *const *inputs) {
    if (!g->committed && ko_commit(g)) return -1;
    if (g->n_inputs > 0 && !inputs) { g_last_error = "oracle: inputs required"; return -1; }
    processor_run(g, inputs);
    return 0;
}
const float *ko_output_block(ko_graph *g) { return g->output.data(); }
uint64_t ko_frame_clock(ko_graph *g) { return g->frame_clock; }
uint64_t ko_log_count(ko_graph *g) { return g->ctx.log_count; }

int ko_add_tap(ko_graph *g, int node, uint32_t channel) {
    if (node < 0 || (size_t)node >= g->nodes.size() || (int)channel >= g->nodes[node].ugen->outputs()) {
        g_last_error = "oracle: bad tap";
        return -1;
    }
    g->taps.push_back({node, channel});
    return (int)g->taps.size() - 1;
}

int ko_render(ko_graph *g, uint64_t n_blocks, float *out, const ko_event *events, size_t n_events, float *taps_out) {
    if (g->n_inputs != 0) { g_last_error = "oracle: ko_render needs a graph without inputs"; return -1; }
    if (!g->committed && ko_commit(g)) return -1;
    const uint64_t bs = g->block_size;
    const uint64_t first_block = g->frame_clock / bs;
    // just-in-time feed order: stable sort by due block
    std::vector<std::pair<uint64_t, size_t>> feed(n_events);
    for (size_t i = 0; i < n_events; i++) {
        uint64_t due_block = first_block;
        if (events[i].time_kind == 1) {
            Seconds s{events[i].seconds, events[i].subsec};
            uint64_t fr = s.to_samples(g->sample_rate);
            due_block = std::max(first_block, fr / bs);
        }
        feed[i] = {due_block, i};
    }
    std::stable_sort(feed.begin(), feed.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    size_t cursor = 0;
    const size_t total = (size_t)n_blocks * bs;
    for (uint64_t b = 0; b < n_blocks; b++) {
        const uint64_t blk = first_block + b;
        while (cursor < n_events && feed[cursor].first <= blk) {
            const ko_event &e = events[feed[cursor].second];
            if (e.node >= g->nodes.size() || (int)e.param >= g->nodes[e.node].ugen->parameters()) {
                g_last_error = "oracle: event node/parameter out of bounds";
                return -1;
            }
            g->ring.push_back(to_sched(e)); // unbounded here: the harness never overflows the ring
            cursor++;
        }
        processor_run(g, nullptr);
        if (out) std::memcpy(out + b * bs * g->n_outputs, g->output.data(), sizeof(F) * bs * g->n_outputs);
        if (taps_out)
            for (size_t t = 0; t < g->taps.size(); t++)
                std::memcpy(taps_out + t * total + b * bs,
                            g->nodes[g->taps[t].node].out_buf.data() + (size_t)g->taps[t].channel * bs, sizeof(F) * bs);
    }
    return 0;
}

void ko_seconds_from_samples(uint64_t samples, uint64_t sr, uint32_t *seconds, uint32_t *subsec) {
    Seconds s = Seconds::from_samples(samples, sr);
    *seconds = s.seconds;
    *subsec = s.sub;
}
uint64_t ko_seconds_to_samples(uint32_t seconds, uint32_t subsec, uint64_t sr) { return Seconds{seconds, subsec}.to_samples(sr); }
void ko_seconds_from_secs_f64(double v, uint32_t *seconds, uint32_t *subsec) {
    Seconds s = Seconds::from_secs_f64(v);
    *seconds = s.seconds;
    *subsec = s.sub;
}
uint64_t ko_seconds_to_tesimals(uint32_t seconds, uint32_t subsec) { return Seconds{seconds, subsec}.to_tesimals(); }
void ko_seconds_from_tesimals(uint64_t t, uint32_t *seconds, uint32_t *subsec) {
    Seconds s = Seconds::from_tesimals(t);
    *seconds = s.seconds;
    *subsec = s.sub;
}
// impl Add for Seconds, knaster_primitives/src/time.rs (carry into seconds)
void ko_seconds_add(uint32_t s0, uint32_t t0, uint32_t s1, uint32_t t1, uint32_t *seconds, uint32_t *subsec) {
    uint64_t sub = (uint64_t)t0 + t1;
    uint32_t sec = s0 + s1;
    if (sub >= TESIMALS) { sub -= TESIMALS; sec += 1; }
    *seconds = sec;
    *subsec = (uint32_t)sub;
}

struct ko_ugen {
    std::unique_ptr<UGen> u;
    Ctx ctx;
};
ko_ugen *ko_ugen_create(const ko_node_desc *desc, uint32_t sample_rate, uint32_t block_size) {
    std::unique_ptr<UGen> u = make_ugen(*desc);
    if (!u) return nullptr;
    ko_ugen *h = new ko_ugen();
    h->ctx.sample_rate = sample_rate;
    h->ctx.block_size = block_size;
    h->ctx.block = BlockMeta{0, block_size, 0};
    u->init(sample_rate, block_size);
    h->u = std::move(u);
    return h;
}
void ko_ugen_destroy(ko_ugen *u) { delete u; }
int ko_ugen_set_delay(ko_ugen *u, uint32_t param, uint32_t delay) {
    u->u->set_delay_within_block_for_param(u->ctx, param, (uint16_t)delay);
    return 0;
}
int ko_ugen_param(ko_ugen *u, uint32_t param, const ko_event *value) { // UGen::param ugen.rs:344-357
    if ((int)param >= u->u->parameters()) { g_last_error = "ParameterIndexOutOfBounds"; return -1; }
    if (value->smoothing_kind != 0) u->u->param_apply(u->ctx, param, event_smoothing(*value));
    if (value->value_kind != 0) u->u->param_apply(u->ctx, param, event_value(*value));
    return 0;
}
int ko_ugen_process_block(ko_ugen *u, const float *in, float *out, uint32_t frames) {
    const F *pin[MAX_CH];
    F *pout[MAX_CH];
    for (int i = 0; i < u->u->inputs(); i++) pin[i] = in + (size_t)i * frames;
    for (int i = 0; i < u->u->outputs(); i++) pout[i] = out + (size_t)i * frames;
    u->ctx.block = BlockMeta{0, frames, 0};
    u->u->process_block(u->ctx, pin, pout);
    return 0;
}
int ko_ugen_process(ko_ugen *u, const float *in, float *out) {
    u->u->process(u->ctx, in, out);
    return 0;
}

} // extern "C"

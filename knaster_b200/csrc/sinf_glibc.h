// sinf_glibc.h -- f32 sine with glibc's results, usable on host and device.
//
// knaster's SinNumeric / PolyBlep call f32::sin (osc.rs:264, polyblep.rs:244), which is the
// platform libm's sinf (num-traits `std`, knaster_primitives/Cargo.toml:12,19): a third-party
// dependency outside /root/reference.  Pinned version here: glibc 2.39 (the image's libm; the
// algorithm is the ARM optimized-routines sinf that glibc adopted in 2.28: sysdeps/ieee754/flt-32/
// s_sinf.c + s_sincosf.h + s_sincosf_data.c).  This is a restatement of that published algorithm:
// fast range reduction by pi/2 in f64 and a degree-7 sine / degree-8 cosine polynomial in f64,
// rounded once to f32.  tests/test_host_plan.py::test_sinf_restatement_matches_libm checks it
// bit-for-bit against the libm the oracle links, over millions of arguments in [-16, 16] (with and
// without FMA contraction the rounded results are identical there).
//
// Why it matters: an FM carrier integrates its modulator's output, so 1-ulp differences between the
// device sine and libm's would random-walk the carrier phase past the 1e-5 budget within seconds.
#pragma once
#include <stdint.h>
#include <string.h>
#if defined(__CUDACC__)
#define KN_SINF_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define KN_SINF_HD inline
#endif

namespace kgpu {

KN_SINF_HD double kn_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return a * b + c; // contraction does not change the rounded f32 (verified exhaustively on the tested range)
#endif
}

// Returns true and the sine in *out for |y| < 120 (glibc's reduce_fast domain); false otherwise.
KN_SINF_HD bool kn_sinf_glibc(float y, float *out) {
    const double HPI_INV = 0x1.45F306DC9C883p+23; // 2^24 / (pi/2)
    const double HPI = 0x1.921FB54442D18p0;       // pi/2
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    uint32_t iy;
#if defined(__CUDA_ARCH__)
    iy = __float_as_uint(y);
#else
    memcpy(&iy, &y, 4);
#endif
    const uint32_t top = (iy >> 20) & 0x7ff;          // abstop12
    double x = (double)y;
    if (top < 0x3f4) {                                // |y| < pi/4  (abstop12(0x1.921FB6p-1f) = 0x3f4)
        if (top < 0x398) {                            // |y| < 2^-12: sin(y) = y to f32 precision
            *out = y;
            return true;
        }
        const double x2 = x * x, x3 = x * x2, s1 = kn_fma(x2, S3, S2), x7 = x3 * x2, s = kn_fma(x3, S1, x);
        *out = (float)kn_fma(x7, s1, s);
        return true;
    }
    if (top >= 0x42f) return false;                   // |y| >= 120: large-argument reduction not restated
    const double r = x * HPI_INV;
#if defined(__CUDA_ARCH__)
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
#else
    const int n = ((int32_t)r + 0x800000) >> 24;
#endif
    x = kn_fma(-(double)n, HPI, x);
    const double x2 = x * x;
    double res;
    if ((n & 1) == 0) {                               // sine polynomial (odd in x)
        const double x3 = x * x2, s1 = kn_fma(x2, S3, S2), x7 = x3 * x2, s = kn_fma(x3, S1, x);
        res = kn_fma(x7, s1, s);
        if (((n + 1) & 2) != 0) res = -res;           // sign table {1,-1,-1,1}[n & 3]
    } else {                                          // cosine polynomial (even in x)
        const double x4 = x2 * x2, c2 = kn_fma(x2, C4, C3), c1 = kn_fma(x2, C1, C0), x6 = x4 * x2, c = kn_fma(x4, C2, c1);
        res = kn_fma(x6, c2, c);
        if ((n & 2) != 0) res = -res;                 // __sincosf_table[1]: negated cosine coefficients
    }
    *out = (float)res;
    return true;
}

} // namespace kgpu

// Microbenchmark: issue rate / latency of FADD vs FADD2 (packed f32x2) on sm_100a with ONE warp
// per SM sub-partition (the occupancy the voice-bank kernels run at).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mul1(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int MODE, int ILP>
__global__ void bench(float *out, long long *cyc, int iters, float seed) {
    float s[ILP]; u64 p[ILP];
    float sel[ILP];
    for (int i = 0; i < ILP; i++) { s[i] = seed + i + threadIdx.x; p[i] = ((u64)__float_as_uint(s[i]) << 32) | __float_as_uint(s[i] + 0.5f); sel[i] = s[i]; }
    const float inc = seed * 0.25f; const u64 inc2 = ((u64)__float_as_uint(inc) << 32) | __float_as_uint(inc);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (MODE == 0) s[i] = add1(s[i], inc);                       // FADD
                if (MODE == 1) p[i] = add2(p[i], inc2);                      // FADD2
                if (MODE == 2) { s[i] = add1(s[i], inc); sel[i] = s[i] >= 1.0f ? sel[i] - 1.0f : sel[i]; }  // FADD + FSETP + @P FADD
                if (MODE == 3) { p[i] = add2(p[i], inc2); sel[i] = sel[i] > seed ? sel[i] : inc; }       // FADD2 + FSETP/FSEL (alu)
                if (MODE == 4) s[i] = mul1(s[i], inc);                       // FMUL
                if (MODE == 5) p[i] = mul2(p[i], inc2);                      // FMUL2
                if (MODE == 6) { s[i] = add1(s[i], inc); sel[i] = sel[i] > seed ? sel[i] : inc; }        // FADD + alu
            }
        }
    }
    long long t1 = clock64();
    float acc = 0; for (int i = 0; i < ILP; i++) acc += s[i] + sel[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP> void run(const char *name, int warps_per_cta, int ops_per) {
    float *out; long long *cyc; int nb = 148 * 4 / (warps_per_cta >= 4 ? 4 : 1);
    if (warps_per_cta == 1) nb = 512;
    cudaMalloc(&out, nb * 32 * warps_per_cta * 4); cudaMalloc(&cyc, nb * 8);
    int iters = 20000;
    bench<MODE, ILP><<<nb, 32 * warps_per_cta>>>(out, cyc, 100, 1.0f);
    bench<MODE, ILP><<<nb, 32 * warps_per_cta>>>(out, cyc, iters, 1.0f);
    long long h[1024]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; i++) avg += h[i]; avg /= nb;
    double per = avg / ((double)iters * 8 * ILP);
    printf("%-28s ILP=%2d warps/cta=%d  cycles per (op group) = %.3f  -> %.3f cycles per warp-instr (%d instr/group)\n", name, ILP, warps_per_cta, per, per / ops_per, ops_per);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 1>("FADD dep chain", 1, 1); run<1, 1>("FADD2 dep chain", 1, 1);
    run<0, 8>("FADD", 1, 1); run<1, 8>("FADD2", 1, 1);
    run<0, 16>("FADD", 1, 1); run<1, 16>("FADD2", 1, 1);
    run<4, 8>("FMUL", 1, 1); run<5, 8>("FMUL2", 1, 1);
    run<2, 8>("FADD+FSETP+@P FADD", 1, 3); run<3, 8>("FADD2+FSETP+FSEL", 1, 3); run<6, 8>("FADD+FSETP+FSEL", 1, 3);
    run<0, 8>("FADD 4 warps/SM-CTA", 4, 1); run<1, 8>("FADD2 4 warps", 4, 1);
    run<0, 8>("FADD 8 warps", 8, 1); run<1, 8>("FADD2 8 warps", 8, 1);
    run<3, 8>("FADD2+FSETP+FSEL 8 warps", 8, 3); run<6, 8>("FADD+FSETP+FSEL 8w", 8, 3);
    return 0;
}

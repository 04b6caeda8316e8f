// Microbenchmark: how FP64 instructions share issue slots with FP32 / integer-pipe instructions on sm_100a, ONE warp per SM
// sub-partition.  Each pattern is ND independent DFMA chains + NF FADD chains + NA select (ALU pipe) chains per round, 8 rounds per
// loop trip, the trip loop NOT unrolled (the body stays inside the instruction cache).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double xdfma(double a, double b, double c) { double r; asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c)); return r; }
__device__ __forceinline__ float xfadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned xlop(unsigned a, unsigned b) { unsigned r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int ND, int NF, int NA>
__global__ void bench(float *out, long long *cyc, int iters, float seed) {
    double d[ND > 0 ? ND : 1]; float f[NF > 0 ? NF : 1]; unsigned u[NA > 0 ? NA : 1];
    for (int i = 0; i < ND; i++) d[i] = seed + i + threadIdx.x;
    for (int i = 0; i < NF; i++) f[i] = seed * 0.5f + i;
    for (int i = 0; i < NA; i++) u[i] = threadIdx.x + i;
    const double c1 = 1.0 + 1e-9 * seed, c2 = 1e-3 * seed; const float finc = seed * 0.25f; const unsigned um = __float_as_uint(seed);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            // interleave: one of each kind in turn
#pragma unroll
            for (int i = 0; i < (ND > NF ? (ND > NA ? ND : NA) : (NF > NA ? NF : NA)); i++) {
                if (i < ND) d[i] = xdfma(d[i], c1, c2);
                if (i < NF) f[i] = xfadd(f[i], finc);
                if (i < NA) u[i] = xlop(u[i], um);
            }
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < ND; i++) acc += (float)d[i];
    for (int i = 0; i < NF; i++) acc += f[i];
    for (int i = 0; i < NA; i++) acc += (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ND, int NF, int NA> void run() {
    const int nb = 512;
    float *out; long long *cyc;
    cudaMalloc(&out, nb * 32 * 4); cudaMalloc(&cyc, nb * 8);
    int iters = 4000;
    bench<ND, NF, NA><<<nb, 32>>>(out, cyc, 100, 1.0f);
    bench<ND, NF, NA><<<nb, 32>>>(out, cyc, iters, 1.0f);
    static long long h[4096]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; i++) avg += h[i]; avg /= nb;
    const double per_round = avg / ((double)iters * 8);
    printf("DFMA x%2d + FADD x%2d + LOP x%2d per round: %7.2f cycles per round (%d instr) = %.3f per instr; model 2*ND=%d, ND+NF+NA=%d\n", ND, NF, NA, per_round,
           ND + NF + NA, per_round / (ND + NF + NA), 2 * ND, ND + NF + NA);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<8, 0, 0>(); run<0, 8, 0>(); run<0, 0, 8>(); run<0, 16, 0>(); run<0, 0, 16>();
    run<8, 8, 0>(); run<8, 16, 0>(); run<8, 24, 0>(); run<4, 12, 0>();
    run<8, 0, 8>(); run<8, 0, 4>(); run<8, 8, 4>(); run<8, 8, 8>(); run<8, 12, 4>(); run<0, 8, 8>(); run<0, 16, 8>();
    run<6, 10, 4>(); run<4, 4, 0>(); run<4, 8, 0>(); run<4, 8, 2>();
    return 0;
}

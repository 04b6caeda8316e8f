// Microbenchmark: latency and issue rate of the FP64 pipe and of the f32<->f64 conversions on sm_100a with ONE warp per SM
// sub-partition (the occupancy render_fm2 runs at: its sine is glibc's, an f64 polynomial).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o fp64_bench fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double xdfma(double a, double b, double c) { double r; asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c)); return r; }
__device__ __forceinline__ double xdmul(double a, double b) { double r; asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b)); return r; }
__device__ __forceinline__ double xdadd(double a, double b) { double r; asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ double up(float a) { double r; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(r) : "f"(a)); return r; }
__device__ __forceinline__ float down(double a) { float r; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(a)); return r; }

template <int MODE, int ILP>
__global__ void bench(float *out, long long *cyc, int iters, float seed) {
    double d[ILP]; float f[ILP], g[ILP];
    for (int i = 0; i < ILP; i++) { d[i] = seed + i + threadIdx.x; f[i] = seed * 0.5f + i; g[i] = f[i]; }
    const double c1 = 1.0 + 1e-9 * seed, c2 = 1e-3 * seed; const float finc = seed * 0.25f;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (MODE == 0) d[i] = xdfma(d[i], c1, c2);
                if (MODE == 1) d[i] = xdmul(d[i], c1);
                if (MODE == 2) d[i] = xdadd(d[i], c2);
                if (MODE == 3) { d[i] = xdfma(d[i], c1, c2); f[i] = fadd(f[i], finc); }             // does an FP32 op fill the FP64 shadow?
                if (MODE == 4) { d[i] = xdfma(d[i], c1, c2); f[i] = fadd(f[i], finc); g[i] = fadd(g[i], finc); }
                if (MODE == 5) f[i] = down(up(f[i]));                                             // 2 conversions, dependent
                if (MODE == 6) { d[i] = xdfma(d[i], c1, c2); f[i] = down(up(f[i])); }               // 1 DFMA + 2 conversions
                if (MODE == 7) { d[i] = xdfma(d[i], c1, c2); g[i] = g[i] > seed ? g[i] : finc; }    // DFMA + FSETP/FSEL (alu)
            }
        }
    }
    long long t1 = clock64();
    float acc = 0; for (int i = 0; i < ILP; i++) acc += (float)d[i] + f[i] + g[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP> void run(const char *name, int warps_per_cta, int nb) {
    float *out; long long *cyc;
    cudaMalloc(&out, nb * 32 * warps_per_cta * 4); cudaMalloc(&cyc, nb * 8);
    int iters = 5000;
    bench<MODE, ILP><<<nb, 32 * warps_per_cta>>>(out, cyc, 100, 1.0f);
    bench<MODE, ILP><<<nb, 32 * warps_per_cta>>>(out, cyc, iters, 1.0f);
    static long long h[4096]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; i++) avg += h[i]; avg /= nb;
    printf("%-34s ILP=%2d warps/cta=%d ctas=%4d  cycles per op group = %.3f\n", name, ILP, warps_per_cta, nb, avg / ((double)iters * 8 * ILP));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 1>("DFMA dependent chain", 1, 512); run<1, 1>("DMUL dependent chain", 1, 512); run<2, 1>("DADD dependent chain", 1, 512);
    run<0, 4>("DFMA", 1, 512); run<0, 8>("DFMA", 1, 512); run<0, 16>("DFMA", 1, 512);
    run<0, 8>("DFMA one warp per SM", 1, 148); run<0, 8>("DFMA 4 warps per SM (one CTA)", 4, 148); run<0, 8>("DFMA 8 warps per SM", 8, 148);
    run<3, 8>("DFMA + FADD", 1, 512); run<4, 8>("DFMA + 2 FADD", 1, 512); run<7, 8>("DFMA + FSETP + FSEL", 1, 512);
    run<5, 1>("cvt up + cvt down dependent", 1, 512); run<5, 8>("cvt up + cvt down", 1, 512); run<5, 16>("cvt up + cvt down", 1, 512);
    run<6, 8>("DFMA + cvt up + cvt down", 1, 512);
    return 0;
}

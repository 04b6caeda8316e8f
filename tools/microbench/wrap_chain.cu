// Microbenchmark: cycles per frame of the phase recurrence t' = wrap01(t + dt) (polyblep.rs:232-235) in the forms the scan kernels
// could use; ONE warp per SM sub-partition, one dependent chain.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o wrap_chain wrap_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void bench(float *out, long long *cyc, int iters, float dt, float T, float c) {
    float t = 0.125f + 0.001f * threadIdx.x * 0.f;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 32; k++) {
            if (MODE == 0) { const float a = t + dt; t = a - (a >= 1.0f ? 1.0f : 0.0f); }                    // FADD -> FSET -> FADD
            if (MODE == 1) { t = (t + dt) - (t >= T ? 1.0f : 0.0f); }                                         // FSET || FADD -> FADD
            if (MODE == 2) { t = (t + dt) - __saturatef(__fmaf_rn(t, 16777216.0f, c)); }                      // FFMA.SAT || FADD -> FADD
            if (MODE == 3) { float a = t + dt; if (t >= T) a = a - 1.0f; t = a; }                             // FSETP || FADD -> @P FADD
            if (MODE == 4) { const float a = t + dt; t = a >= 1.0f ? a - 1.0f : a; }                          // select form
            if (MODE == 5) { const float a = t + dt; const float f = fminf(fmaxf(__fmaf_rn(t, 16777216.0f, c), 0.0f), 1.0f); t = a - f; }
            if (MODE == 6) { const float a = t + dt; const float b = a - 1.0f; t = (t >= T) ? b : a; }        // both candidates, FSEL on old-phase predicate
            if (MODE == 7) { t = t + dt; }                                                                    // no wrap (floor)
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name) {
    const int nb = 512;
    float *out; long long *cyc;
    cudaMalloc(&out, nb * 32 * 4); cudaMalloc(&cyc, nb * 8);
    const float dt = 440.0f / 48000.0f;
    float T = 1.0f - dt;
    while (T + dt >= 1.0f) T = nextafterf(T, 0.0f);
    while (T + dt < 1.0f) T = nextafterf(T, 2.0f);
    const float c = 1.0f - T * 16777216.0f;
    int iters = 20000;
    bench<MODE><<<nb, 32>>>(out, cyc, 100, dt, T, c);
    bench<MODE><<<nb, 32>>>(out, cyc, iters, dt, T, c);
    static long long h[4096]; static float o[512 * 32];
    cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost); cudaMemcpy(o, out, 4, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; i++) avg += h[i]; avg /= nb;
    printf("%-58s %6.2f cycles per frame   (t = %.9g)\n", name, avg / ((double)iters * 32), o[0]);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("a = t + dt; t = a - (a >= 1): FADD -> FSET -> FADD");
    run<1>("t = (t + dt) - (t >= T): FSET || FADD -> FADD");
    run<2>("t = (t + dt) - sat(fma(t, 2^24, c)): FFMA.SAT || FADD -> FADD");
    run<3>("a = t + dt; if (t >= T) a -= 1: FSETP || FADD -> @P FADD");
    run<4>("a = t + dt; t = a >= 1 ? a - 1 : a: FADD -> FSETP, FADD -> FSEL");
    run<5>("flag = min(max(fma(t, 2^24, c), 0), 1): FFMA -> FMNMX x2 -> FADD");
    run<6>("a = t + dt; b = a - 1; t = (t >= T) ? b : a: FSEL on the old-phase predicate");
    run<7>("t = t + dt (no wrap: the floor)");
    return 0;
}

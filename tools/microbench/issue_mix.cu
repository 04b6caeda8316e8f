// Microbenchmark: what ONE warp per SM sub-partition can issue per cycle on the instruction mix of render_sub_asr's
// 8-frame loop (fused.cu): FADD / FMUL / FFMA with three register operands / FFMA with an immediate / FSET, alone and in
// the loop's proportions (per frame 14 FADD, 11 FMUL, 3 + 3 FFMA, 4 FSET), as independent chains (ILP 8: no dependency
// stalls), so that what is left is the dispatch rate of the pipes.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o issue_mix issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#define DEV __device__ __forceinline__
DEV float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
DEV float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
DEV float ffma3(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
DEV float ffmai(float a, float c) { float r; asm volatile("fma.rn.f32 %0, %1, 0f40000000, %2;" : "=f"(r) : "f"(a), "f"(c)); return r; } // a * 2 + c
DEV float fset(float a, float b) { float r; asm volatile("set.ge.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }     // 1.0 / 0.0: FSET.BF

template <int MODE>
__global__ void bench(float *out, long long *cyc, int iters, float seed) {
    constexpr int ILP = 8;
    float s[ILP], u[ILP], w[ILP];
    for (int i = 0; i < ILP; i++) { s[i] = seed + i + threadIdx.x; u[i] = seed * 0.5f + i; w[i] = seed * 0.125f - i; }
    const float a = seed * 0.25f, b = seed * 0.75f;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) { for (int r = 0; r < 8; r++) s[i] = fadd(s[i], a); }
            if (MODE == 1) { for (int r = 0; r < 8; r++) s[i] = fmul(s[i], a); }
            if (MODE == 2) { for (int r = 0; r < 8; r++) s[i] = ffma3(s[i], a, b); }        // 3 registers, two of them shared
            if (MODE == 3) { for (int r = 0; r < 8; r++) s[i] = ffma3(s[i], u[i], w[i]); }  // 3 registers, all per chain
            if (MODE == 4) { for (int r = 0; r < 8; r++) s[i] = ffmai(s[i], b); }           // immediate form
            if (MODE == 5) { for (int r = 0; r < 8; r++) s[i] = fset(s[i], a); }
        }
        if (MODE == 6) { // the loop's mix, op by op across the 8 chains (as the unrolled 8-frame group interleaves its frames): 14 FADD, 11 FMUL, 3 FFMA (3 reg), 3 FFMA (imm), 4 FSET = 35 per chain
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fset(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=ffma3(s[i],u[i],w[i]);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=ffmai(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fmul(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fset(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=ffma3(w[i],s[i],u[i]);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=ffmai(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fset(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fmul(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=ffma3(s[i],w[i],u[i]);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fset(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=ffmai(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fmul(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],b);
        }
        if (MODE == 7) { // the same without the three 3-register FFMAs (FADD in their place)
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fset(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],u[i]);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=ffmai(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fmul(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fset(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],s[i]);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=ffmai(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fset(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fmul(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],w[i]);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fset(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=ffmai(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fmul(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],b);
        }
        if (MODE == 8) { // and without the four FSETs (FADD in their place)
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=ffma3(s[i],u[i],w[i]);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=ffmai(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fmul(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=ffma3(w[i],s[i],u[i]);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=ffmai(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fmul(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=ffma3(s[i],w[i],u[i]);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fadd(u[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fmul(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fadd(s[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=ffmai(u[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) w[i]=fadd(w[i],b);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) s[i]=fmul(s[i],a);
                _Pragma("unroll") for (int i = 0; i < ILP; i++) u[i]=fmul(u[i],b);
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < ILP; i++) acc += s[i] + u[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, int instr_per_chain) {
    float *out; long long *cyc; const int nb = 512;
    cudaMalloc(&out, nb * 32 * 4); cudaMalloc(&cyc, nb * 8);
    const int iters = 20000;
    bench<MODE><<<nb, 32>>>(out, cyc, 100, 1.0f);
    bench<MODE><<<nb, 32>>>(out, cyc, iters, 1.0f);
    long long h[512]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; i++) avg += h[i]; avg /= nb;
    printf("%-52s %.3f cycles per warp instruction (%d per chain, 8 chains)\n", name, avg / ((double)iters * 8 * instr_per_chain), instr_per_chain);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FADD", 8); run<1>("FMUL", 8);
    run<2>("FFMA, 3 registers (two shared by all chains)", 8); run<3>("FFMA, 3 registers (all per chain)", 8);
    run<4>("FFMA, immediate multiplier", 8); run<5>("FSET.BF", 8);
    run<6>("loop mix: 14 FADD 11 FMUL 3 FFMA3 3 FFMAi 4 FSET", 35);
    run<7>("loop mix with FADD for the FFMA3", 35);
    run<8>("loop mix with FADD for the FSET", 35);
    return 0;
}

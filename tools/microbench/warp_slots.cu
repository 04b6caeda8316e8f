// Probe: where do the two warps of 64-thread CTAs land?  Prints, per SM, the hardware warp slot
// (%warpid) of every resident (cta, warp) of a 512-CTA grid whose CTAs stay resident.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_slots warp_slots.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(64, 8) probe(unsigned *out, int spin) {
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if ((threadIdx.x & 31) == 0) {
        unsigned w = threadIdx.x >> 5;
        out[(blockIdx.x * 2 + w) * 2 + 0] = smid;
        out[(blockIdx.x * 2 + w) * 2 + 1] = warpid;
    }
}
int main() {
    const int n = 512;
    unsigned *d, h[n * 4];
    cudaMalloc(&d, sizeof(h));
    probe<<<n, 64>>>(d, 2000000);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int per_sm[160] = {0};
    for (int sm = 0; sm < 6; sm++) {
        printf("SM %d:", sm);
        for (int c = 0; c < n; c++)
            if (h[(c * 2) * 2] == (unsigned)sm) printf("  cta %d -> slots %u,%u", c, h[(c * 2) * 2 + 1], h[(c * 2 + 1) * 2 + 1]);
        printf("\n");
    }
    int hist[4][2] = {{0}};
    for (int c = 0; c < n; c++)
        for (int w = 0; w < 2; w++) hist[h[(c * 2 + w) * 2 + 1] & 3][w]++;
    for (int s = 0; s < 4; s++) printf("SMSP %d: warp0 x%d, warp1 x%d\n", s, hist[s][0], hist[s][1]);
    for (int c = 0; c < n; c++) per_sm[h[c * 4]]++;
    int h3 = 0, h4 = 0, other = 0;
    for (int s = 0; s < 148; s++) (per_sm[s] == 3 ? h3 : per_sm[s] == 4 ? h4 : other)++;
    printf("SMs with 3 CTAs: %d, with 4: %d, other: %d\n", h3, h4, other);
    return 0;
}

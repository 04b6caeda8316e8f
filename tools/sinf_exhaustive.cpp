// sinf_exhaustive.cpp -- compares the two restatements of glibc's sinf in knaster_b200/csrc/sinf_glibc.h (host instantiation: the same
// f64 operations the device executes) with the libm this machine runs, for EVERY float |y| < 120: 2 x 0x42f00000 arguments.
// Test infrastructure (tests/test_sinf_exhaustive.py builds and runs it); prints one JSON line.
//   g++ -O2 -mfma -ffp-contract=off -pthread -Iknaster_b200/csrc tools/sinf_exhaustive.cpp -o sinf_exhaustive
// usage: sinf_exhaustive [stride]   (stride 1 = all arguments; k = every k-th bit pattern)
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "sinf_glibc.h"

int main(int argc, char **argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u;
    const uint32_t LIM = 0x42f00000u; // 120.0f
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> n_args{0}, bad_inrange{0}, bad_lean{0}, lean_zero_sign{0};
    std::vector<std::thread> th;
    for (unsigned w = 0; w < T; w++)
        th.emplace_back([&, w] {
            uint64_t n = 0, a = 0, b = 0, z = 0;
            for (uint64_t i = (uint64_t)w * stride; i < LIM; i += (uint64_t)T * stride)
                for (uint32_t sgn = 0; sgn < 2; sgn++) {
                    const uint32_t bits = (uint32_t)i | (sgn << 31);
                    float y;
                    memcpy(&y, &bits, 4);
                    const float ref = sinf(y), f1 = kgpu::kn_sinf_glibc_inrange(y), f2 = kgpu::kn_sinf_glibc_lean(y);
                    uint32_t rb, b1, b2;
                    memcpy(&rb, &ref, 4);
                    memcpy(&b1, &f1, 4);
                    memcpy(&b2, &f2, 4);
                    n++;
                    if (rb != b1) a++;
                    if (rb != b2) {
                        if (ref == f2) z++; // -0.0 against +0.0
                        else b++;
                    }
                }
            n_args += n, bad_inrange += a, bad_lean += b, lean_zero_sign += z;
        });
    for (auto &t : th) t.join();
    printf("{\"arguments\": %llu, \"stride\": %u, \"threads\": %u, \"inrange_mismatches\": %llu, \"lean_mismatches\": %llu, "
           "\"lean_zero_sign_only\": %llu, \"fma_cpu\": %s}\n",
           (unsigned long long)n_args.load(), stride, T, (unsigned long long)bad_inrange.load(), (unsigned long long)bad_lean.load(),
           (unsigned long long)lean_zero_sign.load(), __builtin_cpu_supports("fma") ? "true" : "false");
    return 0;
}

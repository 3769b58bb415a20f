// TEST INFRASTRUCTURE: m_sinf_glibc (termdaw_b200/csrc/tdg_math.cuh — the function the voice kernels evaluate on
// the device) against the host libm's sinf, over every f32 bit pattern in [lo, hi) with a stride.
//   sinf_check <stride> [threads]   -> prints "checked N mismatches M first 0x........"
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include "../termdaw_b200/csrc/tdg_math.cuh"

int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    const unsigned T = argc > 2 ? (unsigned)atoi(argv[2]) : std::max(1u, std::thread::hardware_concurrency());
    std::vector<uint64_t> bad(T, 0), cnt(T, 0), first(T, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            const uint64_t lo = ((uint64_t)t << 32) / T, hi = ((uint64_t)(t + 1) << 32) / T;
            uint64_t n_bad = 0, n_cnt = 0, f_bad = 0;                      // thread-local: no false sharing
            for (uint64_t u = lo + (stride - lo % stride) % stride; u < hi; u += stride) {
                const uint32_t v = (uint32_t)u;
                float f; memcpy(&f, &v, 4);
                const float a = tdg::m_sinf_glibc(f), b = sinf(f);
                ++n_cnt;
                if (tdg::m_f32_bits(a) != tdg::m_f32_bits(b) && !(a != a && b != b)) { if (!n_bad) f_bad = v; ++n_bad; }
            }
            bad[t] = n_bad; cnt[t] = n_cnt; first[t] = f_bad;
        });
    for (auto& x : th) x.join();
    uint64_t nb = 0, nc = 0, f0 = 0;
    for (unsigned t = 0; t < T; ++t) { nc += cnt[t]; if (bad[t] && !nb) f0 = first[t]; nb += bad[t]; }
    printf("checked %llu mismatches %llu first 0x%08llx\n", (unsigned long long)nc, (unsigned long long)nb, (unsigned long long)f0);
    return 0;
}

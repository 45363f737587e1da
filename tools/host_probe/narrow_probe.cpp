// Host-only probe (no GPU work): how fast do the library's host threads narrow an f32 raster to u16 DNs on this box, against
// the ~55 GB/s of f32 source the PCIe link would carry directly? Variants: plain stores, non-temporal stores, read-only pass.
// Build: g++ -O2 -mavx2 -std=c++17 -I../../sarpro_b200/csrc narrow_probe.cpp -o narrow_probe -lpthread
#include <immintrin.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "host_pool.h"

using namespace sarpro;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <int MODE> // 0 plain stores, 1 non-temporal stores, 2 read only
static void block(const float* s, uint16_t* d, size_t n, float thresh, unsigned* sink) {
    const __m256 vth = _mm256_set1_ps(thresh);
    const __m256i lim = _mm256_set1_epi32(65536);
    __m256i bad = _mm256_setzero_si256();
    for (size_t i = 0; i + 16 <= n; i += 16) {
        const __m256 a = _mm256_loadu_ps(s + i), b = _mm256_loadu_ps(s + i + 8);
        const __m256i ia = _mm256_cvttps_epi32(a), ib = _mm256_cvttps_epi32(b);
        const __m256i va = _mm256_castps_si256(_mm256_cmp_ps(a, vth, _CMP_GE_OQ)), vb = _mm256_castps_si256(_mm256_cmp_ps(b, vth, _CMP_GE_OQ));
        const __m256i oka = _mm256_and_si256(_mm256_castps_si256(_mm256_cmp_ps(a, _mm256_cvtepi32_ps(ia), _CMP_EQ_OQ)), _mm256_cmpgt_epi32(lim, ia));
        const __m256i okb = _mm256_and_si256(_mm256_castps_si256(_mm256_cmp_ps(b, _mm256_cvtepi32_ps(ib), _CMP_EQ_OQ)), _mm256_cmpgt_epi32(lim, ib));
        bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_andnot_si256(oka, va), _mm256_andnot_si256(okb, vb)));
        const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(_mm256_and_si256(ia, _mm256_and_si256(va, oka)), _mm256_and_si256(ib, _mm256_and_si256(vb, okb))), 0xD8);
        if (MODE == 0) _mm256_storeu_si256((__m256i*)(d + i), p);
        else if (MODE == 1) _mm256_stream_si256((__m256i*)(d + i), p);
        else bad = _mm256_or_si256(bad, _mm256_and_si256(p, _mm256_set1_epi32(1 << 30)));
    }
    if (MODE == 1) _mm_sfence();
    *sink |= (unsigned)_mm256_movemask_epi8(bad);
}

int main(int argc, char** argv) {
    const size_t n = (size_t)(argc > 1 ? atof(argv[1]) : 200e6);
    float* src = (float*)aligned_alloc(4096, n * 4);
    uint16_t* dst = (uint16_t*)aligned_alloc(4096, n * 2);
    WorkerPool& pool = WorkerPool::get();
    const size_t kBlock = 256 * 1024, nb = n / kBlock;
    pool.run((uint32_t)nb, 1, [&](uint32_t a, uint32_t b) { for (uint32_t k = a; k < b; ++k) { for (size_t i = 0; i < kBlock; ++i) src[k * kBlock + i] = (float)((k * 31 + i * 7) % 3000); std::memset(dst + k * kBlock, 0, kBlock * 2); } });
    unsigned sink = 0;
    printf("threads %u, %zu M samples\n", pool.width(), n / 1000000);
    for (int mode = 0; mode < 3; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            const double t0 = now();
            pool.run((uint32_t)nb, 1, [&](uint32_t a, uint32_t b) {
                unsigned s = 0;
                for (uint32_t k = a; k < b; ++k) {
                    if (mode == 0) block<0>(src + k * kBlock, dst + k * kBlock, kBlock, 1e-5f, &s);
                    else if (mode == 1) block<1>(src + k * kBlock, dst + k * kBlock, kBlock, 1e-5f, &s);
                    else block<2>(src + k * kBlock, dst + k * kBlock, kBlock, 1e-5f, &s);
                }
                __atomic_fetch_or(&sink, s, __ATOMIC_RELAXED);
            });
            const double dt = now() - t0;
            printf("mode %d (%s): %.1f GB/s of f32 source (%.2f ms)\n", mode, mode == 0 ? "plain stores" : mode == 1 ? "non-temporal stores" : "read only", nb * kBlock * 4 / dt * 1e-9, dt * 1e3);
        }
    { // one thread
        const double t0 = now();
        unsigned s = 0;
        for (size_t k = 0; k < nb / 8; ++k) block<1>(src + k * kBlock, dst + k * kBlock, kBlock, 1e-5f, &s);
        const double dt = now() - t0;
        printf("one thread, non-temporal: %.1f GB/s\n", (nb / 8) * kBlock * 4 / dt * 1e-9);
        sink |= s;
    }
    printf("sink %u\n", sink);
    return 0;
}

// narrow.cpp -- host side of the f32 boundary. The reference hands the library the f32 raster GDAL produced from a u16 GRD
// band (gdal.rs:123): 4 bytes per sample over PCIe for 2 bytes of information. For large host rasters the upload is the
// whole end-to-end time, so the samples are narrowed to their DN on the host, chunk by chunk into pinned staging, while the
// previous chunk is on the wire (api.cu: stage_band). The rule per sample is the one k_f32_to_dn applies on the device
// (kernels_small.cu): not >= valid_thresh (negative, NaN, below -50 dB: pipeline.rs:22) -> DN 0; a valid sample that is not a
// whole number <= 65535 -> the raster is not u16-valued, the caller falls back to the f32 upload and the general path.
#include <immintrin.h>

#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "host_pool.h"
#include "plan.h"

namespace sarpro {
namespace {

bool narrow_scalar(const float* s, uint16_t* d, size_t n, float thresh) {
    bool bad = false;
    for (size_t i = 0; i < n; ++i) {
        const float v = s[i];
        uint16_t o = 0;
        if (v >= thresh) {
            if (v > 65535.0f || v != std::trunc(v)) bad = true;
            else o = (uint16_t)v;
        }
        d[i] = o;
    }
    return !bad;
}

#if defined(__x86_64__)
// 16 samples per iteration. Whole-number test: truncate to i32 and convert back (NaN, infinities and |v| >= 2^31 come back as
// -2^31 and fail it); range test on the integers. The staging slots are far larger than the last-level cache and are read
// next by the DMA engine, not by a core: non-temporal stores skip the read-for-ownership of every destination line (93 instead
// of 79 GB/s of f32 source with 16 threads on the B200 host, profiles/r02B_narrow_probe.log).
template <bool STREAM>
__attribute__((target("avx2"))) bool narrow_avx2(const float* s, uint16_t* d, size_t n, float thresh) {
    const __m256 vth = _mm256_set1_ps(thresh);
    const __m256i lim = _mm256_set1_epi32(65536);
    __m256i bad = _mm256_setzero_si256();
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256 a = _mm256_loadu_ps(s + i), b = _mm256_loadu_ps(s + i + 8);
        const __m256i ia = _mm256_cvttps_epi32(a), ib = _mm256_cvttps_epi32(b);
        const __m256i va = _mm256_castps_si256(_mm256_cmp_ps(a, vth, _CMP_GE_OQ)); // false for NaN
        const __m256i vb = _mm256_castps_si256(_mm256_cmp_ps(b, vth, _CMP_GE_OQ));
        const __m256i oka = _mm256_and_si256(_mm256_castps_si256(_mm256_cmp_ps(a, _mm256_cvtepi32_ps(ia), _CMP_EQ_OQ)), _mm256_cmpgt_epi32(lim, ia));
        const __m256i okb = _mm256_and_si256(_mm256_castps_si256(_mm256_cmp_ps(b, _mm256_cvtepi32_ps(ib), _CMP_EQ_OQ)), _mm256_cmpgt_epi32(lim, ib));
        bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_andnot_si256(oka, va), _mm256_andnot_si256(okb, vb)));
        // packus works per 128-bit lane: [a0-3 b0-3 | a4-7 b4-7] -> quadwords reordered to a0-3 a4-7 b0-3 b4-7
        const __m256i p = _mm256_permute4x64_epi64(
            _mm256_packus_epi32(_mm256_and_si256(ia, _mm256_and_si256(va, oka)), _mm256_and_si256(ib, _mm256_and_si256(vb, okb))), 0xD8);
        if (STREAM) _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), p);
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + i), p);
    }
    if (STREAM) _mm_sfence(); // the copy engine reads the slot next
    const bool tail_ok = narrow_scalar(s + i, d + i, n - i, thresh);
    return _mm256_testz_si256(bad, bad) != 0 && tail_ok;
}
#endif

bool narrow_block(const float* s, uint16_t* d, size_t n, float thresh) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return (reinterpret_cast<uintptr_t>(d) & 31) == 0 ? narrow_avx2<true>(s, d, n, thresh) : narrow_avx2<false>(s, d, n, thresh);
#endif
    return narrow_scalar(s, d, n, thresh);
}

} // namespace

bool narrow_f32_to_dn(const float* src, uint16_t* dst, size_t n, float valid_thresh) {
    constexpr size_t kBlock = 256 * 1024; // samples per task: 1 MB read, 0.5 MB written
    const size_t n_blocks = (n + kBlock - 1) / kBlock;
    std::atomic<bool> ok{true};
    auto body = [&](uint32_t a, uint32_t b) {
        for (uint32_t k = a; k < b; ++k) {
            if (!ok.load(std::memory_order_relaxed)) return; // another block already found a sample that is not a DN
            const size_t o = (size_t)k * kBlock;
            if (!narrow_block(src + o, dst + o, std::min(kBlock, n - o), valid_thresh)) ok.store(false, std::memory_order_relaxed);
        }
    };
    if (n_blocks <= 2) body(0, (uint32_t)n_blocks);
    else WorkerPool::get().run((uint32_t)n_blocks, 1, std::function<void(uint32_t, uint32_t)>(body));
    return ok.load();
}

} // namespace sarpro

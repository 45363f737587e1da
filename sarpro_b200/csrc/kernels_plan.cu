// kernels_plan.cu — the band planner on the device (one CTA), for the strategies whose window needs no pow():
// Robust / Equalized / Clahe / Tamed / Default and the two Tamed-synRGB kinds (gamma == 1, autoscale.rs:492-561, 721-736).
//
// It is plan.cpp's plan_finish() restated for a single thread block: the same f64 operations in the same order on the
// same inputs, so every percentile, the window and the DN -> sample / bin table come out bit-identical to the host
// planner (the tests compare the two on every fixture). What makes that possible:
//   * the DN -> dB table (pipeline.rs:19-20) is data-independent: the host computes it once per context with the libm the
//     reference uses and uploads it (512 KB); no transcendental runs here;
//   * everything else in compute_histogram_stats (autoscale.rs:35-160) is +, -, *, /, floor and comparisons on f64, all
//     IEEE-exact on the device when nothing is contracted (explicit __d*_rn intrinsics below; the build also passes --fmad=false);
//   * gamma == 1: powf(x, 1.0) == x, so the quantisation (autoscale.rs:647-651) is a multiply and a truncating cast;
//   * scale_u16_to_u8 (autoscale.rs:348-364) is f32 arithmetic with roundf — the same on both sides.
// Not bit-identical to the reference (nor was the host planner): mean / std (serial Welford over pixels, order-dependent);
// they only feed log lines on these strategies. Standard and Adaptive (pow with gamma != 1, the skew test on mean / std)
// stay on the host planner.
//
// With the plan on the device no host round trip is left between pass A and pass B: the kernels downstream read the table
// range (`hot`), the brightest present DN and the kernel choice from PlanDev.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace sarpro {

namespace {

constexpr uint32_t kPlanThreads = 1024;
constexpr uint32_t kDnPerThread = 65536 / kPlanThreads; // 64 DNs per thread, interleaved: DN k * 1024 + tid (coalesced loads)
constexpr int kStatBinsDev = 4096;

__device__ __forceinline__ unsigned long long cast_u64_dev(double x) { // Rust `as u64`: truncate, saturate, NaN -> 0
    if (!(x == x) || x <= 0.0) return 0ull;
    if (x >= 18446744073709551616.0) return 0xffffffffffffffffull;
    return (unsigned long long)x;
}
__device__ __forceinline__ uint32_t cast_u16_dev(double x) {
    if (!(x == x) || x <= 0.0) return 0u;
    if (x >= 65535.0) return 65535u;
    return (uint32_t)x;
}
__device__ __forceinline__ uint32_t cast_u8_dev(double x) {
    if (!(x == x) || x <= 0.0) return 0u;
    if (x >= 255.0) return 255u;
    return (uint32_t)x;
}
__device__ __forceinline__ double clampd_dev(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// block-wide reductions over 1024 threads (32 warps); every thread gets the result
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* scratch /* 33 entries */) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads(); // scratch may still be read from the previous reduction
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        T w = scratch[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w = op(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (lane == 0) scratch[32] = w;
    }
    __syncthreads();
    return scratch[32];
}
struct OpAddU64 { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a + b; } };
struct OpMinU32 { __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a < b ? a : b; } };
struct OpMaxU32 { __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };
struct OpMaxI32 { __device__ int operator()(int a, int b) const { return a > b ? a : b; } };
struct OpAddF64 { __device__ double operator()(double a, double b) const { return __dadd_rn(a, b); } };

} // namespace

// Entries of the present DNs, compacted by the kernel itself: in shared memory when there are at most kPlanCap of them (a GRD
// band has a few thousand), else in a global scratch buffer (every u16 value present: same code, slower).
constexpr uint32_t kPlanCap = 8192;
constexpr size_t kPlanDynSmem = (size_t)kPlanCap * (8 + 4 + 2 + 2) + 4096 * 8; // entries + the prefix sums of the stat histogram
constexpr size_t kPlanScratchBytes = (size_t)65536 * (8 + 4 + 2 + 2);

__global__ void __launch_bounds__(kPlanThreads, 1)
k_plan_band(PlanJobs jobs, const double* __restrict__ db) {
    // one CTA per band (blockIdx.x): the two bands of a pair can be planned by one launch
    const uint32_t* __restrict__ total = jobs.j[blockIdx.x].total;
    uint16_t* __restrict__ lut = jobs.j[blockIdx.x].lut;
    PlanDev* __restrict__ out = jobs.j[blockIdx.x].out;
    unsigned char* __restrict__ scratch = reinterpret_cast<unsigned char*>(jobs.j[blockIdx.x].scratch);
    const PlanParams pr = jobs.j[blockIdx.x].params;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    unsigned long long* const s_cum = reinterpret_cast<unsigned long long*>(s_dyn + (size_t)kPlanCap * 16); // inclusive prefix sums
    __shared__ uint32_t s_hist[kStatBinsDev]; // 4096-bin histogram (autoscale.rs:103-117); a bin holds < 2^32 pixels (rasters are below 2^32 samples)
    __shared__ unsigned long long s_u64[33];
    __shared__ uint32_t s_u32[33];
    __shared__ int s_i32[33];
    __shared__ double s_f64[33];
    __shared__ double s_pct[11];
    __shared__ uint8_t s_remap[256];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;

    // ---- the present DNs, compacted -------------------------------------------------------------------------------------
    // Thread t looks at DNs k * 1024 + t (coalesced, 64 independent loads), keeps a presence mask, and the block scans the
    // per-thread counts into offsets. The list is not in DN order; nothing below depends on the order (sums of products
    // run in list order, which is fixed, so the result is deterministic).
    unsigned long long present_bits = 0;
    uint32_t hh[kDnPerThread];
#pragma unroll
    for (uint32_t k = 0; k < kDnPerThread; ++k) hh[k] = total[k * kPlanThreads + tid];
#pragma unroll
    for (uint32_t k = 0; k < kDnPerThread; ++k) present_bits |= hh[k] ? (1ull << k) : 0ull;
    const uint32_t mine = (uint32_t)__popcll(present_bits);
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t nb = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += nb;
    }
    if (lane == 31) s_u32[wid] = inc;
    __syncthreads();
    uint32_t warp_off = 0, n_present = 0;
    for (uint32_t w = 0; w < 32; ++w) {
        const uint32_t c = s_u32[w];
        if (w < wid) warp_off += c;
        n_present += c;
    }
    __syncthreads();
    const bool in_smem = n_present <= kPlanCap;
    unsigned char* base = in_smem ? s_dyn : scratch;
    const size_t cap = in_smem ? kPlanCap : 65536;
    double* e_db = reinterpret_cast<double*>(base);
    uint32_t* e_h = reinterpret_cast<uint32_t*>(base + cap * 8);
    uint16_t* e_dn = reinterpret_cast<uint16_t*>(base + cap * 12);
    uint16_t* e_q = reinterpret_cast<uint16_t*>(base + cap * 14);
    {
        uint32_t pos = warp_off + inc - mine;
#pragma unroll
        for (uint32_t k = 0; k < kDnPerThread; ++k)
            if (hh[k]) {
                e_dn[pos] = (uint16_t)(k * kPlanThreads + tid);
                e_h[pos] = hh[k];
                ++pos;
            }
    }
    // every table word starts at 0: absent and invalid DNs keep it (invalid pixels are written as 0, autoscale.rs:444, 653, 738)
    {
        uint4* lut4 = reinterpret_cast<uint4*>(lut);
        for (uint32_t i = tid; i < 65536u * 2u / 16u; i += kPlanThreads) lut4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    for (uint32_t i = tid; i < n_present; i += kPlanThreads) e_db[i] = db[e_dn[i]];
    __syncthreads();

    // valid <=> dB > -50 (pipeline.rs:22); for a u16 DN that is DN >= 1 (dB(1) = 0, dB(0) = -100), but the table decides.
    unsigned long long cnt = 0, px = 0, ge1k = 0, ge2k = 0;
    uint32_t first_valid = 0xffffffffu, last_valid = 0, last_present = 0;
    uint32_t invalid_present = 0;
    for (uint32_t i = tid; i < n_present; i += kPlanThreads) {
        const uint32_t d = e_dn[i], h = e_h[i];
        px += h;
        if (d >= 1024) ge1k += h;
        if (d >= 2048) ge2k += h;
        last_present = max(last_present, d);
        if (e_db[i] > -50.0) {
            cnt += h;
            first_valid = min(first_valid, d);
            last_valid = max(last_valid, d);
        } else {
            invalid_present = 1;
        }
    }
    // one combined block reduction (two barriers) instead of eight: a single-CTA kernel is paced by its barriers
    __shared__ unsigned long long s_r64[4][32];
    __shared__ uint32_t s_r32[4][32];
    __shared__ unsigned long long s_o64[4];
    __shared__ uint32_t s_o32[4];
    {
        unsigned long long a[4] = {cnt, px, ge1k, ge2k};
        uint32_t m[4] = {first_valid, ~last_valid, ~last_present, ~invalid_present}; // all as minima
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
                m[j] = min(m[j], __shfl_xor_sync(0xffffffffu, m[j], o));
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { s_r64[j][wid] = a[j]; s_r32[j][wid] = m[j]; }
        }
        __syncthreads();
        if (wid < 4) { // warp j finishes value j of both groups
            unsigned long long x = s_r64[wid][lane];
            uint32_t y = s_r32[wid][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                x += __shfl_xor_sync(0xffffffffu, x, o);
                y = min(y, __shfl_xor_sync(0xffffffffu, y, o));
            }
            if (lane == 0) { s_o64[wid] = x; s_o32[wid] = y; }
        }
        __syncthreads();
    }
    const unsigned long long count = s_o64[0], px_total = s_o64[1], px_ge1024 = s_o64[2], px_ge2048 = s_o64[3];
    const uint32_t min_dn = s_o32[0], max_valid_dn = ~s_o32[1], max_present_dn = ~s_o32[2];
    const bool have_invalid = (~s_o32[3]) != 0;

    if (count == 0) { // all-zero output (autoscale.rs:376-378, 466-468, 716-718): every table word is 0
        if (tid == 0) {
            PlanDev p{};
            p.max_present_dn = max_present_dn;
            p.have_invalid = have_invalid;
            p.hot = 64; // every present DN (only invalid ones) carries word 0
            p.hot_top = 0;
            p.clahe = pr.clahe;
            p.px_total = px_total;
            p.px_ge1024 = px_ge1024;
            p.px_ge2048 = px_ge2048;
            *out = p;
        }
        return;
    }

    // ---- pass 1 of compute_histogram_stats over distinct values (autoscale.rs:37-55) ----------------------------------
    const double min_db = db[min_dn], max_db = db[max_valid_dn]; // dB is monotone in the DN
    double sum = 0.0;
    for (uint32_t i = tid; i < n_present; i += kPlanThreads)
        if (e_db[i] > -50.0) sum = __dadd_rn(sum, __dmul_rn((double)e_h[i], e_db[i]));
    const double mean_db = __ddiv_rn(block_reduce(sum, OpAddF64(), s_f64), (double)count);
    double m2 = 0.0;
    for (uint32_t i = tid; i < n_present; i += kPlanThreads)
        if (e_db[i] > -50.0) {
            const double dd = __dsub_rn(e_db[i], mean_db);
            m2 = __dadd_rn(m2, __dmul_rn((double)e_h[i], __dmul_rn(dd, dd)));
        }
    const double m2_all = block_reduce(m2, OpAddF64(), s_f64);
    const double std_db = count > 1 ? sqrt(__ddiv_rn(m2_all, (double)count)) : 0.0;

    // ---- pass 2 (autoscale.rs:103-117) + percentiles (:120-159) ---------------------------------------------------------
    const bool degenerate = fabs(__dsub_rn(max_db, min_db)) < DBL_EPSILON; // autoscale.rs:81
    if (!degenerate) {
        for (uint32_t i = tid; i < (uint32_t)kStatBinsDev; i += kPlanThreads) s_hist[i] = 0u;
        __syncthreads();
        const double span = __dsub_rn(max_db, min_db);
        const double inv_span = __ddiv_rn(1.0, span);
        for (uint32_t i = tid; i < n_present; i += kPlanThreads)
            if (e_db[i] > -50.0) {
                const double t = clampd_dev(__dmul_rn(__dsub_rn(e_db[i], min_db), inv_span), 0.0, 1.0);
                unsigned long long idx = cast_u64_dev(__dmul_rn(t, (double)kStatBinsDev));
                if (idx >= (unsigned long long)kStatBinsDev) idx = kStatBinsDev - 1;
                atomicAdd(&s_hist[idx], e_h[i]);
            }
        __syncthreads();
        // inclusive prefix sums in place: four bins per thread, then the block-wide offsets
        unsigned long long v[4], run = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { run += s_hist[tid * 4 + i]; v[i] = run; }
        {
            unsigned long long inc64 = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long nb = __shfl_up_sync(0xffffffffu, inc64, o);
                if ((int)lane >= o) inc64 += nb;
            }
            __syncthreads();
            if (lane == 31) s_u64[wid] = inc64;
            __syncthreads();
            unsigned long long woff = 0;
            for (uint32_t w = 0; w < wid; ++w) woff += s_u64[w];
            const unsigned long long off = woff + inc64 - run;
#pragma unroll
            for (int i = 0; i < 4; ++i) s_cum[tid * 4 + i] = v[i] + off;
        }
        __syncthreads();
        if (tid < 11) { // estimate_percentile (autoscale.rs:120-140) for p = 1, 2, 5, 10, 25, 50, 75, 90, 95, 98, 99 %
            const double ps[11] = {0.01, 0.02, 0.05, 0.10, 0.25, 0.5, 0.75, 0.90, 0.95, 0.98, 0.99};
            unsigned long long target = cast_u64_dev(floor(__dmul_rn(ps[tid], (double)count)));
            if (target >= count) target = count - 1;
            // first bin b with target < cumsum(b) (the inclusive sums are non-decreasing)
            int lo = 0, hi = kStatBinsDev - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_cum[mid] > target) hi = mid; else lo = mid + 1;
            }
            const int b = lo;
            const unsigned long long before = b ? s_cum[b - 1] : 0ull;
            const unsigned long long h = s_cum[b] - before;
            const unsigned long long within = target >= before ? target - before : 0ull;
            const double frac = h > 0 ? __ddiv_rn((double)within, (double)h) : 0.0;
            const double bin_width = __ddiv_rn(span, (double)kStatBinsDev);
            const double bin_start = __dadd_rn(min_db, __dmul_rn((double)b, bin_width));
            s_pct[tid] = __dadd_rn(bin_start, __dmul_rn(frac, bin_width));
        }
    } else if (tid < 11) {
        s_pct[tid] = tid <= 5 ? min_db : max_db; // autoscale.rs:81-100: p01..p25 and the median = min, p75..p99 = max
    }
    __syncthreads();
    const double p01 = s_pct[0], p02 = s_pct[1], p05 = s_pct[2], p10 = s_pct[3], p25 = s_pct[4], med = s_pct[5], p75 = s_pct[6],
                 p90 = s_pct[7], p95 = s_pct[8], p98 = s_pct[9], p99 = s_pct[10];

    // ---- window (plan.cpp choose_window; only the gamma == 1 arms) ------------------------------------------------------
    double low, high; // (every thread computes the window: cheaper than a barrier)
    if (pr.kind == 1) { // TamedSynRgbCopol, autoscale.rs:721-723
        low = fmin(p02, p05);
        high = p99;
    } else if (pr.kind == 2) { // TamedSynRgbCross, autoscale.rs:724-727
        low = p05;
        high = p99;
    } else if (pr.strategy == SARPRO_STRATEGY_ROBUST) { // autoscale.rs:492-499
        const double iqr = __dsub_rn(p75, p25);
        const double thr = __dmul_rn(2.5, iqr);
        low = fmax(fmax(__dsub_rn(p25, thr), p01), min_db);
        high = fmin(fmin(__dadd_rn(p75, thr), p99), max_db);
    } else if (pr.strategy == SARPRO_STRATEGY_EQUALIZED || pr.strategy == SARPRO_STRATEGY_CLAHE) { // :539-548
        low = p01;
        high = p99;
    } else if (pr.strategy == SARPRO_STRATEGY_TAMED) { // :549-553
        low = p25;
        high = p99;
    } else { // Default, :554-561
        low = p05;
        high = p95;
    }
    const double range = fmax(__dsub_rn(high, low), 1.0); // autoscale.rs:429, 564, 729

    // ---- the table ------------------------------------------------------------------------------------------------------
    const bool tamed_rgb = pr.kind != 0;
    const bool clahe = !tamed_rgb && pr.strategy == SARPRO_STRATEGY_CLAHE;
    const double max_val = (tamed_rgb || pr.bit_depth == SARPRO_U8) ? 255.0 : 65535.0;
    uint32_t mn = 65535u, mx = 0u;
    for (uint32_t i = tid; i < n_present; i += kPlanThreads) {
        uint32_t w = 0;
        if (e_db[i] > -50.0) {
            const double v = e_db[i];
            const double a = v < low ? low : v;           // v.max(low).min(high), autoscale.rs:440, 583, 649, 734
            const double clipped = a > high ? high : a;
            const double n = __ddiv_rn(__dsub_rn(clipped, low), range);
            if (clahe) { // autoscale.rs:585-587 then the bin of :263 / :320
                const double c = clampd_dev(n, 0.0, 1.0);
                const double s = __dmul_rn(c, 255.0);
                const double t = (double)(int)s;
                double b = (__dsub_rn(s, t) >= 0.5) ? __dadd_rn(t, 1.0) : t; // f64::round for 0 <= s <= 255
                long long bin = (b == b) ? (long long)b : 0;
                if (bin < 0) bin = 0;
                if (bin > 255) bin = 255;
                w = (uint32_t)bin;
            } else if (tamed_rgb) { // autoscale.rs:734-736
                w = cast_u8_dev(clampd_dev(__dmul_rn(n, 255.0), 0.0, 255.0));
            } else {                // autoscale.rs:649-651 with gamma == 1 (powf(x, 1) == x)
                w = cast_u16_dev(clampd_dev(__dmul_rn(n, max_val), 0.0, max_val));
            }
            mn = min(mn, w);
            mx = max(mx, w);
        }
        e_q[i] = (uint16_t)w;
    }
    uint32_t pre_min = 0, pre_max = 0;
    if (!clahe) {
        if (have_invalid) { mn = 0; } // invalid pixels are samples of value 0
        {   // {min, max} in one reduction: 16-bit values, packed as (mn << 16) | (0xffff - mx), minimum of both halves
            uint32_t a = mn, b = 0xffffu - mx;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
                b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
            }
            if (lane == 0) { s_r32[0][wid] = a; s_r32[1][wid] = b; }
            __syncthreads();
            if (wid < 2) {
                uint32_t y = s_r32[wid][lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) y = min(y, __shfl_xor_sync(0xffffffffu, y, o));
                if (lane == 0) s_o32[wid] = y;
            }
            __syncthreads();
            pre_min = s_o32[0];
            pre_max = 0xffffu - s_o32[1];
        }
        if (!tamed_rgb && pr.bit_depth == SARPRO_U8) { // scale_u16_to_u8 over ALL pixels (autoscale.rs:352-363, :669-670, :691-693)
            if (tid < 256) {
                const float fmn = (float)pre_min, fmx = (float)pre_max;
                const float scale = fmx > fmn ? __fdiv_rn(255.0f, __fsub_rn(fmx, fmn)) : 1.0f;
                float val = roundf(__fmul_rn(__fsub_rn((float)tid, fmn), scale));
                val = val < 0.0f ? 0.0f : (val > 255.0f ? 255.0f : val);
                s_remap[tid] = (uint8_t)val;
            }
            __syncthreads();
            for (uint32_t i = tid; i < n_present; i += kPlanThreads)
                if (e_db[i] > -50.0) e_q[i] = s_remap[e_q[i] > 255u ? 255u : e_q[i]];
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < n_present; i += kPlanThreads) lut[e_dn[i]] = e_q[i];

    // ---- table range of the tensor-core pass B (plan.cpp set_sat_from, api.cu hmma_hot_from_plan) -----------------------
    // sat_from = the lowest valid present DN from which every valid present DN up to the brightest present one carries the
    // brightest one's table word (low byte).
    uint32_t top_word = 0;
    for (uint32_t i = tid; i < n_present; i += kPlanThreads)
        if (e_dn[i] == max_present_dn) s_u32[0] = e_q[i] & 255u;
    __syncthreads();
    top_word = s_u32[0];
    __syncthreads();
    int last_nontop = -1;
    for (uint32_t i = tid; i < n_present; i += kPlanThreads)
        if (e_db[i] > -50.0 && (e_q[i] & 255u) != top_word) last_nontop = max(last_nontop, (int)e_dn[i]);
    last_nontop = block_reduce(last_nontop, OpMaxI32(), s_i32);
    uint32_t first_after = 0xffffffffu;
    for (uint32_t i = tid; i < n_present; i += kPlanThreads)
        if (e_db[i] > -50.0 && (int)e_dn[i] > last_nontop) first_after = min(first_after, (uint32_t)e_dn[i]);
    first_after = block_reduce(first_after, OpMinU32(), s_u32);
    if (tid == 0) {
        uint32_t h = first_after == 0xffffffffu ? max_present_dn : first_after;
        if (last_nontop < 0 && have_invalid && top_word == 0) h = 0; // invalid DNs are present with word 0
        const uint32_t need = max(64u, (h + 1u + 7u) & ~7u);
        PlanDev p{};
        p.any_valid = 1;
        p.have_invalid = have_invalid;
        p.max_present_dn = max_present_dn;
        p.sat_from_dn = h;
        p.hot = need <= kHmmaMaxHot ? need : 0u;
        p.hot_top = top_word;
        p.use_generic = p.hot == 0;
        p.clahe = clahe;
        p.pre_min = pre_min;
        p.pre_max = pre_max;
        p.px_total = px_total;
        p.px_ge1024 = px_ge1024;
        p.px_ge2048 = px_ge2048;
        sarpro_stats& st = p.stats;
        st.valid_count = count;
        st.min_db = min_db;
        st.max_db = max_db;
        st.mean_db = mean_db;
        st.std_db = std_db;
        st.median_db = med;
        st.p01 = p01; st.p02 = p02; st.p05 = p05; st.p10 = p10; st.p25 = p25;
        st.p75 = p75; st.p90 = p90; st.p95 = p95; st.p98 = p98; st.p99 = p99;
        st.low_clip = low;
        st.high_clip = high;
        st.gamma = 1.0;
        *out = p;
    }
}

bool plan_on_device_supported(int strategy, int kind) {
    if (kind != 0) return true; // Tamed-synRGB windows (autoscale.rs:721-727)
    return strategy == SARPRO_STRATEGY_ROBUST || strategy == SARPRO_STRATEGY_EQUALIZED || strategy == SARPRO_STRATEGY_CLAHE ||
           strategy == SARPRO_STRATEGY_TAMED || strategy == SARPRO_STRATEGY_DEFAULT;
}

size_t plan_scratch_bytes() { return kPlanScratchBytes; }

cudaError_t launch_plan_bands(const PlanJobs& jobs, int n_bands, const double* db_table, cudaStream_t stream) {
    if (n_bands < 1 || n_bands > 2) return cudaErrorInvalidValue;
    if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_plan_band), kPlanDynSmem)) return e;
    k_plan_band<<<n_bands, kPlanThreads, kPlanDynSmem, stream>>>(jobs, db_table);
    return cudaGetLastError();
}

} // namespace sarpro

// kernels_f32.cu — rasters whose samples are not u16-valued: polarization ratios / normalised
// differences (ops.rs:10-44) and calibrated f32 inputs. The op is fused into every loader, so the
// combined f32 plane of the reference (a new Array2<f32>, ops.rs) is never written. Every kernel takes
// ONE or TWO operations over the same operand pair (NOPS): the two-band products of a pair (BASELINE config 4:
// log-ratio and normalised difference of VV / VH) read the operands once per pass instead of once per band.
//
// Exactness without device transcendentals: every index the reference derives from a sample
//   4096-bin stat index  autoscale.rs:113-116      quantised level  autoscale.rs:440-442 / 649-651
//   CLAHE bin            autoscale.rs:263
// is a monotone non-decreasing function of the sample value, so the host (plan_f32.cpp) converts each
// index boundary into an f32 threshold with the same libm the reference uses, and the device only
// compares: index(v) = #{k : v >= edge[k]}. A fast __log2f-based guess lands within a step or two of the
// answer; the two correction loops make the result independent of the guess.
//
// Loads: a thread takes 8 consecutive samples (one 128-bit load per u16 operand, two per f32 operand) when the
// operand pointers are 16-byte aligned; the last partial vector and unaligned rasters take the scalar loader.
#include "common.cuh"
#include "kernels.h"

namespace sarpro {

__device__ __forceinline__ float f32_pol(int op, float a, float b) {
    switch (op) {
    case 0: return __fadd_rn(a, b);
    case 1: return __fsub_rn(a, b);
    case 3: {
        const float denom = __fadd_rn(a, b);
        return fabsf(denom) > 1e-10f ? __fdiv_rn(__fsub_rn(a, b), denom) : 0.0f;
    }
    default: return fabsf(b) > 1e-10f ? __fdiv_rn(a, b) : 0.0f;
    }
}

__device__ __forceinline__ uint4 f32_ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// 8 consecutive samples of one operand, starting at element i (i % 8 == 0); cnt of them exist
__device__ __forceinline__ void f32_load8(const void* p, int is_u16, uint64_t i, uint32_t cnt, bool vec, float (&x)[8]) {
    if (vec && cnt == 8) {
        if (is_u16) {
            const uint4 w = f32_ldg_stream(reinterpret_cast<const uint16_t*>(p) + i);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                x[2 * j] = (float)(ww[j] & 0xffffu);
                x[2 * j + 1] = (float)(ww[j] >> 16);
            }
        } else {
            const uint4 w0 = f32_ldg_stream(reinterpret_cast<const float*>(p) + i);
            const uint4 w1 = f32_ldg_stream(reinterpret_cast<const float*>(p) + i + 4);
            x[0] = __uint_as_float(w0.x); x[1] = __uint_as_float(w0.y); x[2] = __uint_as_float(w0.z); x[3] = __uint_as_float(w0.w);
            x[4] = __uint_as_float(w1.x); x[5] = __uint_as_float(w1.y); x[6] = __uint_as_float(w1.z); x[7] = __uint_as_float(w1.w);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[k] = 0.0f;
            if ((uint32_t)k < cnt) x[k] = is_u16 ? (float)reinterpret_cast<const uint16_t*>(p)[i + k] : reinterpret_cast<const float*>(p)[i + k];
        }
    }
}

struct F32Src {
    const void* a;
    const void* b;
    int a_u16, b_u16;
    int op[2];   // op[1] only read by the NOPS == 2 kernels
    int vec;     // operand pointers are 16-byte aligned
    // samples of vector v (elements 8v .. 8v+7) for every operation; returns how many of them exist
    template <int NOPS>
    __device__ __forceinline__ uint32_t get8(uint64_t v, uint64_t n, float (&out)[NOPS][8]) const {
        const uint64_t i = v * 8;
        const uint32_t cnt = (uint32_t)min((uint64_t)8, n - i);
        float x[8], y[8];
        f32_load8(a, a_u16, i, cnt, vec != 0, x);
        const bool need_b = op[0] >= 0 || (NOPS == 2);
        if (need_b) f32_load8(b, b_u16, i, cnt, vec != 0, y);
        // the operation is warp-uniform: one switch per vector and operation, straight-line arithmetic inside
#pragma unroll
        for (int o = 0; o < NOPS; ++o) {
            switch (op[o]) {
            case 0:
#pragma unroll
                for (int k = 0; k < 8; ++k) out[o][k] = __fadd_rn(x[k], y[k]);
                break;
            case 1:
#pragma unroll
                for (int k = 0; k < 8; ++k) out[o][k] = __fsub_rn(x[k], y[k]);
                break;
            case 2:
            case 4: // log-ratio == ratio (ops.rs:35-44)
#pragma unroll
                for (int k = 0; k < 8; ++k) out[o][k] = fabsf(y[k]) > 1e-10f ? __fdiv_rn(x[k], y[k]) : 0.0f;
                break;
            case 3:
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float denom = __fadd_rn(x[k], y[k]);
                    out[o][k] = fabsf(denom) > 1e-10f ? __fdiv_rn(__fsub_rn(x[k], y[k]), denom) : 0.0f;
                }
                break;
            default:
#pragma unroll
                for (int k = 0; k < 8; ++k) out[o][k] = x[k];
                break;
            }
        }
        return cnt;
    }
};

static F32Src make_src(const void* a, const void* b, int a_u16, int b_u16, int op0, int op1) {
    F32Src s{a, b, a_u16, b_u16, {op0, op1}, 0};
    const bool need_b = op0 >= 0 || op1 >= 0;
    s.vec = (reinterpret_cast<uintptr_t>(a) & 15u) == 0 && (!need_b || (reinterpret_cast<uintptr_t>(b) & 15u) == 0);
    return s;
}

// ---- pass 1: min / max / count over valid samples (autoscale.rs:37-55) --------------------------------
template <int NOPS>
__global__ void __launch_bounds__(256) k_f32_scan(F32Src src, uint64_t n, float valid_thresh, F32Scan* __restrict__ out) {
    uint32_t mn[NOPS], mx[NOPS], cnt[NOPS];
#pragma unroll
    for (int o = 0; o < NOPS; ++o) { mn[o] = 0xffffffffu; mx[o] = 0; cnt[o] = 0; }
    const uint64_t nvec = (n + 7) / 8, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        float s[NOPS][8];
        const uint32_t c = src.get8<NOPS>(v, n, s);
#pragma unroll
        for (int o = 0; o < NOPS; ++o)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((uint32_t)k < c && s[o][k] >= valid_thresh) { // valid samples are positive: the bit pattern orders like the value
                    const uint32_t key = __float_as_uint(s[o][k]);
                    mn[o] = min(mn[o], key);
                    mx[o] = max(mx[o], key);
                    cnt[o]++; // < 2^32 per thread (n < 2^35)
                }
    }
#pragma unroll
    for (int o = 0; o < NOPS; ++o) {
        const uint32_t a = warp_reduce_min(mn[o]), b = warp_reduce_max(mx[o]);
        unsigned long long c = cnt[o];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if ((threadIdx.x & 31) == 0 && c) {
            atomicMin(&out[o].min_key, a);
            atomicMax(&out[o].max_key, b);
            atomicAdd(&out[o].valid_count, c);
        }
    }
}
cudaError_t launch_f32_scan(const void* a, const void* b, int a_u16, int b_u16, int op0, int op1, int nops, uint64_t n,
                            float valid_thresh, F32Scan* out, int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const F32Src src = make_src(a, b, a_u16, b_u16, op0, nops == 2 ? op1 : -1);
    if (nops == 2) k_f32_scan<2><<<sm_count * 8, 256, 0, stream>>>(src, n, valid_thresh, out);
    else k_f32_scan<1><<<sm_count * 8, 256, 0, stream>>>(src, n, valid_thresh, out);
    return cudaGetLastError();
}

// index(v) = #{k in [1, n_edges] : v >= edges[k]}; edges[0] unused, edges ascending (ties allowed)
__device__ __forceinline__ uint32_t edge_index(const float* __restrict__ edges, uint32_t n_edges, float v, int guess) {
    int g = guess < 0 ? 0 : (guess > (int)n_edges ? (int)n_edges : guess);
    while (g < (int)n_edges && v >= edges[g + 1]) ++g;
    while (g > 0 && v < edges[g]) --g;
    return (uint32_t)g;
}

// Guarded direct index. For the linear-in-dB index functions (stat bins; quantised levels with gamma == 1) the position
//   t = (log2(v) - y_lo) * scale,   y_lo = low_dB / (10 log10 2),   scale = n * 10 log10 2 / range_dB
// is evaluated in fp32 with a proven error bound E (host: f32_guard): log2(v) = e + lg2(m) with the exponent e exact and
// MUFU.LG2 on the mantissa m in [1, 2) (absolute error <= 2^-22), y_lo split into integer and fraction so that the integer
// parts cancel exactly. When frac(t) lies in [E, 1 - E] the reference's f64 index is floor(t) and no threshold is read; the
// other samples (a share of 2E) take the exact threshold comparison. E >= 0.5 switches the shortcut off.
struct F32Guard {
    int e0;        // floor(y_lo)
    float f0;      // y_lo - floor(y_lo)
    float scale;   // n * 10 log10(2) / range
    float guard;   // E
};
__device__ __forceinline__ float f32_split_log2(float x, int* e) {
    const uint32_t b = __float_as_uint(x); // positive, normal (>= valid_thresh)
    *e = (int)(b >> 23) - 127;
    return __log2f(__uint_as_float((b & 0x7fffffu) | 0x3f800000u));
}
// Returns true and the index when the shortcut decides. The reference's index is clamp(floor(t*), 0, top) for the exact
// position t* (stat bins: top = 4095 with n = 4096; levels: top = n). With |t - t*| <= E < guard, a fraction of t that keeps
// `guard` away from both integers puts t* strictly inside (floor(t), floor(t) + 1): floor(t*) == floor(t), in range or not.
__device__ __forceinline__ bool f32_guarded_index(const F32Guard& gd, int e, float lg, uint32_t top, uint32_t* idx) {
    const float d = __fadd_rn((float)(e - gd.e0), __fsub_rn(lg, gd.f0));
    const float t = __fmul_rn(d, gd.scale);
    const float fl = floorf(t), fr = t - fl; // exact
    *idx = (uint32_t)min(max((int)fl, 0), (int)top); // (the conversion saturates)
    return fabsf(fr - 0.5f) <= 0.5f - gd.guard;      // false for NaN and for |t| >= 2^23 (fr == 0)
}

// ---- pass 2: 4096-bin histogram over [min_db, max_db] + mean / M2 accumulators -----------------------
struct F32HistArgs {
    float valid_thresh;
    float min_db[2], inv_span4096[2]; // guess: (db - min_db) * inv_span * 4096
    const float* edges[2];            // [4096]: edges[k], k = 1..4095
    unsigned long long* hist[2];      // [4096]
    double* sums[2];                  // [0] = sum(db - min_db), [1] = sum((db - min_db)^2)   (fp32 logs, f64 accumulation)
    F32Guard guard[2];
};
template <int NOPS>
__global__ void __launch_bounds__(256) k_f32_hist4096(F32Src src, uint64_t n, F32HistArgs h) {
    extern __shared__ uint4 f32_smem[];
    float* s_edges = reinterpret_cast<float*>(f32_smem);                 // [NOPS][4096]
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_edges + NOPS * 4096); // [NOPS][4096]
    __shared__ double s_red[2][2][8];
#pragma unroll
    for (int o = 0; o < NOPS; ++o)
        for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x) {
            s_edges[o * 4096 + i] = h.edges[o][i];
            s_hist[o * 4096 + i] = 0;
        }
    __syncthreads();
    double s1[NOPS], s2[NOPS];
#pragma unroll
    for (int o = 0; o < NOPS; ++o) { s1[o] = 0.0; s2[o] = 0.0; }
    const uint64_t nvec = (n + 7) / 8, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        float s[NOPS][8];
        const uint32_t c = src.get8<NOPS>(v, n, s);
#pragma unroll
        for (int o = 0; o < NOPS; ++o) {
            float r1 = 0.f, r2 = 0.f; // per-vector partial sums in fp32 (8 terms), then f64
            uint32_t todo = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float x = s[o][k];
                if ((uint32_t)k < c && x >= h.valid_thresh) {
                    int e;
                    const float lg = f32_split_log2(x, &e);
                    const float rel = __fsub_rn(3.0102999566f * __fadd_rn((float)e, lg), h.min_db[o]);
                    uint32_t idx;
                    if (f32_guarded_index(h.guard[o], e, lg, 4095u, &idx)) atomicAdd(&s_hist[o * 4096 + idx], 1u);
                    else todo |= 1u << k;
                    r1 += rel;
                    r2 += rel * rel;
                }
            }
            while (todo) { // the few samples the guard did not decide: exact threshold comparison, one per iteration
                const int k = __ffs((int)todo) - 1;
                todo &= todo - 1;
                float x = s[o][0];
#pragma unroll
                for (int j = 1; j < 8; ++j) x = k == j ? s[o][j] : x;
                int e;
                const float lg = f32_split_log2(x, &e);
                const float rel = __fsub_rn(3.0102999566f * __fadd_rn((float)e, lg), h.min_db[o]);
                const uint32_t idx = edge_index(s_edges + o * 4096, 4095, x, (int)(rel * h.inv_span4096[o]));
                atomicAdd(&s_hist[o * 4096 + idx], 1u);
            }
            s1[o] += (double)r1;
            s2[o] += (double)r2;
        }
    }
#pragma unroll
    for (int o = 0; o < NOPS; ++o) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s1[o] += __shfl_xor_sync(0xffffffffu, s1[o], d);
            s2[o] += __shfl_xor_sync(0xffffffffu, s2[o], d);
        }
        if ((threadIdx.x & 31) == 0) { s_red[o][0][threadIdx.x >> 5] = s1[o]; s_red[o][1][threadIdx.x >> 5] = s2[o]; }
    }
    __syncthreads();
    if (threadIdx.x < NOPS) {
        const int o = threadIdx.x;
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { a += s_red[o][0][w]; b += s_red[o][1][w]; }
        atomicAdd(&h.sums[o][0], a);
        atomicAdd(&h.sums[o][1], b);
    }
#pragma unroll
    for (int o = 0; o < NOPS; ++o)
        for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x)
            if (s_hist[o * 4096 + i]) atomicAdd(&h.hist[o][i], (unsigned long long)s_hist[o * 4096 + i]);
}
cudaError_t launch_f32_hist4096(const void* a, const void* b, int a_u16, int b_u16, int op0, int op1, int nops, uint64_t n,
                                float valid_thresh, const float* min_db, const float* inv_span4096, const float* const* edges4096,
                                unsigned long long* const* hist4096, double* const* sums, const F32GuardHost* guards, int sm_count,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    F32HistArgs h{};
    h.valid_thresh = valid_thresh;
    for (int o = 0; o < nops; ++o) {
        h.min_db[o] = min_db[o]; h.inv_span4096[o] = inv_span4096[o]; h.edges[o] = edges4096[o]; h.hist[o] = hist4096[o]; h.sums[o] = sums[o];
        h.guard[o] = F32Guard{guards[o].e0, guards[o].f0, guards[o].scale, guards[o].guard};
    }
    const F32Src src = make_src(a, b, a_u16, b_u16, op0, nops == 2 ? op1 : -1);
    const size_t smem = (size_t)nops * 4096 * 8;
    if (nops == 2) {
        if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_f32_hist4096<2>), smem)) return e;
        k_f32_hist4096<2><<<sm_count * 3, 256, smem, stream>>>(src, n, h);
    } else {
        k_f32_hist4096<1><<<sm_count * 4, 256, smem, stream>>>(src, n, h);
    }
    return cudaGetLastError();
}

// ---- pass 3: quantisation by level thresholds -----------------------------------------------------------
struct F32QuantArgs {
    float valid_thresh;
    float low_db[2], high_db[2], inv_range[2], gamma[2]; // guess only
    const float* edges[2];                               // [n_levels + 1]: edges[k], k = 1..n_levels
    uint32_t n_levels;                                   // 255, 65535 or 255 (CLAHE bins)
    const uint8_t* remap[2];                             // 256-entry scale_u16_to_u8 table or nullptr
    int key_plane;                                       // 1: write u16 key = level + 1 for valid, 0 for invalid (CLAHE bridge)
    void* out[2];
    int out_vec;                                         // outputs are 16-byte aligned
    F32Guard guard[2];
};
template <int NOPS, typename OutT>
__global__ void __launch_bounds__(256) k_f32_quantize(F32Src src, uint64_t n, F32QuantArgs qa) {
    __shared__ float s_edges[NOPS][257];
    __shared__ uint8_t s_remap[NOPS][256];
    const bool small = qa.n_levels <= 256;
#pragma unroll
    for (int o = 0; o < NOPS; ++o) {
        if (small)
            for (uint32_t i = threadIdx.x; i <= qa.n_levels; i += blockDim.x) s_edges[o][i] = qa.edges[o][i];
        s_remap[o][threadIdx.x] = qa.remap[o] ? qa.remap[o][threadIdx.x] : (uint8_t)threadIdx.x;
    }
    __syncthreads();
    const float fl = (float)qa.n_levels;
    const uint64_t nvec = (n + 7) / 8, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        float s[NOPS][8];
        const uint32_t c = src.get8<NOPS>(v, n, s);
#pragma unroll
        for (int o = 0; o < NOPS; ++o) {
            const float* edges = small ? s_edges[o] : qa.edges[o];
            uint32_t q[8];
            // first the guarded direct levels of all 8 samples (straight-line code), then the threshold comparisons of the few
            // samples the guard did not decide, one per iteration: the slow path is issued once per vector, not once per sample
            uint32_t todo = 0, valid = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { // branch-free: every lane evaluates, invalid samples are masked afterwards
                const float x0 = s[o][k];
                const bool ok = (uint32_t)k < c && x0 >= qa.valid_thresh;
                int e;
                const float lg = f32_split_log2(x0, &e);
                const bool decided = f32_guarded_index(qa.guard[o], e, lg, qa.n_levels, &q[k]);
                valid |= (ok ? 1u : 0u) << k;
                todo |= ((ok && !decided) ? 1u : 0u) << k;
            }
            while (todo) {
                const int k = __ffs((int)todo) - 1;
                todo &= todo - 1;
                float x0 = s[o][0];
#pragma unroll
                for (int j = 1; j < 8; ++j) x0 = k == j ? s[o][j] : x0;
                int e;
                const float lg = f32_split_log2(x0, &e);
                float x = 3.0102999566f * __fadd_rn((float)e, lg);
                x = fminf(fmaxf(x, qa.low_db[o]), qa.high_db[o]);
                x = (x - qa.low_db[o]) * qa.inv_range[o];
                if (qa.gamma[o] != 1.0f) x = __powf(fmaxf(x, 0.0f), qa.gamma[o]);
                const uint32_t lvl = edge_index(edges, qa.n_levels, x0, (int)(x * fl));
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = k == j ? lvl : q[j];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool v = (valid >> k) & 1u;
                if (qa.key_plane) q[k] = v ? q[k] + 1u : 0u;
                else if (sizeof(OutT) == 1) q[k] = s_remap[o][v ? (q[k] & 255u) : 0u]; // invalid samples are 0 before scale_u16_to_u8 (autoscale.rs:444, 669-670)
                else q[k] = v ? q[k] : 0u;
            }
            OutT* out = reinterpret_cast<OutT*>(qa.out[o]) + v * 8;
            if (c == 8 && qa.out_vec) {
                if (sizeof(OutT) == 2) {
                    *reinterpret_cast<uint4*>(out) = make_uint4(q[0] | (q[1] << 16), q[2] | (q[3] << 16), q[4] | (q[5] << 16), q[6] | (q[7] << 16));
                } else {
                    *reinterpret_cast<uint2*>(out) = make_uint2(q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24), q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((uint32_t)k < c) out[k] = (OutT)q[k];
            }
        }
    }
}
cudaError_t launch_f32_quantize(const void* a, const void* b, int a_u16, int b_u16, int op0, int op1, int nops, uint64_t n,
                                float valid_thresh, const float* low_db, const float* high_db, const float* gamma,
                                const float* const* level_edges, uint32_t n_levels, const uint8_t* const* remap, int key_plane,
                                void* const* out, int out_u8, const F32GuardHost* guards, int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    F32QuantArgs qa{};
    qa.valid_thresh = valid_thresh;
    qa.n_levels = n_levels;
    qa.key_plane = key_plane;
    qa.out_vec = 1;
    for (int o = 0; o < nops; ++o) {
        const float range = fmaxf(high_db[o] - low_db[o], 1.0f);
        qa.low_db[o] = low_db[o]; qa.high_db[o] = high_db[o]; qa.inv_range[o] = 1.0f / range; qa.gamma[o] = gamma[o];
        qa.edges[o] = level_edges[o];
        qa.remap[o] = remap ? remap[o] : nullptr;
        qa.out[o] = out[o];
        qa.guard[o] = F32Guard{guards[o].e0, guards[o].f0, guards[o].scale, guards[o].guard};
        if (reinterpret_cast<uintptr_t>(out[o]) & 15u) qa.out_vec = 0;
    }
    const F32Src src = make_src(a, b, a_u16, b_u16, op0, nops == 2 ? op1 : -1);
    const int grid = sm_count * 8;
    if (nops == 2) {
        if (out_u8) k_f32_quantize<2, uint8_t><<<grid, 256, 0, stream>>>(src, n, qa);
        else k_f32_quantize<2, uint16_t><<<grid, 256, 0, stream>>>(src, n, qa);
    } else {
        if (out_u8) k_f32_quantize<1, uint8_t><<<grid, 256, 0, stream>>>(src, n, qa);
        else k_f32_quantize<1, uint16_t><<<grid, 256, 0, stream>>>(src, n, qa);
    }
    return cudaGetLastError();
}

} // namespace sarpro

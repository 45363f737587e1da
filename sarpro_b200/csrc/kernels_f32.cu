// kernels_f32.cu — rasters whose samples are not u16-valued: polarization ratios / normalised
// differences (ops.rs:10-44) and calibrated f32 inputs. The op is fused into every loader, so the
// combined f32 plane of the reference (a new Array2<f32>, ops.rs) is never written.
//
// Exactness without device transcendentals: every index the reference derives from a sample
//   4096-bin stat index  autoscale.rs:113-116      quantised level  autoscale.rs:440-442 / 649-651
//   CLAHE bin            autoscale.rs:263
// is a monotone non-decreasing function of the sample value, so the host (plan_f32.cpp) converts each
// index boundary into an f32 threshold with the same libm the reference uses, and the device only
// compares: index(v) = #{k : v >= edge[k]}. A fast __log2f-based guess lands within a step or two of the
// answer; the two correction loops make the result independent of the guess.
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

struct F32Src {
    const void* a;
    const void* b;
    int a_u16, b_u16, op;
    __device__ __forceinline__ float get(uint64_t i) const {
        const float x = a_u16 ? (float)reinterpret_cast<const uint16_t*>(a)[i] : reinterpret_cast<const float*>(a)[i];
        if (op < 0) return x;
        const float y = b_u16 ? (float)reinterpret_cast<const uint16_t*>(b)[i] : reinterpret_cast<const float*>(b)[i];
        return f32_pol(op, x, y);
    }
};

// ---- pass 1: min / max / count over valid samples (autoscale.rs:37-55) --------------------------------
__global__ void __launch_bounds__(256) k_f32_scan(F32Src src, uint64_t n, float valid_thresh, F32Scan* __restrict__ out) {
    uint32_t mn = 0xffffffffu, mx = 0;
    unsigned long long cnt = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = src.get(i);
        if (v >= valid_thresh) { // valid samples are positive: the bit pattern orders like the value
            const uint32_t k = __float_as_uint(v);
            mn = min(mn, k);
            mx = max(mx, k);
            cnt++;
        }
    }
    mn = warp_reduce_min(mn);
    mx = warp_reduce_max(mx);
    unsigned c32 = warp_reduce_add((unsigned)cnt); // < 2^32 per warp by construction (n < 2^32)
    if ((threadIdx.x & 31) == 0 && c32) {
        atomicMin(&out->min_key, mn);
        atomicMax(&out->max_key, mx);
        atomicAdd(&out->valid_count, (unsigned long long)c32);
    }
}
cudaError_t launch_f32_scan(const void* a, const void* b, int a_u16, int b_u16, int op, uint64_t n, float valid_thresh,
                            F32Scan* out, int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_f32_scan<<<sm_count * 8, 256, 0, stream>>>(F32Src{a, b, a_u16, b_u16, op}, n, valid_thresh, out);
    return cudaGetLastError();
}

// index(v) = #{k in [1, n_edges] : v >= edges[k]}; edges[0] unused, edges ascending (ties allowed)
__device__ __forceinline__ uint32_t edge_index(const float* __restrict__ edges, uint32_t n_edges, float v, int guess) {
    int g = guess < 0 ? 0 : (guess > (int)n_edges ? (int)n_edges : guess);
    while (g < (int)n_edges && v >= edges[g + 1]) ++g;
    while (g > 0 && v < edges[g]) --g;
    return (uint32_t)g;
}

// ---- pass 2: 4096-bin histogram over [min_db, max_db] + mean / M2 accumulators -----------------------
struct F32HistArgs {
    float valid_thresh;
    float min_db, inv_span4096; // guess: (db - min_db) * inv_span * 4096
    const float* edges;         // [4096]: edges[k], k = 1..4095
    unsigned long long* hist;   // [4096]
    double* sums;               // [0] = sum(db - min_db), [1] = sum((db - min_db)^2)   (fp32 logs, f64 accumulation)
};
__global__ void __launch_bounds__(256) k_f32_hist4096(F32Src src, uint64_t n, F32HistArgs h) {
    __shared__ float s_edges[4096];
    __shared__ uint32_t s_hist[4096];
    __shared__ double s_red[2][8];
    for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x) {
        s_edges[i] = h.edges[i];
        s_hist[i] = 0;
    }
    __syncthreads();
    double s1 = 0.0, s2 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = src.get(i);
        if (v >= h.valid_thresh) {
            const float rel = __fsub_rn(3.0102999566f * __log2f(v), h.min_db);
            const uint32_t idx = edge_index(s_edges, 4095, v, (int)(rel * h.inv_span4096));
            atomicAdd(&s_hist[idx], 1u);
            s1 += (double)rel;
            s2 += (double)rel * (double)rel;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = s1; s_red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
        atomicAdd(&h.sums[0], a);
        atomicAdd(&h.sums[1], b);
    }
    for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&h.hist[i], (unsigned long long)s_hist[i]);
}
cudaError_t launch_f32_hist4096(const void* a, const void* b, int a_u16, int b_u16, int op, uint64_t n, float valid_thresh,
                                float min_db, float inv_span4096, const float* edges4096, unsigned long long* hist4096,
                                double* sums, int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    F32HistArgs h{valid_thresh, min_db, inv_span4096, edges4096, hist4096, sums};
    k_f32_hist4096<<<sm_count * 4, 256, 0, stream>>>(F32Src{a, b, a_u16, b_u16, op}, n, h);
    return cudaGetLastError();
}

// ---- pass 3: quantisation by level thresholds -----------------------------------------------------------
struct F32QuantArgs {
    float valid_thresh;
    float low_db, high_db, inv_range, gamma; // guess only
    const float* edges;                      // [n_levels + 1]: edges[k], k = 1..n_levels
    uint32_t n_levels;                       // 255, 65535 or 255 (CLAHE bins)
    const uint8_t* remap;                    // 256-entry scale_u16_to_u8 table or nullptr
    int key_plane;                           // 1: write u16 key = level + 1 for valid, 0 for invalid (CLAHE bridge)
};
template <typename OutT>
__global__ void __launch_bounds__(256) k_f32_quantize(F32Src src, uint64_t n, F32QuantArgs qa, OutT* __restrict__ out) {
    __shared__ float s_edges[257];
    __shared__ uint8_t s_remap[256];
    const bool small = qa.n_levels <= 256;
    if (small)
        for (uint32_t i = threadIdx.x; i <= qa.n_levels; i += blockDim.x) s_edges[i] = qa.edges[i];
    s_remap[threadIdx.x] = qa.remap ? qa.remap[threadIdx.x] : (uint8_t)threadIdx.x;
    __syncthreads();
    const float* edges = small ? s_edges : qa.edges;
    const float fl = (float)qa.n_levels;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = src.get(i);
        uint32_t o = 0;
        if (v >= qa.valid_thresh) {
            float x = 3.0102999566f * __log2f(v);
            x = fminf(fmaxf(x, qa.low_db), qa.high_db);
            x = (x - qa.low_db) * qa.inv_range;
            if (qa.gamma != 1.0f) x = __powf(fmaxf(x, 0.0f), qa.gamma);
            const uint32_t lvl = edge_index(edges, qa.n_levels, v, (int)(x * fl));
            o = qa.key_plane ? lvl + 1 : (sizeof(OutT) == 1 ? (uint32_t)s_remap[lvl & 255u] : lvl);
        } else if (!qa.key_plane && sizeof(OutT) == 1) {
            o = s_remap[0]; // invalid samples are 0 before scale_u16_to_u8 (autoscale.rs:444, 669-670)
        }
        out[i] = (OutT)o;
    }
}
cudaError_t launch_f32_quantize(const void* a, const void* b, int a_u16, int b_u16, int op, uint64_t n, float valid_thresh,
                                float low_db, float high_db, float gamma, const float* level_edges, uint32_t n_levels,
                                const uint8_t* remap, int key_plane, uint8_t* out_u8, uint16_t* out_u16, int sm_count,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const float range = fmaxf(high_db - low_db, 1.0f);
    F32QuantArgs qa{valid_thresh, low_db, high_db, 1.0f / range, gamma, level_edges, n_levels, remap, key_plane};
    const F32Src src{a, b, a_u16, b_u16, op};
    if (out_u8) k_f32_quantize<uint8_t><<<sm_count * 8, 256, 0, stream>>>(src, n, qa, out_u8);
    else k_f32_quantize<uint16_t><<<sm_count * 8, 256, 0, stream>>>(src, n, qa, out_u16);
    return cudaGetLastError();
}

} // namespace sarpro

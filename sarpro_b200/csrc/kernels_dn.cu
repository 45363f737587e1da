// kernels_dn.cu — the u16-DN passes of the raster path for sm_100a.
//
// Design: for a u16 raster every per-pixel quantity of the reference's path
// (dB pipeline.rs:19-20, validity :22, 4096-bin stat histogram autoscale.rs:108-117, clip/gamma/
// quantise :437-446 / :647-655, scale_u16_to_u8 :348-364, CLAHE bin :263) is a pure function of the
// DN, so pass A only counts DNs (per CLAHE tile) and pass B is "load DN, look up, consume".
// No transcendental runs on the device; the host planner (plan.cpp) evaluates them once per
// distinct DN with the same libm the reference uses.
//
// All kernels are HBM-streaming: 128-bit coalesced loads, shared-memory privatised histograms /
// tables, grids sized to a multiple of the SM count.
#include "common.cuh"
#include "kernels.h"

namespace sarpro {

// =============================================================================================
// Pass A — per-tile DN histogram.  2 B/px read; one shared-memory atomic per non-zero pixel.
// =============================================================================================
// ---- general kernel (any column count) --------------------------------------------------------------------
// Per 8-pixel vector: one range test on the OR of the four words, then per pixel one shift/mask (ALU pipe),
// one IMAD (FMA pipe) and one shared-memory reduction — no per-pixel predicates. DN 0 is counted like any other
// value (a run of identical DNs is one POPC-merged update per replica). Vectors that hold a DN >= HOT or that
// straddle the unit's column range take the per-pixel path (two lanes per row for the ragged ends).
// 1024 threads, one CTA per SM, NCOPY = 1 << LOGC lane-interleaved replicas: word (dn << LOGC) | (lane % NCOPY).
__device__ __forceinline__ void red_shared_inc(uint32_t addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

template <int HOT, int LOGC>
__global__ void __launch_bounds__(1024, 1) k_dn_hist2(const uint16_t* __restrict__ dn, uint64_t cols,
                                                      const HistUnit* __restrict__ units, uint32_t n_units,
                                                      uint32_t* __restrict__ tile_hist) {
    extern __shared__ uint32_t sh[];
    constexpr uint32_t NCOPY = 1u << LOGC;
    constexpr uint32_t kHotMask = ~(uint32_t)(HOT - 1) & 0xffffu; // HOT is a power of two
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sh) + (lane & (NCOPY - 1)) * 4u;
    for (uint32_t i = tid; i < (uint32_t)HOT * NCOPY; i += blockDim.x) sh[i] = 0;
    __syncthreads();

    const uint16_t* dn_al = reinterpret_cast<const uint16_t*>(reinterpret_cast<uintptr_t>(dn) & ~uintptr_t(15));
    const uint64_t eoff = (uint64_t)(dn - dn_al);

    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const HistUnit un = units[u];
        uint32_t* __restrict__ gh = tile_hist + (size_t)un.tile * 65536u;
        const uint32_t seg = un.c1 - un.c0;

        auto add_px = [&](uint32_t d) {
            if (d < (uint32_t)HOT) red_shared_inc(sbase + (d << (LOGC + 2)));
            else atomicAdd(&gh[d], 1u);
        };
        auto add_vec = [&](const uint4& q) {
            if (((q.x | q.y | q.z | q.w) & (kHotMask * 0x10001u)) == 0) {
                red_shared_inc(sbase + (q.x & 0xffffu) * (4u * NCOPY));
                red_shared_inc(sbase + (q.x >> 16) * (4u * NCOPY));
                red_shared_inc(sbase + (q.y & 0xffffu) * (4u * NCOPY));
                red_shared_inc(sbase + (q.y >> 16) * (4u * NCOPY));
                red_shared_inc(sbase + (q.z & 0xffffu) * (4u * NCOPY));
                red_shared_inc(sbase + (q.z >> 16) * (4u * NCOPY));
                red_shared_inc(sbase + (q.w & 0xffffu) * (4u * NCOPY));
                red_shared_inc(sbase + (q.w >> 16) * (4u * NCOPY));
            } else {
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) { add_px(w[k] & 0xffffu); add_px(w[k] >> 16); }
            }
        };

        for (uint32_t r = un.r0 + warp; r < un.r1; r += nwarps) {
            const uint64_t e0 = (uint64_t)r * cols + un.c0 + eoff, e1 = e0 + seg;
            const uint64_t f0 = (e0 + 7) >> 3, f1 = e1 >> 3; // full vectors [f0, f1)
            if (f0 < f1) {
                for (uint64_t v = f0 + lane; v < f1; v += 128) {
                    uint4 q[4];
                    bool ok[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint64_t vv = v + 32u * j;
                        ok[j] = vv < f1;
                        if (ok[j]) q[j] = ld_stream_u4(dn_al + (vv << 3));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (ok[j]) add_vec(q[j]);
                }
                // ragged ends: elements [e0, f0*8) and [f1*8, e1)
                if (lane < 2) {
                    const uint64_t a = lane == 0 ? e0 : (f1 << 3), b = lane == 0 ? (f0 << 3) : e1;
                    for (uint64_t e = a; e < b; ++e) add_px(dn_al[e]);
                }
            } else {
                for (uint64_t e = e0 + lane; e < e1; e += 32) add_px(dn_al[e]);
            }
        }
        __syncthreads();
        for (uint32_t b = tid; b < (uint32_t)HOT; b += blockDim.x) {
            uint32_t s = 0;
            if (NCOPY >= 4) {
                uint4* p = reinterpret_cast<uint4*>(&sh[b << LOGC]);
#pragma unroll
                for (uint32_t c = 0; c < NCOPY / 4; ++c) {
                    // rotate the starting replica per bin so that the lanes of a warp spread over the banks
                    const uint32_t cc = (c + b) & (NCOPY / 4 - 1);
                    const uint4 t = p[cc];
                    s += t.x + t.y + t.z + t.w;
                    p[cc] = make_uint4(0, 0, 0, 0);
                }
            } else {
#pragma unroll
                for (uint32_t c = 0; c < NCOPY; ++c) { s += sh[(b << LOGC) | c]; sh[(b << LOGC) | c] = 0; }
            }
            if (s) atomicAdd(&gh[b], s);
        }
        __syncthreads();
    }
}

// ---- third generation: four rows in flight per warp, loads software-pipelined one step ahead, units fetched
// from a device counter (no tail imbalance). Needs cols % 8 == 0 (all rows of a unit share the vector phase).
template <int HOT, int LOGC>
__global__ void __launch_bounds__(1024, 1) k_dn_hist3(const uint16_t* __restrict__ dn, uint64_t cols,
                                                      const HistUnit* __restrict__ units, uint32_t n_units,
                                                      uint32_t* __restrict__ tile_hist, uint32_t* __restrict__ counter) {
    extern __shared__ uint32_t sh[];
    __shared__ uint32_t s_unit;
    constexpr uint32_t NCOPY = 1u << LOGC;
    constexpr uint32_t kHotMask = ~(uint32_t)(HOT - 1) & 0xffffu;
    constexpr int R = 4; // rows in flight per warp
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sh) + (lane & (NCOPY - 1)) * 4u;
    for (uint32_t i = tid; i < (uint32_t)HOT * NCOPY; i += 1024) sh[i] = 0;

    const uint16_t* dn_al = reinterpret_cast<const uint16_t*>(reinterpret_cast<uintptr_t>(dn) & ~uintptr_t(15));
    const uint32_t eoff = (uint32_t)(dn - dn_al);
    const uint4* const dn4 = reinterpret_cast<const uint4*>(dn_al);
    const uint32_t cols8 = (uint32_t)(cols >> 3);

    for (;;) {
        __syncthreads(); // table zeroed / previous flush done; s_unit free
        if (tid == 0) s_unit = atomicAdd(counter, 1u);
        __syncthreads();
        const uint32_t u = s_unit;
        if (u >= n_units) break;
        const HistUnit un = units[u];
        uint32_t* __restrict__ gh = tile_hist + (size_t)un.tile * 65536u;
        const uint32_t off = un.c0 + eoff, seg = un.c1 - un.c0;
        const uint32_t vf0 = (off + 7) >> 3, vf1 = (off + seg) >> 3; // full vectors [vf0, vf1) relative to the row start
        const uint32_t nfull = vf1 > vf0 ? vf1 - vf0 : 0;

        auto add_px = [&](uint32_t d) {
            if (d < (uint32_t)HOT) red_shared_inc(sbase + (d << (LOGC + 2)));
            else atomicAdd(&gh[d], 1u);
        };
        auto add_vec = [&](const uint4& q) {
            if (((q.x | q.y | q.z | q.w) & (kHotMask * 0x10001u)) == 0) {
                red_shared_inc(sbase + (q.x & 0xffffu) * (4u * NCOPY));
                red_shared_inc(sbase + (q.x >> 16) * (4u * NCOPY));
                red_shared_inc(sbase + (q.y & 0xffffu) * (4u * NCOPY));
                red_shared_inc(sbase + (q.y >> 16) * (4u * NCOPY));
                red_shared_inc(sbase + (q.z & 0xffffu) * (4u * NCOPY));
                red_shared_inc(sbase + (q.z >> 16) * (4u * NCOPY));
                red_shared_inc(sbase + (q.w & 0xffffu) * (4u * NCOPY));
                red_shared_inc(sbase + (q.w >> 16) * (4u * NCOPY));
            } else {
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) { add_px(w[k] & 0xffffu); add_px(w[k] >> 16); }
            }
        };

        for (uint32_t rbase = un.r0 + warp; rbase < un.r1; rbase += 32 * R) {
            // vector index of (row j, full vector 0); rows beyond the unit are skipped
            uint32_t vb[R];
            bool rok[R];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const uint32_t r = rbase + 32u * j;
                rok[j] = r < un.r1;
                vb[j] = (rok[j] ? r : un.r0) * cols8 + vf0; // < 2^32 vectors for any raster in HBM
            }
            if (nfull) {
                uint4 q[R], qn[R];
                auto load = [&](uint4* dst, uint32_t i) {
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if (rok[j] && i < nfull) dst[j] = ld_stream_u4(dn4 + (size_t)(vb[j] + i));
                };
                load(q, lane);
                for (uint32_t i = lane; i < nfull + 0u; i += 32) {
                    load(qn, i + 32);
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if (rok[j]) add_vec(q[j]);
#pragma unroll
                    for (int j = 0; j < R; ++j) q[j] = qn[j];
                }
            }
            // ragged ends: lanes 0..7 take (row j, side) = (lane >> 1, lane & 1)
            if (lane < 2 * R) {
                const int j = lane >> 1;
                const uint32_t r = rbase + 32u * j;
                if (r < un.r1) {
                    const uint64_t rs = (uint64_t)r * cols; // element index of the row start (aligned view)
                    uint64_t a, b;
                    if (nfull) {
                        a = (lane & 1) ? rs + ((uint64_t)vf1 << 3) : rs + off;
                        b = (lane & 1) ? rs + off + seg : rs + ((uint64_t)vf0 << 3);
                    } else {
                        a = rs + off;
                        b = (lane & 1) ? a : rs + off + seg;
                    }
                    for (uint64_t e = a; e < b; ++e) add_px(dn_al[e]);
                }
            }
        }
        __syncthreads();
        for (uint32_t b = tid; b < (uint32_t)HOT; b += 1024) {
            uint32_t s = 0;
            uint4* p = reinterpret_cast<uint4*>(&sh[b << LOGC]);
#pragma unroll
            for (uint32_t c = 0; c < NCOPY / 4; ++c) {
                const uint32_t cc = (c + b) & (NCOPY / 4 - 1);
                const uint4 t = p[cc];
                s += t.x + t.y + t.z + t.w;
                p[cc] = make_uint4(0, 0, 0, 0);
            }
            if (s) atomicAdd(&gh[b], s);
        }
    }
}

template <int HOT, int LOGC>
static cudaError_t launch_dn_hist3_t(const uint16_t* dn, uint64_t cols, const HistUnit* units, uint32_t n_units,
                                     uint32_t* tile_hist, uint32_t* counter, int sm_count, cudaStream_t stream) {
    const size_t smem = (size_t)HOT * (1u << LOGC) * sizeof(uint32_t);
    if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_dn_hist3<HOT, LOGC>), smem)) return e;
    uint32_t grid = (uint32_t)sm_count;
    if (grid > n_units) grid = n_units;
    if (grid == 0) return cudaSuccess;
    k_dn_hist3<HOT, LOGC><<<grid, 1024, smem, stream>>>(dn, cols, units, n_units, tile_hist, counter);
    return cudaGetLastError();
}

template <int HOT, int LOGC>
static cudaError_t launch_dn_hist2_t(const uint16_t* dn, uint64_t cols, const HistUnit* units, uint32_t n_units,
                                     uint32_t* tile_hist, int sm_count, cudaStream_t stream) {
    const size_t smem = (size_t)HOT * (1u << LOGC) * sizeof(uint32_t);
    if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_dn_hist2<HOT, LOGC>), smem)) return e;
    uint32_t grid = (uint32_t)sm_count;
    if (grid > n_units) grid = n_units;
    if (grid == 0) return cudaSuccess;
    k_dn_hist2<HOT, LOGC><<<grid, 1024, smem, stream>>>(dn, cols, units, n_units, tile_hist);
    return cudaGetLastError();
}

cudaError_t launch_dn_hist(const uint16_t* dn, uint64_t cols, const HistUnit* units, uint32_t n_units,
                           uint32_t* tile_hist, int sm_count, int variant, cudaStream_t stream, uint32_t* counter) {
    // variant: table shape 20 / 21 / 22 = DN < 4096 / 2048 / 1024 in the replicated shared histogram (8 / 16 / 32 replicas).
    // The production kernel (k_dn_hist3) needs the unit counter and cols % 8 == 0; other rasters take the general kernel.
    const bool gen3 = counter && cols % 8 == 0;
    switch (variant) {
    case 20: return gen3 ? launch_dn_hist3_t<4096, 3>(dn, cols, units, n_units, tile_hist, counter, sm_count, stream)
                         : launch_dn_hist2_t<4096, 3>(dn, cols, units, n_units, tile_hist, sm_count, stream);
    case 22: return gen3 ? launch_dn_hist3_t<1024, 5>(dn, cols, units, n_units, tile_hist, counter, sm_count, stream)
                         : launch_dn_hist2_t<1024, 5>(dn, cols, units, n_units, tile_hist, sm_count, stream);
    default: return gen3 ? launch_dn_hist3_t<2048, 4>(dn, cols, units, n_units, tile_hist, counter, sm_count, stream)
                         : launch_dn_hist2_t<2048, 4>(dn, cols, units, n_units, tile_hist, sm_count, stream);
    }
}

// total[dn] = sum over tiles; max_dn = highest non-empty DN; and the non-empty bins as (dn, count) pairs for the host
// planner (a GRD band uses ~1e3 of the 65,536 DNs: the planner reads 12 KB of pinned memory instead of scanning 256 KB
// of it cold). Each block compacts its 256 DNs and takes pairs[off .. off+cnt) with one atomic; blk[block] = {off, cnt}.
// The allocation order is arbitrary, the order inside a block and the block index give the DN order back. Pairs beyond
// `cap` are dropped (the host sees off + cnt > cap and reads the dense totals instead).
__global__ void __launch_bounds__(256) k_hist_total(const uint32_t* __restrict__ tile_hist, uint32_t n_tiles,
                                                    uint32_t* __restrict__ total, uint32_t* __restrict__ max_dn,
                                                    uint32_t* __restrict__ n_present, uint2* __restrict__ blk,
                                                    uint2* __restrict__ pairs, uint32_t cap) {
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_off;
    const uint32_t d = blockIdx.x * 256u + threadIdx.x; // < 65536
    uint32_t s = 0;
    for (uint32_t t = 0; t < n_tiles; ++t) s += tile_hist[(size_t)t * 65536u + d];
    total[d] = s;
    unsigned m = warp_reduce_max(s ? d : 0u);
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (lane == 0 && m) atomicMax(max_dn, m);
    if (!blk) return;
    const uint32_t bal = __ballot_sync(0xffffffffu, s != 0u);
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, cnt = 0;
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
        const uint32_t c = s_w[i];
        if (i < wid) before += c;
        cnt += c;
    }
    if (threadIdx.x == 0) {
        const uint32_t off = cnt ? atomicAdd(n_present, cnt) : 0u;
        s_off = off;
        blk[blockIdx.x] = make_uint2(off, cnt);
    }
    __syncthreads();
    if (s) {
        const uint32_t pos = s_off + before + __popc(bal & ((1u << lane) - 1u));
        if (pos < cap) pairs[pos] = make_uint2(d, s);
    }
}
cudaError_t launch_hist_total(const uint32_t* tile_hist, uint32_t n_tiles, uint32_t* total, uint32_t* max_dn,
                              uint32_t* n_present, uint2* present, uint32_t cap, cudaStream_t stream) {
    k_hist_total<<<65536 / 256, 256, 0, stream>>>(tile_hist, n_tiles, total, max_dn, n_present, present,
                                                  present ? present + 256 : nullptr, cap);
    return cudaGetLastError();
}

// =============================================================================================
// CLAHE tile statistics
// =============================================================================================
__global__ void k_clahe_tile256(ClaheStatJobs jobs) {
    __shared__ uint32_t h[256];
    const ClaheStatJob& jb = jobs.j[blockIdx.z]; // one band per grid layer
    const uint32_t* __restrict__ tile_hist = jb.tile_hist;
    const uint16_t* __restrict__ lut = jb.lut;
    uint32_t* __restrict__ tile256 = jb.tile256;
    const uint32_t max_dn = jb.plan->max_present_dn;
    const uint32_t t = blockIdx.x, part = blockIdx.y, nparts = gridDim.y;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t* th = tile_hist + (size_t)t * 65536u;
    for (uint32_t d = 1 + part * blockDim.x + threadIdx.x; d <= max_dn; d += nparts * blockDim.x) {
        const uint32_t c = th[d];
        if (c) atomicAdd(&h[lut[d] & 255u], c);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&tile256[t * 256u + threadIdx.x], h[threadIdx.x]);
}
cudaError_t launch_clahe_tile256(const uint32_t* tile_hist, const uint16_t* lut, uint32_t n_tiles, const PlanDev* plan,
                                 uint32_t* tile256, cudaStream_t stream) {
    // the brightest present DN is only known on the device: a fixed split of the DN range (a GRD band ends below DN 4096
    // except for point targets; CTAs whose share lies beyond max_dn return at once)
    ClaheStatJobs jobs{};
    jobs.j[0] = ClaheStatJob{tile_hist, lut, plan, tile256, nullptr, nullptr};
    return launch_clahe_tile256_2(jobs, 1, n_tiles, stream);
}
cudaError_t launch_clahe_tile256_2(const ClaheStatJobs& jobs, int n_bands, uint32_t n_tiles, cudaStream_t stream) {
    const uint32_t parts = 8;
    k_clahe_tile256<<<dim3(n_tiles, parts, (unsigned)n_bands), 256, 0, stream>>>(jobs);
    return cudaGetLastError();
}

// autoscale.rs:271-302. Every quantity is an integer or an integer multiple of 1/128 below 2^53, so the
// f64 sums are exact and independent of summation order; products/quotients use explicit round-to-nearest
// intrinsics (never contracted to FMA).
__global__ void __launch_bounds__(256) k_clahe_cdf(ClaheStatJobs jobs, const uint64_t* __restrict__ tile_px) {
    __shared__ double s_ex[256];
    __shared__ unsigned long long s_pre[256];
    const ClaheStatJob& jb = jobs.j[blockIdx.y]; // one band per grid row
    const uint32_t* __restrict__ tile256 = jb.tile256;
    double* __restrict__ cdf = jb.cdf;
    float* __restrict__ cdf32 = jb.cdf32;
    const uint32_t t = blockIdx.x, i = threadIdx.x;
    uint32_t h = tile256[t * 256u + i];
    const double avg = __ddiv_rn((double)tile_px[t], 256.0);       // :242-245
    const double thr = fmax(__dmul_rn(2.0, avg), 1.0);             // :273
    double ex = 0.0;
    if ((double)h > thr) {                                         // :276-279
        ex = __dsub_rn((double)h, thr);
        h = (uint32_t)thr;
    }
    s_ex[i] = ex;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)i < o) s_ex[i] = __dadd_rn(s_ex[i], s_ex[i + o]);
        __syncthreads();
    }
    const double excess = s_ex[0];
    const double add_per_bin = floor(__ddiv_rn(excess, 256.0));    // :282
    const double remf = round(__dsub_rn(excess, __dmul_rn(add_per_bin, 256.0))); // :283
    const unsigned long long remainder = remf > 0.0 ? (unsigned long long)remf : 0ull;
    h = (uint32_t)__dadd_rn((double)h, add_per_bin);               // :285
    h += (uint32_t)(remainder / 256ull) + (i < (uint32_t)(remainder % 256ull) ? 1u : 0u); // :287-292
    s_pre[i] = h;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) { // inclusive scan (integers)
        unsigned long long v = (int)i >= o ? s_pre[i - o] : 0ull;
        __syncthreads();
        s_pre[i] += v;
        __syncthreads();
    }
    const double total = fmax((double)s_pre[255], 1.0);            // :295
    double c = __ddiv_rn((double)s_pre[i], total);                 // :299-300
    c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);
    cdf[t * 256u + i] = c;
    if (cdf32) cdf32[t * 256u + i] = (float)c;
}
cudaError_t launch_clahe_cdf(const uint32_t* tile256, const uint64_t* tile_px, uint32_t n_tiles, double* cdf,
                             float* cdf32, cudaStream_t stream) {
    ClaheStatJobs jobs{};
    jobs.j[0] = ClaheStatJob{nullptr, nullptr, nullptr, const_cast<uint32_t*>(tile256), cdf, cdf32};
    return launch_clahe_cdf_2(jobs, 1, tile_px, n_tiles, stream);
}
cudaError_t launch_clahe_cdf_2(const ClaheStatJobs& jobs, int n_bands, const uint64_t* tile_px, uint32_t n_tiles, cudaStream_t stream) {
    k_clahe_cdf<<<dim3(n_tiles, (unsigned)n_bands), 256, 0, stream>>>(jobs, tile_px);
    return cudaGetLastError();
}

// autoscale.rs:308-318 for one axis: f = g/tile - 0.5; t = max(floor(f),0); d = f - t; neighbours clamped.
__global__ void k_clahe_axis(uint32_t n, uint32_t global_offset, uint32_t tile_size, uint32_t n_tiles,
                             double* __restrict__ d, double* __restrict__ omd, uint16_t* __restrict__ t01,
                             int32_t* __restrict__ m, uint16_t* __restrict__ sat, uint16_t* __restrict__ sat1) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = global_offset + i;
    const double f = __dsub_rn(__ddiv_rn((double)g, (double)tile_size), 0.5);
    const long long t = (long long)fmax(floor(f), 0.0);
    const double dd = __dsub_rn(f, (double)t);
    const double om = __dsub_rn(1.0, dd);
    const long long hi = (long long)n_tiles - 1;
    const long long t0 = t < 0 ? 0 : (t > hi ? hi : t);
    const long long t1 = (t + 1) < 0 ? 0 : ((t + 1) > hi ? hi : (t + 1));
    const double one = __dadd_rn(om, dd); // value of c*(1-d) + c*d for c == 1.0
    d[i] = dd;
    omd[i] = om;
    const double one_m = 0x1.fffffffffffffp-1; // 1 - 2^-53, the other value fl(om + d) takes
    t01[i] = (uint16_t)(t0 | (t1 << 8) | (one == 1.0 ? 0x80 : 0) | (one == one_m ? 0x40 : 0));
    if (m) m[i] = (int32_t)(2ll * (long long)g - (long long)tile_size * (2 * t + 1));
    if (sat) {
        const double c = one < 0.0 ? 0.0 : (one > 1.0 ? 1.0 : one);
        sat[i] = (uint16_t)__dmul_rn(c, 255.0);
    }
    if (sat1) { // the same sample when the other axis gives fl(om' + d') == 1 - 2^-53 (autoscale.rs:327-329 with CDFs 1.0)
        const double v = __dadd_rn(__dmul_rn(one_m, om), __dmul_rn(one_m, dd));
        const double c = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
        sat1[i] = (uint16_t)__dmul_rn(c, 255.0);
    }
}
cudaError_t launch_clahe_axis(uint32_t n, uint32_t global_offset, uint32_t tile_size, uint32_t n_tiles, double* d,
                              double* omd, uint16_t* t01, int32_t* m, uint16_t* sat, cudaStream_t stream, uint16_t* sat1) {
    if (n == 0) return cudaSuccess;
    k_clahe_axis<<<(n + 255) / 256, 256, 0, stream>>>(n, global_offset, tile_size, n_tiles, d, omd, t01, m, sat, sat1);
    return cudaGetLastError();
}

// =============================================================================================
// Pass B — full-resolution apply (LUT only). 2 B/px read, 1 or 2 B/px written.
// =============================================================================================
constexpr uint32_t kLutHot = 8192; // DN range staged in shared memory; brighter DNs read the global LUT

template <typename OutT>
__global__ void __launch_bounds__(512) k_apply_lut(const uint16_t* __restrict__ dn, uint64_t n,
                                                   const uint16_t* __restrict__ lut, OutT* __restrict__ out) {
    __shared__ OutT s_lut[kLutHot];
    for (uint32_t i = threadIdx.x; i < kLutHot; i += blockDim.x) s_lut[i] = (OutT)lut[i];
    __syncthreads();
    auto look = [&](uint32_t d) -> uint32_t { return d < kLutHot ? (uint32_t)s_lut[d] : (uint32_t)__ldg(&lut[d]); };
    const bool aligned = ((reinterpret_cast<uintptr_t>(dn) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(out) & (8 * sizeof(OutT) - 1)) == 0);
    const uint64_t nvec = aligned ? (n >> 3) : 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 q = ld_stream_u4(dn + (v << 3));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[2 * k] = look(w[k] & 0xffffu);
            o[2 * k + 1] = look(w[k] >> 16);
        }
        if (sizeof(OutT) == 1) {
            uint2 p;
            p.x = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
            p.y = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
            st_stream_u2(out + (v << 3), p);
        } else {
            uint4 p;
            p.x = o[0] | (o[1] << 16);
            p.y = o[2] | (o[3] << 16);
            p.z = o[4] | (o[5] << 16);
            p.w = o[6] | (o[7] << 16);
            st_stream_u4(out + (v << 3), p);
        }
    }
    for (uint64_t e = (nvec << 3) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
        out[e] = (OutT)look(dn[e]);
}
cudaError_t launch_apply_lut(const uint16_t* dn, uint64_t n, const uint16_t* lut, uint8_t* out_u8, uint16_t* out_u16,
                             int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    uint64_t want = (n / 8 + 511) / 512;
    uint32_t grid = (uint32_t)(want < (uint64_t)sm_count * 4 ? (want ? want : 1) : (uint64_t)sm_count * 4);
    if (out_u8) k_apply_lut<uint8_t><<<grid, 512, 0, stream>>>(dn, n, lut, out_u8);
    else k_apply_lut<uint16_t><<<grid, 512, 0, stream>>>(dn, n, lut, out_u16);
    return cudaGetLastError();
}

// =============================================================================================
// CLAHE blend (exact): autoscale.rs:320-329 then :602.
// =============================================================================================
// Explicit _rn intrinsics keep the reference's operation order and forbid FMA contraction.
__device__ __forceinline__ double clahe_blend_exact(double c00, double c01, double c10, double c11, double dx,
                                                    double omdx, double dy, double omdy) {
    const double top = __dadd_rn(__dmul_rn(c00, omdx), __dmul_rn(c01, dx));
    const double bottom = __dadd_rn(__dmul_rn(c10, omdx), __dmul_rn(c11, dx));
    return __dadd_rn(__dmul_rn(top, omdy), __dmul_rn(bottom, dy));
}
__device__ __forceinline__ uint32_t clahe_quantize(double v, double max_val) {
    const double n = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); // clamp(0,1); NaN cannot occur
    return (uint32_t)__dmul_rn(n, max_val);               // `as u16` truncation
}

template <typename OutT>
__global__ void __launch_bounds__(256) k_apply_clahe(const uint16_t* __restrict__ dn, uint32_t rows, uint32_t cols,
                                                     const uint16_t* __restrict__ lut, ClaheDev cl, double max_val,
                                                     OutT* __restrict__ out, uint32_t* __restrict__ minmax) {
    uint32_t mn = 0xffffffffu, mx = 0;
    const uint32_t vec_per_row = (cols + 7) / 8;
    const uint64_t total = (uint64_t)rows * vec_per_row;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(idx / vec_per_row);
        const uint32_t c0 = (uint32_t)(idx % vec_per_row) * 8;
        const double dy = cl.row_dy[r], omdy = cl.row_omdy[r];
        const uint32_t ty = cl.row_t[r];
        const double* cdf_t0 = cl.cdf + (size_t)(ty & 7u) * cl.tiles_x * 256u;
        const double* cdf_t1 = cl.cdf + (size_t)((ty >> 8) & 7u) * cl.tiles_x * 256u;
        const uint64_t base = (uint64_t)r * cols + c0;
#pragma unroll 1
        for (uint32_t k = 0; k < 8 && c0 + k < cols; ++k) {
            const uint32_t c = c0 + k;
            const uint32_t d = dn[base + k];
            uint32_t o = 0;
            if (d != 0) {
                const uint32_t bin = lut[d] & 255u;
                const uint32_t tx = cl.col_t[c];
                const uint32_t x0 = (tx & 7u) * 256u + bin, x1 = ((tx >> 8) & 7u) * 256u + bin;
                const double v = clahe_blend_exact(cdf_t0[x0], cdf_t0[x1], cdf_t1[x0], cdf_t1[x1], cl.col_dx[c],
                                                   cl.col_omdx[c], dy, omdy);
                o = clahe_quantize(v, max_val);
            }
            out[base + k] = (OutT)o;
            mn = min(mn, o);
            mx = max(mx, o);
        }
    }
    mn = warp_reduce_min(mn);
    mx = warp_reduce_max(mx);
    if ((threadIdx.x & 31) == 0 && mn != 0xffffffffu) {
        atomicMin(&minmax[0], mn);
        atomicMax(&minmax[1], mx);
    }
}
cudaError_t launch_apply_clahe(const uint16_t* dn, uint32_t rows, uint32_t cols, const uint16_t* lut, ClaheDev cl,
                               int max_val, uint8_t* out_u8, uint16_t* out_u16, uint32_t* minmax, int sm_count,
                               cudaStream_t stream) {
    if (rows == 0 || cols == 0) return cudaSuccess;
    const uint32_t grid = (uint32_t)sm_count * 8;
    if (out_u8) k_apply_clahe<uint8_t><<<grid, 256, 0, stream>>>(dn, rows, cols, lut, cl, (double)max_val, out_u8, minmax);
    else k_apply_clahe<uint16_t><<<grid, 256, 0, stream>>>(dn, rows, cols, lut, cl, (double)max_val, out_u16, minmax);
    return cudaGetLastError();
}

// =============================================================================================
// scale_u16_to_u8 pieces (autoscale.rs:348-364)
// =============================================================================================
__global__ void __launch_bounds__(512) k_remap_u8(uint8_t* __restrict__ data, uint64_t n,
                                                  const uint8_t* __restrict__ remap, const uint32_t* __restrict__ skip) {
    __shared__ uint8_t s[256];
    if (skip && *skip) return;
    if (threadIdx.x < 256) s[threadIdx.x] = remap[threadIdx.x];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
    const uint64_t nvec = aligned ? n >> 4 : 0;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        uint4 q = *reinterpret_cast<const uint4*>(data + (v << 4));
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = (uint32_t)s[w[k] & 255u] | ((uint32_t)s[(w[k] >> 8) & 255u] << 8) |
                   ((uint32_t)s[(w[k] >> 16) & 255u] << 16) | ((uint32_t)s[w[k] >> 24] << 24);
        *reinterpret_cast<uint4*>(data + (v << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (uint64_t e = (nvec << 4) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride)
        data[e] = s[data[e]];
}
cudaError_t launch_remap_u8(uint8_t* data, uint64_t n, const uint8_t* remap256, int sm_count, cudaStream_t stream,
                            const uint32_t* skip) {
    if (n == 0) return cudaSuccess;
    k_remap_u8<<<sm_count * 4, 512, 0, stream>>>(data, n, remap256, skip);
    return cudaGetLastError();
}

// autoscale.rs:352-363 for the 256 possible CLAHE samples; same f32 operations as plan.cpp make_u16_to_u8_remap
__global__ void __launch_bounds__(256) k_clahe_remap_decide(const uint32_t* __restrict__ minmax, uint8_t* __restrict__ remap,
                                                            uint32_t* __restrict__ skip, const PlanDev* __restrict__ plan) {
    uint32_t mn = minmax[0], mx = minmax[1];
    if (mn == 0xffffffffu) { mn = 0; mx = 0; }
    const bool identity = (mn == 0 && mx == 255) || (mn == 0 && mx == 0);
    const float fmn = (float)mn, fmx = (float)mx;
    const float scale = fmx > fmn ? __fdiv_rn(255.0f, __fsub_rn(fmx, fmn)) : 1.0f;
    float val = roundf(__fmul_rn(__fsub_rn((float)threadIdx.x, fmn), scale));
    val = val < 0.0f ? 0.0f : (val > 255.0f ? 255.0f : val);
    remap[threadIdx.x] = (uint8_t)val;
    if (threadIdx.x == 0) {
        skip[0] = identity ? 1u : 0u;
        // skip[1]: the vertical pass behind the re-run may be skipped only when neither the re-stretch nor the generic
        // horizontal kernel (plan->use_generic: the first vertical pass ran on rows the tensor-core kernel never wrote) ran
        skip[1] = (identity && !(plan && plan->use_generic)) ? 1u : 0u;
    }
}
cudaError_t launch_clahe_remap_decide(const uint32_t* minmax, uint8_t* remap256, uint32_t* skip, cudaStream_t stream,
                                      const PlanDev* plan) {
    k_clahe_remap_decide<<<1, 256, 0, stream>>>(minmax, remap256, skip, plan);
    return cudaGetLastError();
}

// {max, ~min} of both bands in one 4-word vector, so that a single all-reduce(max) merges the ranks' CLAHE sample
// extrema; unpack writes the merged values back to the bands' {min, max} words.
__global__ void k_minmax_pack(uint32_t* __restrict__ s0, uint32_t* __restrict__ s1, uint32_t* __restrict__ packed, int unpack) {
    uint32_t* s = threadIdx.x == 0 ? s0 : s1;
    if (!unpack) {
        packed[2 * threadIdx.x] = s[1];
        packed[2 * threadIdx.x + 1] = ~s[0];
    } else {
        s[1] = packed[2 * threadIdx.x];
        s[0] = ~packed[2 * threadIdx.x + 1];
    }
}
cudaError_t launch_minmax_pack(uint32_t* scalars0, uint32_t* scalars1, uint32_t* packed4, int unpack, cudaStream_t stream) {
    k_minmax_pack<<<1, 2, 0, stream>>>(scalars0, scalars1, packed4, unpack);
    return cudaGetLastError();
}

// ---- re-pitched rasters: replicate the edge sample into the padding columns -------------------------------------------------
__global__ void __launch_bounds__(256) k_pad_cols(uint16_t* __restrict__ dn, uint32_t rows, uint32_t width, uint32_t pitch) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    uint16_t* row = dn + (size_t)r * pitch;
    const uint16_t v = row[width - 1];
    for (uint32_t c = width; c < pitch; ++c) row[c] = v;
}
// dst[r][c] = src[r][min(c, width - 1)], dst rows `pitch` samples apart (a multiple of 8, 16-byte aligned), src rows `width`
// apart at whatever alignment: a thread writes one 128-bit vector from eight 2-byte loads (consecutive lanes read consecutive
// 16-byte pieces, L1 merges the 2-byte accesses of a line)
__global__ void __launch_bounds__(256) k_repitch(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, uint32_t rows,
                                                 uint32_t width, uint32_t pitch) {
    const uint32_t vpr = pitch / 8u; // vectors per row
    const uint64_t nvec = (uint64_t)rows * vpr, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint32_t r = (uint32_t)(v / vpr), c = (uint32_t)(v % vpr) * 8u;
        const uint16_t* s = src + (size_t)r * width;
        uint32_t x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = s[min(c + (uint32_t)k, width - 1u)];
        *reinterpret_cast<uint4*>(dst + (size_t)r * pitch + c) =
            make_uint4(x[0] | (x[1] << 16), x[2] | (x[3] << 16), x[4] | (x[5] << 16), x[6] | (x[7] << 16));
    }
}
cudaError_t launch_repitch(const uint16_t* src, uint16_t* dst, uint32_t rows, uint32_t width, uint32_t pitch, int sm_count,
                           cudaStream_t stream) {
    if (rows == 0 || width == 0) return cudaSuccess;
    k_repitch<<<sm_count * 8, 256, 0, stream>>>(src, dst, rows, width, pitch);
    return cudaGetLastError();
}
cudaError_t launch_pad_cols(uint16_t* dn, uint32_t rows, uint32_t width, uint32_t pitch, cudaStream_t stream) {
    if (rows == 0 || pitch <= width || width == 0) return cudaSuccess;
    k_pad_cols<<<(rows + 255) / 256, 256, 0, stream>>>(dn, rows, width, pitch);
    return cudaGetLastError();
}

// ---- sharded scene: rows + extrema out of the all-gathered slots (see GatherGeom in kernels.h) ----------------------------
__global__ void k_gather_tail(const uint32_t* __restrict__ s0, const uint32_t* __restrict__ s1, uint32_t* __restrict__ tail) {
    const uint32_t* s = threadIdx.x < 2 ? s0 : s1;
    tail[threadIdx.x] = s[threadIdx.x & 1u];
}
cudaError_t launch_gather_tail(const uint32_t* scalars0, const uint32_t* scalars1, uint32_t* tail4, cudaStream_t stream) {
    k_gather_tail<<<1, 4, 0, stream>>>(scalars0, scalars1, tail4);
    return cudaGetLastError();
}
__global__ void __launch_bounds__(128) k_gather_unpack(const unsigned char* __restrict__ gathered, GatherGeom gg,
                                                       unsigned char* __restrict__ canvas0, unsigned char* __restrict__ canvas1,
                                                       uint32_t* __restrict__ scalars0, uint32_t* __restrict__ scalars1,
                                                       uint32_t* __restrict__ flag) {
    const uint32_t oy = blockIdx.x, b = blockIdx.y;
    if (oy == 0 && b == 0 && threadIdx.x < 2 && gg.clahe) { // merged extrema of band threadIdx.x
        uint32_t mn = 0xffffffffu, mx = 0;
        for (uint32_t r = 0; r < gg.world; ++r) {
            const uint32_t* t = reinterpret_cast<const uint32_t*>(gathered + (size_t)(r + 1) * gg.slot_bytes - 16) + 2 * threadIdx.x;
            if (t[0] != 0xffffffffu) { mn = min(mn, t[0]); mx = max(mx, t[1]); }
        }
        uint32_t* s = threadIdx.x ? scalars1 : scalars0;
        s[0] = mn;
        s[1] = mx;
        if (mn == 0xffffffffu) { mn = 0; mx = 0; }
        const bool identity = (mn == 0 && mx == 255) || (mn == 0 && mx == 0); // as k_clahe_remap_decide
        if (!identity) atomicOr(flag, 1u);
    }
    uint32_t r = 0;
    while (r + 1 < gg.world && !(oy >= gg.oy0[r] && oy < gg.oy1[r])) ++r;
    if (!(oy >= gg.oy0[r] && oy < gg.oy1[r])) return; // a row no rank owns (cannot happen: the bands partition the rows)
    const unsigned char* src = gathered + (size_t)r * gg.slot_bytes + ((size_t)b * gg.max_rows + (oy - gg.oy0[r])) * gg.out_pitch;
    unsigned char* dst = (b ? canvas1 : canvas0) + (size_t)(gg.pad_top + oy) * gg.out_pitch;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | gg.out_pitch) & 15u) == 0) {
        for (uint32_t i = threadIdx.x; i < gg.out_pitch / 16u; i += blockDim.x)
            reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    } else {
        for (uint32_t i = threadIdx.x; i < gg.out_pitch; i += blockDim.x) dst[i] = src[i];
    }
}
cudaError_t launch_gather_unpack(const unsigned char* gathered, GatherGeom gg, unsigned char* canvas0, unsigned char* canvas1,
                                 uint32_t* scalars0, uint32_t* scalars1, uint32_t* flag, cudaStream_t stream) {
    if (gg.out_rows == 0) return cudaSuccess;
    k_gather_unpack<<<dim3(gg.out_rows, 2), 128, 0, stream>>>(gathered, gg, canvas0, canvas1, scalars0, scalars1, flag);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(512) k_minmax_u16(const uint16_t* __restrict__ data, uint64_t n,
                                                    uint32_t* __restrict__ minmax) {
    uint32_t mn = 0xffffffffu, mx = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(data) & 15) == 0;
    const uint64_t nvec = aligned ? n >> 3 : 0;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 q = ld_stream_u4(data + (v << 3));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mn = min(mn, min(w[k] & 0xffffu, w[k] >> 16));
            mx = max(mx, max(w[k] & 0xffffu, w[k] >> 16));
        }
    }
    for (uint64_t e = (nvec << 3) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        mn = min(mn, (uint32_t)data[e]);
        mx = max(mx, (uint32_t)data[e]);
    }
    mn = warp_reduce_min(mn);
    mx = warp_reduce_max(mx);
    if ((threadIdx.x & 31) == 0 && mn != 0xffffffffu) {
        atomicMin(&minmax[0], mn);
        atomicMax(&minmax[1], mx);
    }
}
cudaError_t launch_minmax_u16(const uint16_t* data, uint64_t n, uint32_t* minmax, int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_minmax_u16<<<sm_count * 4, 512, 0, stream>>>(data, n, minmax);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(512) k_scale_u16_to_u8(const uint16_t* __restrict__ data, uint64_t n,
                                                         const uint32_t* __restrict__ minmax,
                                                         uint8_t* __restrict__ out) {
    const float mn = (float)minmax[0], mx = (float)minmax[1];
    const float scale = mx > mn ? __fdiv_rn(255.0f, __fsub_rn(mx, mn)) : 1.0f;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        float val = roundf(__fmul_rn(__fsub_rn((float)data[e], mn), scale));
        val = val < 0.0f ? 0.0f : (val > 255.0f ? 255.0f : val);
        out[e] = (uint8_t)val;
    }
}
cudaError_t launch_scale_u16_to_u8(const uint16_t* data, uint64_t n, const uint32_t* minmax, uint8_t* out, int sm_count,
                                   cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_scale_u16_to_u8<<<sm_count * 4, 512, 0, stream>>>(data, n, minmax, out);
    return cudaGetLastError();
}

} // namespace sarpro

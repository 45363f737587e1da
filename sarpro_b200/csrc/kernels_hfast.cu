// kernels_hfast.cu — the production horizontal Lanczos pass for u8 samples, fused with the per-pixel
// stage (pass B). One CTA (256 threads) = a strip of output columns x a block of source rows.
//
// Per group of 4 source rows:
//   1. thread t owns the 8-sample vector column t of the strip's source span (span <= 256 vectors); the
//      raw samples of the NEXT group are prefetched into registers (128-bit streaming loads) before the
//      current group is accumulated, so HBM latency overlaps the arithmetic of the same CTA;
//   2. samples are produced (LUT, or CLAHE blend) and stored row-interleaved in shared memory: word w of
//      rows 0..3 is one 16-byte slot, so the accumulate loop fetches 4 rows with one LDS.128;
//   3. thread t < strip width accumulates output column t for the 4 rows with dp2a (two i16 taps x two
//      u8 samples per instruction); the taps live in registers for the whole kernel (template MAXP).
//
// CLAHE (autoscale.rs:307-330, :602) fast path: the bilinear CDF blend is evaluated in fp32 from a shared
// float4 table (the four tile CDFs of a bilinear cell, per bin), scaled by 255*2^16 and floored to fixed
// point. The sample k = u >> 16 is accepted only when the fraction u & 0xffff is at least kGuardQ/65536
// away from both neighbouring integers — the fp32 error analysis in DESIGN.md bounds
// |255*(v32 - v_ref)| < 4.5e-4 < kGuardQ/65536 — otherwise the pixel is queued and recomputed with the
// exact f64 operation order of the reference once the group is staged. Bins whose four CDFs are exactly
// 1.0 (everything above the p99 clip) take a closed form: v_ref = fl(omdy+dy) whenever fl(omdx+dx) == 1.
// The output is bit-identical to the exact kernel (kernels_resize.cu); tests compare both to the oracle.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace sarpro {

constexpr int kFRows = 4;
constexpr uint32_t kDeferCap = 1020;
constexpr int kGuardQ = 48;                    // 48/65536 = 7.3e-4 sample units
constexpr float kScaleQ = 255.0f * 65536.0f;
// shared memory map (bytes); everything the per-pixel code touches sits at a compile-time offset
constexpr uint32_t kOffLut = 0;                // u8 [8192]   DN -> sample / CLAHE bin
constexpr uint32_t kLutHot = 8192;
constexpr uint32_t kOffQuad = 8192;            // float4 [8][256] (CLAHE only)
constexpr uint32_t kOffDefer = kOffQuad + 8 * 256 * 16; // u32 [kDeferCap] + counter + pad
constexpr uint32_t kOffRemap = kOffDefer + (kDeferCap + 4) * 4;
constexpr uint32_t kOffRowsClahe = kOffRemap + 256;
constexpr uint32_t kOffRowsLut = kOffLut + kLutHot;
constexpr uint32_t kOffRowsImage = 0;

__device__ __forceinline__ double hf_blend_exact(double c00, double c01, double c10, double c11, double dx,
                                                 double omdx, double dy, double omdy) {
    const double top = __dadd_rn(__dmul_rn(c00, omdx), __dmul_rn(c01, dx));
    const double bottom = __dadd_rn(__dmul_rn(c10, omdx), __dmul_rn(c11, dx));
    return __dadd_rn(__dmul_rn(top, omdy), __dmul_rn(bottom, dy));
}

// exact sample of pixel (r, c) with DN d (reference operation order, f64). Arguments are passed by value so
// that the kernel-parameter struct never has to be materialised in local memory.
__device__ __noinline__ uint32_t hf_clahe_exact_impl(const uint16_t* __restrict__ lut, const double* __restrict__ cdf,
                                                     const double* __restrict__ col_dx, const double* __restrict__ col_omdx,
                                                     const uint16_t* __restrict__ col_t, const double* __restrict__ row_dy,
                                                     const double* __restrict__ row_omdy, const uint16_t* __restrict__ row_t,
                                                     uint32_t r, uint32_t c, uint32_t d) {
    if (d == 0) return 0;
    const uint32_t bin = lut[d] & 255u;
    const uint32_t ty = row_t[r], tx = col_t[c];
    const double* t0 = cdf + (size_t)(ty & 7u) * 8u * 256u;
    const double* t1 = cdf + (size_t)((ty >> 8) & 7u) * 8u * 256u;
    const uint32_t x0 = (tx & 7u) * 256u + bin, x1 = ((tx >> 8) & 7u) * 256u + bin;
    double v = hf_blend_exact(t0[x0], t0[x1], t1[x0], t1[x1], col_dx[c], col_omdx[c], row_dy[r], row_omdy[r]);
    v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    return (uint32_t)__dmul_rn(v, 255.0);
}
#define hf_clahe_exact(LUT, CL, R, C, D) \
    hf_clahe_exact_impl((LUT), (CL).cdf, (CL).col_dx, (CL).col_omdx, (CL).col_t, (CL).row_dy, (CL).row_omdy, (CL).row_t, (R), (C), (D))

template <int SRC, int MAXP>
__global__ void __launch_bounds__(256, (SRC == HSRC_DN_CLAHE ? 2 : 3)) k_hfast(HResizeArgs a, const HStrip* __restrict__ strips,
                                               const uint2* __restrict__ rowblocks, uint32_t strip_w) {
    extern __shared__ uint4 smem4[];
    if (a.skip && *a.skip) return;
    unsigned char* const smem = reinterpret_cast<unsigned char*>(smem4);
    constexpr uint32_t kOffRows = SRC == HSRC_DN_CLAHE ? kOffRowsClahe : (SRC == HSRC_DN_LUT ? kOffRowsLut : kOffRowsImage);
    uint4* const s_rows = reinterpret_cast<uint4*>(smem + kOffRows);
    const uint32_t tid = threadIdx.x;
    const HStrip st = strips[blockIdx.x];
    const uint2 rb = rowblocks[blockIdx.y];
    const uint32_t ox = blockIdx.x * strip_w + tid;
    const bool have_ox = tid < strip_w && ox < a.ax.out_size;

    // ---- one-time setup ---------------------------------------------------------------------
    int taps[MAXP];
#pragma unroll
    for (int i = 0; i < MAXP; ++i)
        taps[i] = (have_ox && (uint32_t)i < a.ax.pairs) ? (int)a.ax.packed[(size_t)ox * a.ax.pairs + i] : 0;
    const uint32_t woff = have_ox ? ((((a.ax.start[ox]) & ~3u) - st.sc0) >> 2) : 0;
    const int prec = a.ax.precision;

    if (SRC != HSRC_IMAGE)
        for (uint32_t i = tid; i < kLutHot; i += 256) smem[kOffLut + i] = (uint8_t)a.lut[i];
    for (uint32_t i = tid; i < a.rbw_words; i += 256) s_rows[i] = make_uint4(0, 0, 0, 0);

    uint32_t p_lo = 0;
    if (SRC == HSRC_DN_CLAHE) {
        const ClaheDev& cl = a.clahe;
        const uint32_t ty = cl.row_t[rb.x]; // the row block lies inside one vertical bilinear cell
        const uint32_t ty0 = ty & 7u, ty1 = (ty >> 8) & 7u;
        p_lo = cl.col_t[st.sc0 < a.src_cols ? st.sc0 : a.src_cols - 1] & 7u;
        float4* s_quad = reinterpret_cast<float4*>(smem + kOffQuad);
        for (uint32_t i = tid; i < 8 * 256; i += 256) {
            const uint32_t pc = p_lo + i / 256, bin = i & 255u;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pc < 8) {
                const uint32_t p1 = pc + 1 < 8 ? pc + 1 : 7;
                const double c00 = cl.cdf[((size_t)ty0 * 8 + pc) * 256 + bin], c01 = cl.cdf[((size_t)ty0 * 8 + p1) * 256 + bin];
                const double c10 = cl.cdf[((size_t)ty1 * 8 + pc) * 256 + bin], c11 = cl.cdf[((size_t)ty1 * 8 + p1) * 256 + bin];
                if (c00 == 1.0 && c01 == 1.0 && c10 == 1.0 && c11 == 1.0) q = make_float4(8.f, 8.f, 8.f, 8.f); // saturated bin
                else q = make_float4((float)c00, (float)c01, (float)c10, (float)c11);
            }
            s_quad[i] = q;
        }
        if (a.remap) smem[kOffRemap + tid] = a.remap[tid];
        if (tid == 0) reinterpret_cast<uint32_t*>(smem + kOffDefer)[kDeferCap] = 0;
    }
    __syncthreads();

    constexpr uint32_t esz = (SRC == HSRC_IMAGE) ? 1 : 2;
    const bool aligned = (reinterpret_cast<uintptr_t>(a.src) % 16 == 0) && (a.src_cols % 8 == 0);
    const bool have_vec = tid < st.nvec;
    const uint32_t c0 = st.sc0 + tid * 8;
    const bool full_vec = have_vec && aligned && c0 + 8 <= a.src_cols;
    const size_t row_pitch = (size_t)a.src_cols * esz;
    const unsigned char* const col_base = reinterpret_cast<const unsigned char*>(a.src) + (size_t)c0 * esz;
    uint32_t mn = 0xffffffffu, mx = 0;

    // CLAHE: per-column bilinear geometry of my 8 columns, fixed for the whole kernel
    float cdx[8];
    uint32_t cpack = 0; // 4 bits per column: tile-pair slot (3 bits) | fl(omdx+dx)==1 flag
    if (SRC == HSRC_DN_CLAHE) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t cc = c0 + k < a.src_cols ? c0 + k : a.src_cols - 1;
            const uint32_t ct = a.clahe.col_t[cc];
            cdx[k] = __fmul_rn((float)a.clahe.col_m[cc], a.clahe.inv2tw);
            cpack |= ((((ct & 7u) - p_lo) & 7u) | ((ct & 0x80u) ? 8u : 0u)) << (4 * k);
        }
    }

    uint4 q[kFRows]; // raw samples of the next group: 8 samples of my vector column per row
    auto prefetch = [&](uint32_t g) {
        const unsigned char* p = col_base + (size_t)g * row_pitch;
#pragma unroll
        for (int rr = 0; rr < kFRows; ++rr, p += row_pitch) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (g + rr < rb.y) {
                if (full_vec) {
                    if (SRC == HSRC_IMAGE) { const uint2 t = ld_stream_u2(p); v.x = t.x; v.y = t.y; }
                    else v = ld_stream_u4(p);
                } else if (have_vec) {
                    uint32_t e[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        e[k] = 0;
                        if (c0 + k < a.src_cols) e[k] = (SRC == HSRC_IMAGE) ? (uint32_t)p[k] : (uint32_t) reinterpret_cast<const uint16_t*>(p)[k];
                    }
                    if (SRC == HSRC_IMAGE) {
                        v.x = e[0] | (e[1] << 8) | (e[2] << 16) | (e[3] << 24);
                        v.y = e[4] | (e[5] << 8) | (e[6] << 16) | (e[7] << 24);
                    } else {
                        v.x = e[0] | (e[1] << 16); v.y = e[2] | (e[3] << 16); v.z = e[4] | (e[5] << 16); v.w = e[6] | (e[7] << 16);
                    }
                }
            }
            q[rr] = v;
        }
    };

    prefetch(rb.x);
    for (uint32_t g = rb.x; g < rb.y; g += kFRows) {
        // ---- produce the samples of this group and store them row-interleaved ------------------
        if (have_vec) {
            uint32_t w0[kFRows], w1[kFRows];
            if (SRC == HSRC_IMAGE) {
#pragma unroll
                for (int rr = 0; rr < kFRows; ++rr) { w0[rr] = q[rr].x; w1[rr] = q[rr].y; }
            } else if (SRC == HSRC_DN_LUT) {
                uint32_t any = 0;
#pragma unroll
                for (int rr = 0; rr < kFRows; ++rr) any |= q[rr].x | q[rr].y | q[rr].z | q[rr].w;
                if ((any & 0xE000E000u) == 0) { // all 32 DNs inside the shared LUT
#pragma unroll
                    for (int rr = 0; rr < kFRows; ++rr) {
                        const uint32_t w[4] = {q[rr].x, q[rr].y, q[rr].z, q[rr].w};
                        uint32_t o[8];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            o[2 * k] = smem[kOffLut + (w[k] & 0xffffu)];
                            o[2 * k + 1] = smem[kOffLut + (w[k] >> 16)];
                        }
                        w0[rr] = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
                        w1[rr] = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
                    }
                } else { // a DN beyond the shared LUT (rare): re-read the samples, look up in the global LUT
#pragma unroll
                    for (int rr = 0; rr < kFRows; ++rr) {
                        uint32_t lo = 0, hi = 0;
                        if (g + rr < rb.y) {
                            const uint16_t* src = reinterpret_cast<const uint16_t*>(a.src) + (size_t)(g + rr) * a.src_cols + c0;
#pragma unroll 1
                            for (int k = 0; k < 8 && c0 + k < a.src_cols; ++k) {
                                const uint32_t o = __ldg(&a.lut[src[k]]) & 255u;
                                if (k < 4) lo |= o << (8 * k); else hi |= o << (8 * (k - 4));
                            }
                        }
                        w0[rr] = lo;
                        w1[rr] = hi;
                    }
                }
            } else {
                const ClaheDev& cl = a.clahe;
                uint32_t any = 0;
#pragma unroll
                for (int rr = 0; rr < kFRows; ++rr) any |= q[rr].x | q[rr].y | q[rr].z | q[rr].w;
                const bool fast = ((any & 0xE000E000u) == 0) && (c0 + 8 <= a.src_cols) && (g + kFRows <= rb.y);
#pragma unroll
                for (int rr = 0; rr < kFRows; ++rr) { w0[rr] = 0; w1[rr] = 0; }
                if (fast) {
                    float sdy[kFRows], somdy[kFRows];
                    uint32_t satv[kFRows];
                    uint32_t needmask = 0; // bit rr*8+k: sample must be recomputed exactly
#pragma unroll
                    for (int rr = 0; rr < kFRows; ++rr) {
                        sdy[rr] = __fmul_rn((float)cl.row_dy[g + rr], kScaleQ);
                        somdy[rr] = __fmul_rn((float)cl.row_omdy[g + rr], kScaleQ);
                        satv[rr] = cl.row_sat[g + rr];
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float dx = cdx[k];
                        const float omdx = __fsub_rn(1.0f, dx);
                        const uint32_t cbase = kOffQuad + ((cpack >> (4 * k)) & 7u) * 4096u;
                        const bool col_one = ((cpack >> (4 * k)) & 8u) != 0;
#pragma unroll
                        for (int rr = 0; rr < kFRows; ++rr) {
                            const uint32_t wq = (k >> 1) == 0 ? q[rr].x : ((k >> 1) == 1 ? q[rr].y : ((k >> 1) == 2 ? q[rr].z : q[rr].w));
                            const uint32_t d = (k & 1) ? (wq >> 16) : (wq & 0xffffu);
                            const uint32_t bin = smem[kOffLut + d];
                            const float4 cq = *reinterpret_cast<const float4*>(smem + cbase + bin * 16u);
                            const float top = __fmaf_rn(cq.y, dx, __fmul_rn(cq.x, omdx));
                            const float bot = __fmaf_rn(cq.w, dx, __fmul_rn(cq.z, omdx));
                            const int u = __float2int_rd(__fmaf_rn(bot, sdy[rr], __fmul_rn(top, somdy[rr])));
                            const uint32_t kq = (uint32_t)(u >> 16);
                            const bool accept = ((uint32_t)((u & 0xffff) - kGuardQ) <= (uint32_t)(65535 - 2 * kGuardQ)) && (kq < 255u);
                            const bool sat = cq.x > 4.0f;
                            const bool valid = d != 0;
                            uint32_t o = sat ? satv[rr] : kq;
                            const bool need = valid && (sat ? !col_one : !accept);
                            o = (valid && !need) ? o : 0u;      // queued samples: placeholder 0, patched after the barrier
                            needmask |= need ? (1u << (rr * 8 + k)) : 0u;
                            mx = max(mx, o);
                            mn = min(mn, need ? 255u : o);      // a queued sample is neutral for the running min
                            if (k < 4) w0[rr] |= o << (8 * k); else w1[rr] |= o << (8 * (k - 4));
                        }
                    }
                    if (needmask) { // rare: queue for the exact path
                        uint32_t* s_defer = reinterpret_cast<uint32_t*>(smem + kOffDefer);
                        while (needmask) {
                            const uint32_t b = __ffs(needmask) - 1;
                            needmask &= needmask - 1;
                            const uint32_t slot = atomicAdd(&s_defer[kDeferCap], 1u);
                            const uint32_t rr = b >> 3, k = b & 7u;
                            if (slot < kDeferCap) {
                                s_defer[slot] = (rr << 28) | (tid * 8 + k);
                            } else { // queue full: resolve in place
                                const uint32_t d = reinterpret_cast<const uint16_t*>(a.src)[(size_t)(g + rr) * a.src_cols + c0 + k];
                                uint32_t o = hf_clahe_exact(a.lut, cl, g + rr, c0 + k, d);
                                mn = min(mn, o);
                                mx = max(mx, o);
                                const uint32_t sh = 8 * (k & 3u);
                                uint32_t word = k < 4 ? w0[0] : w1[0]; // select row rr without dynamic indexing
                                if (rr == 1) word = k < 4 ? w0[1] : w1[1];
                                if (rr == 2) word = k < 4 ? w0[2] : w1[2];
                                if (rr == 3) word = k < 4 ? w0[3] : w1[3];
                                word = (word & ~(0xffu << sh)) | (o << sh);
                                if (k < 4) { if (rr == 0) w0[0] = word; if (rr == 1) w0[1] = word; if (rr == 2) w0[2] = word; if (rr == 3) w0[3] = word; }
                                else { if (rr == 0) w1[0] = word; if (rr == 1) w1[1] = word; if (rr == 2) w1[2] = word; if (rr == 3) w1[3] = word; }
                            }
                        }
                    }
                } else {
                    // edge vectors / bright DNs / short last group: exact path for every sample (re-read from global)
#pragma unroll
                    for (int rr = 0; rr < kFRows; ++rr) {
                        uint32_t lo = 0, hi = 0;
                        if (g + rr < rb.y) {
                            const uint16_t* src = reinterpret_cast<const uint16_t*>(a.src) + (size_t)(g + rr) * a.src_cols + c0;
#pragma unroll 1
                            for (int k = 0; k < 8 && c0 + k < a.src_cols; ++k) {
                                const uint32_t o = hf_clahe_exact(a.lut, cl, g + rr, c0 + k, src[k]);
                                mn = min(mn, o);
                                mx = max(mx, o);
                                if (k < 4) lo |= o << (8 * k); else hi |= o << (8 * (k - 4));
                            }
                        }
                        w0[rr] = lo;
                        w1[rr] = hi;
                    }
                }
                if (a.remap) { // second run with a non-identity scale_u16_to_u8 (rare)
#pragma unroll
                    for (int rr = 0; rr < kFRows; ++rr) {
                        uint32_t x = w0[rr], y = w1[rr];
                        w0[rr] = (uint32_t)smem[kOffRemap + (x & 255u)] | ((uint32_t)smem[kOffRemap + ((x >> 8) & 255u)] << 8) |
                                 ((uint32_t)smem[kOffRemap + ((x >> 16) & 255u)] << 16) | ((uint32_t)smem[kOffRemap + (x >> 24)] << 24);
                        w1[rr] = (uint32_t)smem[kOffRemap + (y & 255u)] | ((uint32_t)smem[kOffRemap + ((y >> 8) & 255u)] << 8) |
                                 ((uint32_t)smem[kOffRemap + ((y >> 16) & 255u)] << 16) | ((uint32_t)smem[kOffRemap + (y >> 24)] << 24);
                    }
                }
            }
            s_rows[tid * 2] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
            s_rows[tid * 2 + 1] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        }
        __syncthreads();
        if (SRC == HSRC_DN_CLAHE) {
            // ---- exact recomputation of the queued samples ------------------------------------------
            uint32_t* s_defer = reinterpret_cast<uint32_t*>(smem + kOffDefer);
            const uint32_t nd = min(s_defer[kDeferCap], kDeferCap);
            if (nd) {
                for (uint32_t i = tid; i < nd; i += 256) {
                    const uint32_t e = s_defer[i];
                    const uint32_t rr = e >> 28, lc = e & 0x0fffffffu;
                    const uint32_t r = g + rr, c = st.sc0 + lc;
                    const uint32_t d = reinterpret_cast<const uint16_t*>(a.src)[(size_t)r * a.src_cols + c];
                    uint32_t o = hf_clahe_exact(a.lut, a.clahe, r, c, d);
                    mn = min(mn, o);
                    mx = max(mx, o);
                    if (a.remap) o = smem[kOffRemap + o];
                    smem[kOffRows + (size_t)(lc >> 2) * 16 + rr * 4 + (lc & 3u)] = (uint8_t)o;
                }
                __syncthreads();
                if (tid == 0) s_defer[kDeferCap] = 0;
            }
        }
        // ---- prefetch the next group, then accumulate this one ------------------------------------
        if (g + kFRows < rb.y) prefetch(g + kFRows);
        if (have_ox) {
            int acc[kFRows];
#pragma unroll
            for (int rr = 0; rr < kFRows; ++rr) acc[rr] = prec > 0 ? (1 << (prec - 1)) : 0;
#pragma unroll
            for (int m = 0; m < MAXP / 2; ++m) {
                if ((uint32_t)(2 * m) < a.ax.pairs) {
                    const uint4 w = s_rows[woff + m];
                    acc[0] = dp2a_hi_su(taps[2 * m + 1], w.x, dp2a_lo_su(taps[2 * m], w.x, acc[0]));
                    acc[1] = dp2a_hi_su(taps[2 * m + 1], w.y, dp2a_lo_su(taps[2 * m], w.y, acc[1]));
                    acc[2] = dp2a_hi_su(taps[2 * m + 1], w.z, dp2a_lo_su(taps[2 * m], w.z, acc[2]));
                    acc[3] = dp2a_hi_su(taps[2 * m + 1], w.w, dp2a_lo_su(taps[2 * m], w.w, acc[3]));
                }
            }
#pragma unroll
            for (int rr = 0; rr < kFRows; ++rr)
                if (g + rr < rb.y) {
                    int v = acc[rr] >> prec;
                    v = v < 0 ? 0 : (v > 255 ? 255 : v);
                    reinterpret_cast<uint8_t*>(a.temp)[(size_t)(g + rr - a.row0) * a.ax.out_size + ox] = (uint8_t)v;
                }
        }
        __syncthreads();
    }
    if (SRC == HSRC_DN_CLAHE && a.minmax) {
        mn = warp_reduce_min(mn);
        mx = warp_reduce_max(mx);
        if ((tid & 31) == 0 && mn != 0xffffffffu) {
            atomicMin(&a.minmax[0], mn);
            atomicMax(&a.minmax[1], mx);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------
bool hfast_supported(uint32_t pairs) { return pairs >= 2 && pairs <= 64; }

// Strips for the production kernel: the widest output strip whose source span fits 256 vector columns.
cudaError_t hfast_build_strips(const uint32_t* start_h, const uint32_t* size_h, uint32_t out_size, uint32_t in_size,
                               uint32_t window, uint32_t* strip_w_out, std::vector<HStrip>* strips, uint32_t* rbw_words) {
    if (out_size == 0) return cudaErrorInvalidConfiguration;
    auto span_ok = [&](uint32_t w) {
        for (uint32_t ox0 = 0; ox0 < out_size; ox0 += w) {
            const uint32_t ox1 = std::min(out_size, ox0 + w);
            uint32_t lo = ox0 == 0 ? 0 : start_h[ox0], hi = 0;
            for (uint32_t x = ox0; x < ox1; ++x) {
                lo = std::min(lo, start_h[x]);
                hi = std::max(hi, start_h[x] + size_h[x]);
            }
            if (ox1 == out_size) hi = in_size;
            if ((hi - (lo & ~7u) + 7) / 8 > 256) return false;
        }
        return true;
    };
    uint32_t w = std::min(256u, out_size);
    while (w > 1 && !span_ok(w)) w = w > 16 ? w - 8 : w - 1;
    if (!span_ok(w)) return cudaErrorInvalidConfiguration;
    strips->clear();
    uint32_t max_nvec = 0;
    for (uint32_t ox0 = 0; ox0 < out_size; ox0 += w) {
        const uint32_t ox1 = std::min(out_size, ox0 + w);
        uint32_t lo = ox0 == 0 ? 0 : start_h[ox0], hi = 0; // every source column is staged by some strip
        for (uint32_t x = ox0; x < ox1; ++x) {
            lo = std::min(lo, start_h[x]);
            hi = std::max(hi, start_h[x] + size_h[x]);
        }
        if (ox1 == out_size) hi = in_size;                   // (the CLAHE min/max must see the whole raster)
        HStrip st;
        st.sc0 = lo & ~7u;
        st.nvec = (hi - st.sc0 + 7) / 8;
        max_nvec = std::max(max_nvec, st.nvec);
        strips->push_back(st);
    }
    *strip_w_out = w;
    *rbw_words = max_nvec * 2 + (window + 16) / 4 + 2; // staged words + zero-tap overrun of the packed table
    return cudaSuccess;
}

template <int SRC, int MAXP>
static cudaError_t launch_hfast_t(const HResizeArgs& a, const HStrip* strips_dev, uint32_t n_strips, const uint2* rowblocks,
                                  uint32_t n_rowblocks, uint32_t strip_w, size_t smem, cudaStream_t stream) {
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_hfast<SRC, MAXP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    k_hfast<SRC, MAXP><<<dim3(n_strips, n_rowblocks), 256, smem, stream>>>(a, strips_dev, rowblocks, strip_w);
    return cudaGetLastError();
}

template <int SRC>
static cudaError_t launch_hfast_s(const HResizeArgs& a, const HStrip* strips_dev, uint32_t n_strips, const uint2* rowblocks,
                                  uint32_t n_rowblocks, uint32_t strip_w, size_t smem, cudaStream_t stream) {
    const uint32_t p = a.ax.pairs;
#define SARPRO_HF(P) return launch_hfast_t<SRC, P>(a, strips_dev, n_strips, rowblocks, n_rowblocks, strip_w, smem, stream)
    if (p <= 8) SARPRO_HF(8);
    if (p <= 16) SARPRO_HF(16);
    if (p <= 24) SARPRO_HF(24);
    if (p <= 32) SARPRO_HF(32);
    if (p <= 40) SARPRO_HF(40);
    if (p <= 48) SARPRO_HF(48);
    SARPRO_HF(64);
#undef SARPRO_HF
}

cudaError_t launch_hfast(const HResizeArgs& a, int src_kind, const HStrip* strips_dev, uint32_t n_strips,
                         const uint2* rowblocks_dev, uint32_t n_rowblocks, uint32_t strip_w, cudaStream_t stream) {
    if (a.n_rows == 0 || a.ax.out_size == 0 || n_rowblocks == 0) return cudaSuccess;
    const uint32_t off = src_kind == HSRC_DN_CLAHE ? kOffRowsClahe : (src_kind == HSRC_DN_LUT ? kOffRowsLut : kOffRowsImage);
    const size_t smem = off + (size_t)a.rbw_words * 16;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    if (src_kind == HSRC_IMAGE) return launch_hfast_s<HSRC_IMAGE>(a, strips_dev, n_strips, rowblocks_dev, n_rowblocks, strip_w, smem, stream);
    if (src_kind == HSRC_DN_LUT) return launch_hfast_s<HSRC_DN_LUT>(a, strips_dev, n_strips, rowblocks_dev, n_rowblocks, strip_w, smem, stream);
    return launch_hfast_s<HSRC_DN_CLAHE>(a, strips_dev, n_strips, rowblocks_dev, n_rowblocks, strip_w, smem, stream);
}

} // namespace sarpro

// kernels_hpipe.cu — pass B for u8 samples, second generation: the per-pixel stage (DN -> sample) fused
// with the horizontal Lanczos pass, organised around the two resources that bound it on an SM: shared-memory
// wavefronts (table gathers) and ALU-pipe issue slots.
//
// Persistent CTAs (one per SM) of NSUB independent 256-thread sub-blocks that share the look-up tables but work on
// alternating 4-row groups of the same piece (strip of output columns x run of source rows inside one vertical
// CLAHE cell); each sub-block synchronises with its own named barrier once per group. The host cuts the
// (strip, row) space into equal-weight contiguous pieces, one run per CTA, so there is no wave quantisation and
// the pipeline drains only at piece ends. Within a sub-block the work is software-pipelined three deep over
// triple-buffered staging:   produce(group i) | fix(group i-1) | accumulate(group i-2).
//
// produce — thread t owns the 8-sample vector column t of the strip's source span for 4 rows (DN prefetched
//   one group ahead with 128-bit streaming loads). Per 2 pixels: one VIMNMX.U16x2 clamps both DNs to the
//   staged table range and one IMAD turns them into two 16-bit shared-memory addresses of an R-way
//   lane-interleaved table (word (idx*R + lane%R), R = 16 when the range fits 64 KB, else 8: the lanes that
//   share a replica are the only ones that can conflict).
//     LUT strategies: the table word is the sample.
//     CLAHE (autoscale.rs:307-330, :602): the table word is the shared address of the pixel's bin entry in
//       an 8-way replicated float4 table (conflict-free LDS.128) holding the bilinear form of the four tile
//       CDFs of the cell: u = A + B*dx + (C + D*dx)*dy, in sample units with +1536 folded into A, so that the
//       integer part of 255*v sits at mantissa bits 13..20. Three FFMAs give u; u-6ulp must have the same
//       integer part (the fp32 error bound), otherwise the 8-pixel row segment is queued and
//       recomputed with the reference's exact f64 operation order by the fix stage. One FADD.RD moves
//       floor(u) into the low 16 bits (biased by 512); DPX 16x2 instructions track min/max and clamp.
// fix — queued row segments (and vectors that straddle a CLAHE cell boundary or the raster edge) are
//   recomputed exactly, 8 lanes per segment (by the warps that own no output column), from the f64 CDFs of
//   the piece staged in shared memory, and patched into the staged bytes.
// accumulate — thread t < strip width owns output column t for the 4 rows: taps in registers, samples staged
//   row-interleaved (one LDS.128 = 4 rows x 4 samples), dp2a.
//
// The result is bit-identical to the exact kernels (kernels_resize.cu); the GPU tests compare both to the oracle.
#include <algorithm>
#include <vector>

#include "clahe_exact.cuh"
#include "common.cuh"
#include "kernels.h"

namespace sarpro {

namespace hp {
constexpr int kRows = 4;                 // rows per group
constexpr uint32_t kHalf = 256;          // threads per half
constexpr uint32_t kQuadEntries = 257;   // 256 bins + the invalid-pixel entry
constexpr uint32_t kQuadCellBytes = kQuadEntries * 8 * 16; // one cell, 8 replicas
constexpr uint32_t kQueueCap = 254;      // entries per queue (+ counter + pad = 1 KB)
constexpr float kM0 = 1536.0f;           // binade [1024, 2048): ulp 2^-13
constexpr float kUlp = 1.0f / 8192.0f;
constexpr float kShift = 3.0f * kUlp;    // >= fp32 error bound of u (3.4e-4 sample units)
constexpr float kGuard = 6.0f * kUlp;
constexpr float kBigC = 12582912.0f - 1024.0f; // u + kBigC = 2^23*1.5 + 512 + (u - 1536)
} // namespace hp

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <uint32_t OFF>
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
// Pins a loop invariant in a register: ptxas otherwise rematerialises cheap-looking address constants inside the
// per-pixel loops (15 instructions per 8 pixels for the table base alone).
__device__ __forceinline__ uint32_t keep(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ void bar_half(uint32_t id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }

struct HPipeSmem {
    uint32_t lut, quad, cdf, dy, queue, stage, total; // byte offsets / total size
};
__host__ __device__ inline HPipeSmem hpipe_layout(bool clahe, uint32_t nsub, uint32_t lut_bytes, uint32_t max_rows,
                                                  uint32_t rbw_words) {
    HPipeSmem L;
    L.lut = 0;
    uint32_t o = lut_bytes;
    L.quad = o;
    if (clahe) o += 2 * hp::kQuadCellBytes;
    L.cdf = o;
    if (clahe) o += 6 * 256 * 8;
    L.dy = o;
    if (clahe) o += ((max_rows + 3) & ~3u) * 4u;
    L.queue = o;
    o += nsub * 3 * 1024;
    o = (o + 15) & ~15u;
    L.stage = o;
    o += nsub * 3 * rbw_words * 16u;
    L.total = o;
    return L;
}

// A piece (HPiece, kernels.h): rows [r0, r1) of strip `strip`. CTA b runs pieces [first[b], first[b+1]).

struct HPipeParams {
    const HStrip* strips;
    const HPiece* pieces;
    const uint32_t* cta_first;
    uint32_t strip_w, hot, lut_shift, max_rows;
};

template <bool CLAHE, int MAXP, int NSUB>
__global__ void __launch_bounds__(NSUB * 256, 1) k_hpipe(HResizeArgs a, HPipeParams pp) {
    extern __shared__ uint4 smem4[];
    unsigned char* const smem = reinterpret_cast<unsigned char*>(smem4);
    constexpr uint32_t NT = NSUB * 256;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t hot = pp.hot, lut_shift = pp.lut_shift; // table word of DN idx, replica r: byte (idx << lut_shift) + 4r
    const HPipeSmem L = hpipe_layout(CLAHE, NSUB, hot << lut_shift, pp.max_rows, a.rbw_words);
    const uint32_t tid = threadIdx.x, sub = tid >> 8, stid = tid & 255u, lane = tid & 31u;
    if (sbase + (hot << lut_shift) > 65536u) __trap(); // 16-bit table addresses (dynamic smem starts low on sm_100)

    // ---- once per CTA: DN -> table word, R lane-interleaved replicas ------------------------------------
    {
        uint4* s_lut4 = reinterpret_cast<uint4*>(smem + L.lut);
        const uint32_t per = 1u << (lut_shift - 4); // uint4 per table entry
        for (uint32_t i = tid; i < hot * per; i += NT) {
            const uint32_t idx = i >> (lut_shift - 4), r0 = (i & (per - 1)) * 4u;
            const uint32_t e = idx + 1 == hot ? a.hot_top : (a.lut[idx] & 255u);
            uint4 v;
            if (CLAHE) {
                const uint32_t bin = idx ? e : 256u; // DN 0 is the only invalid DN (pipeline.rs:22)
                const uint32_t b = sbase + L.quad + (bin * 8u + (r0 & 7u)) * 16u;
                v = make_uint4(b, b + 16u, b + 32u, b + 48u);
            } else {
                v = make_uint4(e, e, e, e);
            }
            s_lut4[i] = v;
        }
    }
    const bool aligned = (reinterpret_cast<uintptr_t>(a.src) % 16 == 0) && (a.src_cols % 8 == 0);
    const size_t row_pitch = (size_t)a.src_cols * 2;
    const uint32_t cap2 = keep((hot - 1u) * 0x10001u);
    const uint32_t lut_mul = keep(1u << lut_shift);
    const uint32_t cj = keep((sbase + L.lut + (lane & ((lut_mul >> 2) - 1u)) * 4u) * 0x10001u);
    const int prec = a.ax.precision;
    uint4* const s_stage = reinterpret_cast<uint4*>(smem + L.stage) + (size_t)sub * 3 * a.rbw_words;
    uint32_t* const s_queue = reinterpret_cast<uint32_t*>(smem + L.queue) + sub * 3 * 256;
    // scale_u16_to_u8 (autoscale.rs:348-364) takes min/max over ALL samples, invalid pixels (written as 0) included
    uint32_t mn2 = 0xffffffffu, mx2 = 0u;   // u16x2 running min / max of floor(u) - 1536 + 512 (not yet clamped)
    uint32_t mn_e = 0xffffffffu, mx_e = 0;  // exact-path samples

    int taps[MAXP];
    uint32_t taps_strip = 0xffffffffu;

    for (uint32_t pi = pp.cta_first[blockIdx.x]; pi < pp.cta_first[blockIdx.x + 1]; ++pi) {
        const HPiece pc = pp.pieces[pi];
        const HStrip st = pp.strips[pc.strip];
        const uint2 rb = make_uint2(pc.r0, pc.r1);
        const uint32_t ox = pc.strip * pp.strip_w + stid;
        const bool have_ox = stid < pp.strip_w && ox < a.ax.out_size;
        __syncthreads(); // the previous piece is done with the tables and the staging buffers

        // ---- per-piece setup ------------------------------------------------------------------------
        if (taps_strip != pc.strip) {
#pragma unroll
            for (int i = 0; i < MAXP; ++i)
                taps[i] = (have_ox && (uint32_t)i < a.ax.pairs) ? (int)a.ax.packed[(size_t)ox * a.ax.pairs + i] : 0;
            taps_strip = pc.strip;
        }
        const uint32_t woff = have_ox ? ((((a.ax.start[ox]) & ~3u) - st.sc0) >> 2) : 0;
        for (uint32_t i = stid; i < 3 * a.rbw_words; i += hp::kHalf) s_stage[i] = make_uint4(0, 0, 0, 0);
        if (stid < 3) s_queue[stid * 256 + 255] = 0;

        uint32_t cellA = 0, ty0 = 0, ty1 = 0;
        if (CLAHE) {
            const ClaheDev& cl = a.clahe;
            const uint32_t ty = cl.row_t[rb.x]; // the piece lies inside one vertical bilinear cell
            ty0 = ty & 7u;
            ty1 = (ty >> 8) & 7u;
            const uint32_t last_col = min(st.sc0 + st.nvec * 8u, a.src_cols) - 1u;
            cellA = cl.col_t[st.sc0] & 7u;
            const uint32_t cellB = cl.col_t[last_col] & 7u; // == cellA or cellA + 1 (strip span <= tile width)
            // saturated-bin shortcut (all four CDFs exactly 1.0 -> sample 255) is valid for a cell only when
            // fl(omdx+dx) == 1 for all of its columns in the strip and fl(omdy+dy) == 1 for all rows of the piece
            int okA = 1, okB = 1, okR = 1;
            for (uint32_t c = st.sc0 + tid; c <= last_col; c += NT) {
                const uint32_t ct = cl.col_t[c];
                if (!(ct & 0x80u)) { if ((ct & 7u) == cellA) okA = 0; else okB = 0; }
            }
            for (uint32_t r = rb.x + tid; r < rb.y; r += NT) if (cl.row_sat[r] != 255u) okR = 0;
            okA = __syncthreads_and(okA);
            okB = __syncthreads_and(okB);
            okR = __syncthreads_and(okR);
            const bool sat_ok[2] = {okA && okR, okB && okR};
            // f64 CDFs of the (up to) 3 x 2 tiles of the piece, for the exact path
            double* s_cdf = reinterpret_cast<double*>(smem + L.cdf);
            for (uint32_t i = tid; i < 6 * 256; i += NT) {
                const uint32_t t = i >> 8, bin = i & 255u;
                const uint32_t tyy = t >= 3 ? ty1 : ty0, txx = min(cellA + (t % 3u), 7u);
                s_cdf[i] = cl.cdf[((size_t)tyy * 8 + txx) * 256 + bin];
            }
            // bilinear-form table of both cells, 8 replicas per entry
            float4* s_quad = reinterpret_cast<float4*>(smem + L.quad);
            for (uint32_t i = tid; i < 2 * hp::kQuadEntries; i += NT) {
                const uint32_t cs = i / hp::kQuadEntries, bin = i % hp::kQuadEntries;
                const uint32_t pcx = cs ? cellB : cellA;
                float4 qv;
                if (bin == 256) {
                    qv = make_float4(hp::kM0 + 0.5f, 0.f, 0.f, 0.f);   // invalid pixel: sample 0 (autoscale.rs:604)
                } else {
                    const uint32_t p1 = pcx + 1 < 8 ? pcx + 1 : 7;
                    const double c00 = cl.cdf[((size_t)ty0 * 8 + pcx) * 256 + bin], c01 = cl.cdf[((size_t)ty0 * 8 + p1) * 256 + bin];
                    const double c10 = cl.cdf[((size_t)ty1 * 8 + pcx) * 256 + bin], c11 = cl.cdf[((size_t)ty1 * 8 + p1) * 256 + bin];
                    if (c00 == 0.0 && c01 == 0.0 && c10 == 0.0 && c11 == 0.0) {
                        qv = make_float4(hp::kM0 + 0.5f, 0.f, 0.f, 0.f);   // 0*x + 0*y == 0 exactly
                    } else if (c00 == 1.0 && c01 == 1.0 && c10 == 1.0 && c11 == 1.0 && sat_ok[cs]) {
                        qv = make_float4(hp::kM0 + 255.5f, 0.f, 0.f, 0.f);
                    } else {
                        qv.x = (float)(255.0 * c00 + (double)hp::kM0 + (double)hp::kShift);
                        qv.y = (float)(255.0 * (c01 - c00));
                        qv.z = (float)(255.0 * (c10 - c00));
                        qv.w = (float)(255.0 * ((c11 - c10) - (c01 - c00)));
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) s_quad[(cs * hp::kQuadEntries + bin) * 8 + r] = qv;
            }
            float* s_dy = reinterpret_cast<float*>(smem + L.dy);
            for (uint32_t i = tid; i < ((rb.y - rb.x + 3) & ~3u); i += NT)
                s_dy[i] = (rb.x + i < rb.y) ? (float)cl.row_dy[rb.x + i] : 0.f;
        }
        __syncthreads();

        // ---- per-thread produce state -----------------------------------------------------------
        const bool have_vec = stid < st.nvec;
        const uint32_t c0 = st.sc0 + stid * 8;
        const bool full_vec = have_vec && aligned && c0 + 8 <= a.src_cols;
        const unsigned char* const col_base = reinterpret_cast<const unsigned char*>(a.src) + (size_t)c0 * 2;
        float cdx[8];
        uint32_t cellsel = 0;
        bool slow_vec = have_vec && !full_vec; // edge / unaligned vectors: every row through the exact path
        if (CLAHE) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t cc = c0 + k < a.src_cols ? c0 + k : a.src_cols - 1;
                cdx[k] = (float)a.clahe.col_dx[cc];
            }
            if (have_vec) {
                const uint32_t cf = a.clahe.col_t[min(c0, a.src_cols - 1)] & 7u, cl_ = a.clahe.col_t[min(c0 + 7, a.src_cols - 1)] & 7u;
                cellsel = cf - cellA;
                if (cf != cl_) slow_vec = true; // the vector straddles a cell boundary
            }
        }

        // groups of this sub-block: g_i = rb.x + 4*(sub + NSUB*i)
        const uint32_t n_groups_all = (rb.y - rb.x + hp::kRows - 1) / hp::kRows;
        const uint32_t n_it = n_groups_all > sub ? (n_groups_all - sub + NSUB - 1) / NSUB : 0;
        auto group_row = [&](uint32_t i) { return rb.x + hp::kRows * (sub + NSUB * i); };

        uint4 q[hp::kRows];
        auto prefetch = [&](uint32_t g) {
            const unsigned char* p = col_base + (size_t)g * row_pitch;
            if (full_vec && g + hp::kRows <= rb.y) {
#pragma unroll
                for (int rr = 0; rr < hp::kRows; ++rr) q[rr] = ld_stream_u4(p + (size_t)rr * row_pitch);
            } else {
#pragma unroll
                for (int rr = 0; rr < hp::kRows; ++rr) {
                    q[rr] = make_uint4(0, 0, 0, 0);
                    if (full_vec && g + rr < rb.y) q[rr] = ld_stream_u4(p + (size_t)rr * row_pitch);
                }
            }
        };
        auto push = [&](uint32_t it, uint32_t rr) -> bool {
            uint32_t* qu = s_queue + (it % 3) * 256;
            const uint32_t slot = atomicAdd(&qu[255], 1u);
            if (slot < hp::kQueueCap) { qu[slot] = stid | (rr << 8); return true; }
            return false;
        };
        // exact sample of pixel (r, c): one round of independent global loads, tables from shared memory
        auto exact_px = [&](uint32_t r, uint32_t c) -> uint32_t {
            if (c >= a.src_cols || r >= rb.y) return 0u;
            const uint32_t d = reinterpret_cast<const uint16_t*>(a.src)[(size_t)r * a.src_cols + c];
            const uint32_t word = reinterpret_cast<const uint32_t*>(smem + L.lut)[(size_t)min(d, hot - 1u) << (lut_shift - 2)];
            if (!CLAHE) return word;
            const ClaheDev& cl = a.clahe;
            const uint32_t tx = cl.col_t[c];
            const double dx = cl.col_dx[c], omdx = cl.col_omdx[c], dy = cl.row_dy[r], omdy = cl.row_omdy[r];
            uint32_t o = 0;
            if (d) {
                const uint32_t bin = (word - (sbase + L.quad)) >> 7;
                const double* s_cdf = reinterpret_cast<const double*>(smem + L.cdf);
                const uint32_t x0 = ((tx & 7u) - cellA) * 256u + bin, x1 = (((tx >> 8) & 7u) - cellA) * 256u + bin;
                double v = clahe_blend_exact_rn(s_cdf[x0], s_cdf[x1], s_cdf[768 + x0], s_cdf[768 + x1], dx, omdx, dy, omdy);
                v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
                o = (uint32_t)__dmul_rn(v, 255.0);
            }
            mn_e = min(mn_e, o);
            mx_e = max(mx_e, o);
            return o;
        };

        if (n_it) prefetch(group_row(0));
        for (uint32_t it = 0; it < n_it + 2; ++it) {
            // ================= produce(group it) =================
            if (it < n_it) {
                const uint32_t g = group_row(it);
                uint4* const stg = s_stage + (size_t)(it % 3) * a.rbw_words;
                if (have_vec) {
                    uint32_t w0[hp::kRows], w1[hp::kRows];
                    if (slow_vec) {
#pragma unroll
                        for (int rr = 0; rr < hp::kRows; ++rr) {
                            w0[rr] = 0; w1[rr] = 0;
                            if (g + rr < rb.y && !push(it, rr)) { // queue full: resolve in place
#pragma unroll 1
                                for (int k = 0; k < 8; ++k) {
                                    const uint32_t o = exact_px(g + rr, c0 + k);
                                    if (k < 4) w0[rr] |= o << (8 * k); else w1[rr] |= o << (8 * (k - 4));
                                }
                            }
                        }
                    } else if (!CLAHE) {
#pragma unroll
                        for (int rr = 0; rr < hp::kRows; ++rr) {
                            const uint32_t wv[4] = {q[rr].x, q[rr].y, q[rr].z, q[rr].w};
                            uint32_t pr[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t a2 = __vminu2(wv[j], cap2) * lut_mul + cj;
                                const uint32_t e0 = lds_u32(a2 & 0xffffu), e1 = lds_u32(a2 >> 16);
                                pr[j] = __byte_perm(e0, e1, 0x5410);
                            }
                            w0[rr] = __byte_perm(pr[0], pr[1], 0x6420);
                            w1[rr] = __byte_perm(pr[2], pr[3], 0x6420);
                        }
                    } else {
                        const float* s_dy = reinterpret_cast<const float*>(smem + L.dy);
                        // cell_tag 0 / 1: the warp lies in one cell (table offset is an immediate); 2: the warp holds the
                        // cell boundary, every lane adds its own table offset (one more instruction per pixel)
                        auto rows_fast = [&](auto cell_tag) {
                            constexpr uint32_t OFF = decltype(cell_tag)::value == 1 ? hp::kQuadCellBytes : 0u;
                            const uint32_t dyn = decltype(cell_tag)::value == 2 ? cellsel * hp::kQuadCellBytes : 0u;
#pragma unroll
                            for (int rr = 0; rr < hp::kRows; ++rr) {
                                const float dy = s_dy[g - rb.x + rr];
                                const uint32_t wv[4] = {q[rr].x, q[rr].y, q[rr].z, q[rr].w};
                                uint32_t pr[4];
                                uint32_t risk = 0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const uint32_t a2 = __vminu2(wv[j], cap2) * lut_mul + cj;
                                    const uint32_t e0 = lds_u32(a2 & 0xffffu), e1 = lds_u32(a2 >> 16);
                                    const float4 q0 = lds_f4<OFF>(e0 + dyn), q1 = lds_f4<OFF>(e1 + dyn);
                                    const float u0 = __fmaf_rn(__fmaf_rn(q0.w, cdx[2 * j], q0.z), dy, __fmaf_rn(q0.y, cdx[2 * j], q0.x));
                                    const float u1 = __fmaf_rn(__fmaf_rn(q1.w, cdx[2 * j + 1], q1.z), dy, __fmaf_rn(q1.y, cdx[2 * j + 1], q1.x));
                                    risk |= (__float_as_uint(u0) ^ __float_as_uint(__fsub_rn(u0, hp::kGuard))) |
                                            (__float_as_uint(u1) ^ __float_as_uint(__fsub_rn(u1, hp::kGuard)));
                                    pr[j] = __byte_perm(__float_as_uint(__fadd_rd(u0, hp::kBigC)), __float_as_uint(__fadd_rd(u1, hp::kBigC)), 0x5410);
                                }
                                const bool risky = risk >= 8192u; // some integer part differs between u and u - 6ulp
                                if (!risky) {
                                    mn2 = __vimin3_u16x2(mn2, __vimin3_u16x2(pr[0], pr[1], pr[2]), pr[3]);
                                    mx2 = __vimax3_u16x2(mx2, __vimax3_u16x2(pr[0], pr[1], pr[2]), pr[3]);
                                }
                                const uint32_t k0 = __viaddmin_s16x2_relu(pr[0], 0xFE00FE00u, 0x00FF00FFu);
                                const uint32_t k1 = __viaddmin_s16x2_relu(pr[1], 0xFE00FE00u, 0x00FF00FFu);
                                const uint32_t k2 = __viaddmin_s16x2_relu(pr[2], 0xFE00FE00u, 0x00FF00FFu);
                                const uint32_t k3 = __viaddmin_s16x2_relu(pr[3], 0xFE00FE00u, 0x00FF00FFu);
                                w0[rr] = __byte_perm(k0, k1, 0x6420);
                                w1[rr] = __byte_perm(k2, k3, 0x6420);
                                if (risky && !push(it, rr)) { // queue full (never in practice): resolve in place
                                    w0[rr] = 0; w1[rr] = 0;
#pragma unroll 1
                                    for (int k = 0; k < 8; ++k) {
                                        const uint32_t o = exact_px(g + rr, c0 + k);
                                        if (k < 4) w0[rr] |= o << (8 * k); else w1[rr] |= o << (8 * (k - 4));
                                    }
                                }
                            }
                        };
                        const uint32_t in_b = __ballot_sync(__activemask(), cellsel != 0), act = __activemask();
                        if (in_b == 0) rows_fast(std::integral_constant<uint32_t, 0>{});
                        else if (in_b == act) rows_fast(std::integral_constant<uint32_t, 1>{});
                        else rows_fast(std::integral_constant<uint32_t, 2>{});
                    }
                    stg[stid * 2] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
                    stg[stid * 2 + 1] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
                }
                if (it + 1 < n_it) prefetch(group_row(it + 1));
            }
            // ================= fix(group it-1): the warps without output columns go first =================
            if (it >= 1 && it <= n_it) {
                const uint32_t g = group_row(it - 1);
                unsigned char* const stgb = reinterpret_cast<unsigned char*>(s_stage + (size_t)((it - 1) % 3) * a.rbw_words);
                const uint32_t* qu = s_queue + ((it - 1) % 3) * 256;
                const uint32_t nq = min(qu[255], hp::kQueueCap);
                for (uint32_t i = 255u - stid; i < nq * 8; i += hp::kHalf) {
                    const uint32_t e = qu[i >> 3], k = i & 7u;
                    const uint32_t vt = e & 255u, rr = e >> 8;
                    const uint32_t o = exact_px(g + rr, st.sc0 + vt * 8 + k);
                    stgb[(size_t)(vt * 2 + (k >> 2)) * 16 + rr * 4 + (k & 3u)] = (uint8_t)o;
                }
            }
            // the queue pushed to two iterations ago has been consumed (it is pushed to again next iteration)
            if (it >= 2 && stid == 0) s_queue[((it - 2) % 3) * 256 + 255] = 0;
            // ================= accumulate(group it-2) =================
            if (it >= 2 && have_ox) {
                const uint32_t g = group_row(it - 2);
                const uint4* const stg = s_stage + (size_t)((it - 2) % 3) * a.rbw_words;
                int acc[hp::kRows];
#pragma unroll
                for (int rr = 0; rr < hp::kRows; ++rr) acc[rr] = prec > 0 ? (1 << (prec - 1)) : 0;
#pragma unroll
                for (int m = 0; m < MAXP / 2; ++m) { // taps beyond ax.pairs are zero: no per-step test
                    const uint4 w = stg[woff + m];
                    acc[0] = dp2a_hi_su(taps[2 * m + 1], w.x, dp2a_lo_su(taps[2 * m], w.x, acc[0]));
                    acc[1] = dp2a_hi_su(taps[2 * m + 1], w.y, dp2a_lo_su(taps[2 * m], w.y, acc[1]));
                    acc[2] = dp2a_hi_su(taps[2 * m + 1], w.z, dp2a_lo_su(taps[2 * m], w.z, acc[2]));
                    acc[3] = dp2a_hi_su(taps[2 * m + 1], w.w, dp2a_lo_su(taps[2 * m], w.w, acc[3]));
                }
#pragma unroll
                for (int rr = 0; rr < hp::kRows; ++rr)
                    if (g + rr < rb.y) {
                        int v = acc[rr] >> prec;
                        v = v < 0 ? 0 : (v > 255 ? 255 : v);
                        reinterpret_cast<uint8_t*>(a.temp)[(size_t)(g + rr - a.row0) * a.ax.out_size + ox] = (uint8_t)v;
                    }
            }
            bar_half(1 + sub);
        }
    }
    if (CLAHE && a.minmax) {
        // fast-path extrema: biased by 512 and not yet clamped
        if (mn2 != 0xffffffffu) {
            const int lo = (int)min(mn2 & 0xffffu, mn2 >> 16) - 512, hi = (int)max(mx2 & 0xffffu, mx2 >> 16) - 512;
            mn_e = min(mn_e, (uint32_t)(lo < 0 ? 0 : (lo > 255 ? 255 : lo)));
            mx_e = max(mx_e, (uint32_t)(hi < 0 ? 0 : (hi > 255 ? 255 : hi)));
        }
        const uint32_t mnw = warp_reduce_min(mn_e), mxw = warp_reduce_max(mx_e);
        if (lane == 0 && mnw != 0xffffffffu) {
            atomicMin(&a.minmax[0], mnw);
            atomicMax(&a.minmax[1], mxw);
        }
    }
}

// =================================================================================================
// Warp-specialised variant. A sub-block is 8 producer warps (one 8-sample vector column per thread, no taps) and
// 4 consumer warps (one output column per thread, taps in registers; they also run the exact fix-up, which they
// have time for). Producers and consumers hand the triple-buffered staging over with producer/consumer named
// barriers (bar.arrive / bar.sync), so no warp carries both kinds of work and nobody waits for the slowest warp
// of a group: 24 warps per SM at 80 registers instead of 16 at 128. The fix-up needs no global round trip for
// pixel data: queue entries carry the 8 DNs, and the strip's f64 column geometry is staged in shared memory.
// =================================================================================================
namespace hs {
constexpr uint32_t kProd = 256, kCons = 128, kSub = kProd + kCons, kSubs = 2, kThreads = kSub * kSubs;
constexpr uint32_t kEntryWords = 5, kQueueCap = 100, kQueueWords = 512; // id + 4 words of DN; 100 entries + counter in 2 KB
}
__device__ __forceinline__ void bar_sync_n(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(uint32_t id, uint32_t n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct HSpecSmem {
    uint32_t lut, quad, cdf, dy, cdx, comdx, ct, rowgeo, queue, stage, total;
};
__host__ inline HSpecSmem hspec_layout(bool clahe, uint32_t lut_bytes, uint32_t max_rows, uint32_t max_nvec, uint32_t rbw_words) {
    HSpecSmem L;
    L.lut = 0;
    uint32_t o = lut_bytes;
    L.quad = o;
    if (clahe) o += 2 * hp::kQuadCellBytes;
    L.cdf = o;
    if (clahe) o += 6 * 256 * 8;
    L.cdx = o;
    if (clahe) o += max_nvec * 8 * 8;
    L.comdx = o;
    if (clahe) o += max_nvec * 8 * 8;
    L.dy = o;
    if (clahe) o += ((max_rows + 3) & ~3u) * 4u;
    L.ct = o;
    if (clahe) o += ((max_nvec * 8 * 2 + 15) & ~15u);
    L.rowgeo = o; // per sub-block and staging buffer: dy[4], omdy[4] (f64) of the group's rows
    if (clahe) o += hs::kSubs * 3 * 64;
    L.queue = o;
    o += hs::kSubs * 3 * hs::kQueueWords * 4;
    o = (o + 15) & ~15u;
    L.stage = o;
    o += hs::kSubs * 3 * rbw_words * 16u;
    L.total = o;
    return L;
}
struct HSpecParams {
    const HStrip* strips;
    const HPiece* pieces;
    const uint32_t* cta_first;
    uint32_t strip_w, hot, lut_shift, lut_mul, cap2, rbw_words;
    HSpecSmem L;
};

template <bool CLAHE, int MAXP>
__global__ void __launch_bounds__(hs::kThreads, 1) k_hspec(HResizeArgs a, HSpecParams pp) {
    extern __shared__ uint4 smem4[];
    unsigned char* const smem = reinterpret_cast<unsigned char*>(smem4);
    constexpr uint32_t NT = hs::kThreads;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    // block = (384, 2): threadIdx.y is the sub-block, threadIdx.x < 256 are its producers
    const uint32_t sub = threadIdx.y, lt = threadIdx.x, tid = sub * hs::kSub + lt, lane = lt & 31u;
    const bool producer = lt < hs::kProd;
    const uint32_t bar_ready = 1 + 7 * sub, bar_free = bar_ready + 3, bar_cons = bar_ready + 6;
    if (sbase + (pp.hot << pp.lut_shift) > 65536u) __trap();

    {   // once per CTA: DN -> table word, R lane-interleaved replicas
        uint4* s_lut4 = reinterpret_cast<uint4*>(smem + pp.L.lut);
        const uint32_t per = 1u << (pp.lut_shift - 4);
        for (uint32_t i = tid; i < pp.hot * per; i += NT) {
            const uint32_t idx = i >> (pp.lut_shift - 4), r0 = (i & (per - 1)) * 4u;
            const uint32_t e = idx + 1 == pp.hot ? a.hot_top : (a.lut[idx] & 255u);
            uint4 v;
            if (CLAHE) {
                const uint32_t bin = idx ? e : 256u;
                const uint32_t b = sbase + pp.L.quad + (bin * 8u + (r0 & 7u)) * 16u;
                v = make_uint4(b, b + 16u, b + 32u, b + 48u);
            } else {
                v = make_uint4(e, e, e, e);
            }
            s_lut4[i] = v;
        }
    }
    uint32_t mn2 = 0xffffffffu, mx2 = 0u;   // fast path: u16x2 running min / max of floor(u) - 1536 + 512
    uint32_t mn_e = 0xffffffffu, mx_e = 0;  // exact-path samples
    uint32_t tab_strip = 0xffffffffu;

    for (uint32_t pi = pp.cta_first[blockIdx.x]; pi < pp.cta_first[blockIdx.x + 1]; ++pi) {
        const HPiece pc = pp.pieces[pi];
        const HStrip st = pp.strips[pc.strip];
        const uint2 rb = make_uint2(pc.r0, pc.r1);
        uint4* const s_stage = reinterpret_cast<uint4*>(smem + pp.L.stage) + (size_t)sub * 3 * pp.rbw_words;
        uint32_t* const s_queue = reinterpret_cast<uint32_t*>(smem + pp.L.queue) + sub * 3 * hs::kQueueWords;
        __syncthreads(); // the previous piece is done with the tables and the staging buffers

        // ---- per-piece setup (all threads) -------------------------------------------------------
        for (uint32_t i = lt; i < 3 * pp.rbw_words; i += hs::kSub) s_stage[i] = make_uint4(0, 0, 0, 0);
        if (lt < 3) s_queue[lt * hs::kQueueWords + hs::kQueueWords - 1] = 0;
        uint32_t cellA = 0;
        if (CLAHE) {
            const ClaheDev& cl = a.clahe;
            const uint32_t ty = cl.row_t[rb.x];
            const uint32_t ty0 = ty & 7u, ty1 = (ty >> 8) & 7u;
            const uint32_t last_col = min(st.sc0 + st.nvec * 8u, a.src_cols) - 1u;
            cellA = cl.col_t[st.sc0] & 7u;
            const uint32_t cellB = cl.col_t[last_col] & 7u;
            if (tab_strip != pc.strip) { // f64 column geometry of the strip, for the exact path
                double* s_cdx = reinterpret_cast<double*>(smem + pp.L.cdx);
                double* s_comdx = reinterpret_cast<double*>(smem + pp.L.comdx);
                uint16_t* s_ct = reinterpret_cast<uint16_t*>(smem + pp.L.ct);
                for (uint32_t i = tid; i < st.nvec * 8u; i += NT) {
                    const uint32_t c = min(st.sc0 + i, a.src_cols - 1);
                    s_cdx[i] = cl.col_dx[c];
                    s_comdx[i] = cl.col_omdx[c];
                    s_ct[i] = cl.col_t[c];
                }
                tab_strip = pc.strip;
            }
            int okA = 1, okB = 1, okR = 1;
            for (uint32_t c = st.sc0 + tid; c <= last_col; c += NT) {
                const uint32_t ct = cl.col_t[c];
                if (!(ct & 0x80u)) { if ((ct & 7u) == cellA) okA = 0; else okB = 0; }
            }
            for (uint32_t r = rb.x + tid; r < rb.y; r += NT) if (cl.row_sat[r] != 255u) okR = 0;
            okA = __syncthreads_and(okA);
            okB = __syncthreads_and(okB);
            okR = __syncthreads_and(okR);
            const bool sat_ok[2] = {okA && okR, okB && okR};
            double* s_cdf = reinterpret_cast<double*>(smem + pp.L.cdf);
            for (uint32_t i = tid; i < 6 * 256; i += NT) {
                const uint32_t t = i >> 8, bin = i & 255u;
                const uint32_t tyy = t >= 3 ? ty1 : ty0, txx = min(cellA + (t % 3u), 7u);
                s_cdf[i] = cl.cdf[((size_t)tyy * 8 + txx) * 256 + bin];
            }
            float4* s_quad = reinterpret_cast<float4*>(smem + pp.L.quad);
            for (uint32_t i = tid; i < 2 * hp::kQuadEntries; i += NT) {
                const uint32_t cs = i / hp::kQuadEntries, bin = i % hp::kQuadEntries;
                const uint32_t pcx = cs ? cellB : cellA;
                float4 qv;
                if (bin == 256) {
                    qv = make_float4(hp::kM0 + 0.5f, 0.f, 0.f, 0.f);   // invalid pixel: sample 0 (autoscale.rs:604)
                } else {
                    const uint32_t p1 = pcx + 1 < 8 ? pcx + 1 : 7;
                    const double c00 = cl.cdf[((size_t)ty0 * 8 + pcx) * 256 + bin], c01 = cl.cdf[((size_t)ty0 * 8 + p1) * 256 + bin];
                    const double c10 = cl.cdf[((size_t)ty1 * 8 + pcx) * 256 + bin], c11 = cl.cdf[((size_t)ty1 * 8 + p1) * 256 + bin];
                    if (c00 == 0.0 && c01 == 0.0 && c10 == 0.0 && c11 == 0.0) {
                        qv = make_float4(hp::kM0 + 0.5f, 0.f, 0.f, 0.f);   // 0*x + 0*y == 0 exactly
                    } else if (c00 == 1.0 && c01 == 1.0 && c10 == 1.0 && c11 == 1.0 && sat_ok[cs]) {
                        qv = make_float4(hp::kM0 + 255.5f, 0.f, 0.f, 0.f);
                    } else {
                        qv.x = (float)(255.0 * c00 + (double)hp::kM0 + (double)hp::kShift);
                        qv.y = (float)(255.0 * (c01 - c00));
                        qv.z = (float)(255.0 * (c10 - c00));
                        qv.w = (float)(255.0 * ((c11 - c10) - (c01 - c00)));
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) s_quad[(cs * hp::kQuadEntries + bin) * 8 + r] = qv;
            }
            float* s_dy = reinterpret_cast<float*>(smem + pp.L.dy);
            for (uint32_t i = tid; i < ((rb.y - rb.x + 3) & ~3u); i += NT)
                s_dy[i] = (rb.x + i < rb.y) ? (float)cl.row_dy[rb.x + i] : 0.f;
        }
        __syncthreads();

        const uint32_t n_groups_all = (rb.y - rb.x + hp::kRows - 1) / hp::kRows;
        const uint32_t n_it = n_groups_all > sub ? (n_groups_all - sub + hs::kSubs - 1) / hs::kSubs : 0;
        auto group_row = [&](uint32_t i) { return rb.x + hp::kRows * (sub + hs::kSubs * i); };

        // exact sample of local column lc (strip-relative) in row r with DN d
        auto exact_sample = [&](uint32_t lc, uint32_t d, double dy, double omdy) -> uint32_t {
            const uint32_t word = reinterpret_cast<const uint32_t*>(smem + pp.L.lut)[(size_t)min(d, pp.hot - 1u) << (pp.lut_shift - 2)];
            if (!CLAHE) return word;
            uint32_t o = 0;
            if (d) {
                const uint32_t tx = reinterpret_cast<const uint16_t*>(smem + pp.L.ct)[lc];
                const double dx = reinterpret_cast<const double*>(smem + pp.L.cdx)[lc];
                const double omdx = reinterpret_cast<const double*>(smem + pp.L.comdx)[lc];
                const uint32_t bin = (word - (sbase + pp.L.quad)) >> 7;
                const double* s_cdf = reinterpret_cast<const double*>(smem + pp.L.cdf);
                const uint32_t x0 = ((tx & 7u) - cellA) * 256u + bin, x1 = (((tx >> 8) & 7u) - cellA) * 256u + bin;
                double v = clahe_blend_exact_rn(s_cdf[x0], s_cdf[x1], s_cdf[768 + x0], s_cdf[768 + x1], dx, omdx, dy, omdy);
                v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
                o = (uint32_t)__dmul_rn(v, 255.0);
            }
            mn_e = min(mn_e, o);
            mx_e = max(mx_e, o);
            return o;
        };
        // pixel (r, strip column lc) read from the raster; columns beyond the raster repeat the last one (they carry
        // zero taps, and must not disturb the min / max), rows beyond the piece are never stored
        auto exact_px = [&](uint32_t r, uint32_t lc) -> uint32_t {
            if (r >= rb.y) return 0u;
            const uint32_t c = min(st.sc0 + lc, a.src_cols - 1);
            double dy = 0.0, omdy = 0.0;
            if (CLAHE) { dy = a.clahe.row_dy[r]; omdy = a.clahe.row_omdy[r]; }
            return exact_sample(c - st.sc0, reinterpret_cast<const uint16_t*>(a.src)[(size_t)r * a.src_cols + c], dy, omdy);
        };

        if (producer) {
            // ================= producers =================
            const uint32_t ptid = lt;
            const bool have_vec = ptid < st.nvec;
            const uint32_t c0 = st.sc0 + ptid * 8;
            const bool aligned = (reinterpret_cast<uintptr_t>(a.src) % 16 == 0) && (a.src_cols % 8 == 0);
            const bool full_vec = have_vec && aligned && c0 + 8 <= a.src_cols;
            const uint32_t stride8 = a.src_cols >> 3; // uint4 per row (only used when aligned)
            const uint4* const src4 = reinterpret_cast<const uint4*>(a.src);
            const uint32_t cj = keep((sbase + pp.L.lut + (lane & ((pp.lut_mul >> 2) - 1u)) * 4u) * 0x10001u);
            float cdx[8];
            uint32_t cellsel = 0;
            bool slow_vec = have_vec && !full_vec; // edge / unaligned vectors: every row through the exact path
            if (CLAHE) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t cc = c0 + k < a.src_cols ? c0 + k : a.src_cols - 1;
                    cdx[k] = (float)a.clahe.col_dx[cc];
                }
                if (have_vec) {
                    const uint32_t cf = a.clahe.col_t[min(c0, a.src_cols - 1)] & 7u, cl_ = a.clahe.col_t[min(c0 + 7, a.src_cols - 1)] & 7u;
                    cellsel = cf - cellA;
                    if (cf != cl_) slow_vec = true; // the vector straddles a cell boundary
                }
            }
            uint4 q[hp::kRows];
            auto prefetch = [&](uint32_t g) {
                const uint4* p = src4 + (size_t)(g * stride8 + (c0 >> 3)); // vector index < 2^32 for any raster in HBM
                if (full_vec && g + hp::kRows <= rb.y) {
#pragma unroll
                    for (int rr = 0; rr < hp::kRows; ++rr) q[rr] = ld_stream_u4(p + (size_t)rr * stride8);
                } else {
#pragma unroll
                    for (int rr = 0; rr < hp::kRows; ++rr) {
                        q[rr] = make_uint4(0, 0, 0, 0);
                        if (full_vec && g + rr < rb.y) q[rr] = ld_stream_u4(p + (size_t)rr * stride8);
                    }
                }
            };
            // queue entry: id (vector | row << 8 | bit 31: DNs not included), then the row's 8 DNs
            auto push = [&](uint32_t k3, uint32_t rr, bool with_dn, const uint4& dn) -> bool {
                uint32_t* qu = s_queue + k3 * hs::kQueueWords;
                const uint32_t slot = atomicAdd(&qu[hs::kQueueWords - 1], 1u);
                if (slot >= hs::kQueueCap) return false;
                uint32_t* e = qu + slot * hs::kEntryWords;
                e[0] = ptid | (rr << 8) | (with_dn ? 0u : 0x80000000u);
                e[1] = dn.x; e[2] = dn.y; e[3] = dn.z; e[4] = dn.w;
                return true;
            };
            if (n_it) prefetch(group_row(0));
            for (uint32_t it = 0; it < n_it; ++it) {
                const uint32_t k3 = it % 3;
                const uint32_t g = group_row(it);
                uint4* const stg = s_stage + (size_t)k3 * pp.rbw_words;
                if (it >= 3) bar_sync_n(bar_free + k3, hs::kSub); // consumers are done with group it-3
                if (have_vec) {
                    uint32_t w0[hp::kRows], w1[hp::kRows];
                    if (slow_vec) {
#pragma unroll
                        for (int rr = 0; rr < hp::kRows; ++rr) {
                            w0[rr] = 0; w1[rr] = 0;
                            if (g + rr < rb.y && !push(k3, rr, full_vec, q[rr])) { // queue full: resolve in place
#pragma unroll 1
                                for (int k = 0; k < 8; ++k) {
                                    const uint32_t o = exact_px(g + rr, ptid * 8 + k);
                                    if (k < 4) w0[rr] |= o << (8 * k); else w1[rr] |= o << (8 * (k - 4));
                                }
                            }
                        }
                    } else if (!CLAHE) {
#pragma unroll
                        for (int rr = 0; rr < hp::kRows; ++rr) {
                            const uint32_t wv[4] = {q[rr].x, q[rr].y, q[rr].z, q[rr].w};
                            uint32_t pr[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t a2 = __vminu2(wv[j], pp.cap2) * pp.lut_mul + cj;
                                const uint32_t e0 = lds_u32(a2 & 0xffffu), e1 = lds_u32(a2 >> 16);
                                pr[j] = __byte_perm(e0, e1, 0x5410);
                            }
                            w0[rr] = __byte_perm(pr[0], pr[1], 0x6420);
                            w1[rr] = __byte_perm(pr[2], pr[3], 0x6420);
                        }
                    } else {
                        const float* s_dy = reinterpret_cast<const float*>(smem + pp.L.dy);
                        // cell_tag 0 / 1: the warp lies in one cell (table offset is an immediate); 2: the warp holds the
                        // cell boundary, every lane adds its own table offset (one more instruction per pixel)
                        auto rows_fast = [&](auto cell_tag) {
                            constexpr uint32_t OFF = decltype(cell_tag)::value == 1 ? hp::kQuadCellBytes : 0u;
                            const uint32_t dyn = decltype(cell_tag)::value == 2 ? cellsel * hp::kQuadCellBytes : 0u;
#pragma unroll
                            for (int rr = 0; rr < hp::kRows; ++rr) {
                                const float dy = s_dy[g - rb.x + rr];
                                const uint32_t wv[4] = {q[rr].x, q[rr].y, q[rr].z, q[rr].w};
                                uint32_t pr[4];
                                uint32_t risk = 0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const uint32_t a2 = __vminu2(wv[j], pp.cap2) * pp.lut_mul + cj;
                                    const uint32_t e0 = lds_u32(a2 & 0xffffu), e1 = lds_u32(a2 >> 16);
                                    const float4 q0 = lds_f4<OFF>(e0 + dyn), q1 = lds_f4<OFF>(e1 + dyn);
                                    const float u0 = __fmaf_rn(__fmaf_rn(q0.w, cdx[2 * j], q0.z), dy, __fmaf_rn(q0.y, cdx[2 * j], q0.x));
                                    const float u1 = __fmaf_rn(__fmaf_rn(q1.w, cdx[2 * j + 1], q1.z), dy, __fmaf_rn(q1.y, cdx[2 * j + 1], q1.x));
                                    risk |= (__float_as_uint(u0) ^ __float_as_uint(__fsub_rn(u0, hp::kGuard))) |
                                            (__float_as_uint(u1) ^ __float_as_uint(__fsub_rn(u1, hp::kGuard)));
                                    pr[j] = __byte_perm(__float_as_uint(__fadd_rd(u0, hp::kBigC)), __float_as_uint(__fadd_rd(u1, hp::kBigC)), 0x5410);
                                }
                                const bool risky = risk >= 8192u; // some integer part differs between u and u - 6ulp
                                if (!risky) {
                                    mn2 = __vimin3_u16x2(mn2, __vimin3_u16x2(pr[0], pr[1], pr[2]), pr[3]);
                                    mx2 = __vimax3_u16x2(mx2, __vimax3_u16x2(pr[0], pr[1], pr[2]), pr[3]);
                                }
                                const uint32_t k0 = __viaddmin_s16x2_relu(pr[0], 0xFE00FE00u, 0x00FF00FFu);
                                const uint32_t k1 = __viaddmin_s16x2_relu(pr[1], 0xFE00FE00u, 0x00FF00FFu);
                                const uint32_t k2 = __viaddmin_s16x2_relu(pr[2], 0xFE00FE00u, 0x00FF00FFu);
                                const uint32_t k3b = __viaddmin_s16x2_relu(pr[3], 0xFE00FE00u, 0x00FF00FFu);
                                w0[rr] = __byte_perm(k0, k1, 0x6420);
                                w1[rr] = __byte_perm(k2, k3b, 0x6420);
                                if (risky && !push(k3, rr, true, q[rr])) { // queue full (never in practice): resolve in place
                                    w0[rr] = 0; w1[rr] = 0;
#pragma unroll 1
                                    for (int k = 0; k < 8; ++k) {
                                        const uint32_t o = exact_px(g + rr, ptid * 8 + k);
                                        if (k < 4) w0[rr] |= o << (8 * k); else w1[rr] |= o << (8 * (k - 4));
                                    }
                                }
                            }
                        };
                        const uint32_t in_b = __ballot_sync(__activemask(), cellsel != 0), act = __activemask();
                        if (in_b == 0) rows_fast(std::integral_constant<uint32_t, 0>{});
                        else if (in_b == act) rows_fast(std::integral_constant<uint32_t, 1>{});
                        else rows_fast(std::integral_constant<uint32_t, 2>{});
                    }
                    stg[ptid * 2] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
                    stg[ptid * 2 + 1] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
                }
                if (it + 1 < n_it) prefetch(group_row(it + 1));
                __threadfence_block();
                bar_arrive_n(bar_ready + k3, hs::kSub);
            }
            // drain: every generation of a FREE barrier needs its 256 producer arrivals
            for (uint32_t it = n_it > 3 ? n_it - 3 : 0; it < n_it; ++it) bar_sync_n(bar_free + it % 3, hs::kSub);
        } else {
            // ================= consumers: exact fix-up, then the horizontal Lanczos taps =================
            const uint32_t ctid = lt - hs::kProd;
            const uint32_t ox = pc.strip * pp.strip_w + ctid;
            const bool have_ox = ctid < pp.strip_w && ox < a.ax.out_size;
            int taps[MAXP]; // (re)loaded per piece so that they are live in the consumer branch only
#pragma unroll
            for (int i = 0; i < MAXP; ++i)
                taps[i] = (have_ox && (uint32_t)i < a.ax.pairs) ? (int)a.ax.packed[(size_t)ox * a.ax.pairs + i] : 0;
            const uint32_t woff = have_ox ? ((((a.ax.start[ox]) & ~3u) - st.sc0) >> 2) : 0;
            const int prec = a.ax.precision;
            const int acc0 = prec > 0 ? (1 << (prec - 1)) : 0;
            for (uint32_t it = 0; it < n_it; ++it) {
                const uint32_t k3 = it % 3;
                const uint32_t g = group_row(it);
                uint4* const stg = s_stage + (size_t)k3 * pp.rbw_words;
                double* const s_rg = reinterpret_cast<double*>(smem + pp.L.rowgeo) + (sub * 3 + k3) * 8;
                if (CLAHE && ctid < 8) { // f64 row geometry of this group for the fix-up, fetched while waiting for the producers
                    const uint32_t r = min(g + (ctid & 3u), rb.y - 1);
                    s_rg[ctid] = ctid < 4 ? a.clahe.row_dy[r] : a.clahe.row_omdy[r];
                }
                bar_sync_n(bar_ready + k3, hs::kSub);
                uint32_t* qu = s_queue + k3 * hs::kQueueWords;
                const uint32_t nq = min(qu[hs::kQueueWords - 1], hs::kQueueCap);
                if (nq) {
                    unsigned char* const stgb = reinterpret_cast<unsigned char*>(stg);
                    for (uint32_t i = ctid; i < nq * 8; i += hs::kCons) {
                        const uint32_t* e = qu + (i >> 3) * hs::kEntryWords;
                        const uint32_t id = e[0], k = i & 7u;
                        const uint32_t vt = id & 255u, rr = (id >> 8) & 3u;
                        uint32_t o;
                        if (id & 0x80000000u) {
                            o = exact_px(g + rr, vt * 8 + k);
                        } else {
                            const uint32_t wd = e[1 + (k >> 1)];
                            o = exact_sample(vt * 8 + k, (k & 1u) ? (wd >> 16) : (wd & 0xffffu), CLAHE ? s_rg[rr] : 0.0, CLAHE ? s_rg[4 + rr] : 0.0);
                        }
                        stgb[(size_t)(vt * 2 + (k >> 2)) * 16 + rr * 4 + (k & 3u)] = (uint8_t)o;
                    }
                    bar_sync_n(bar_cons, hs::kCons);
                    if (ctid == 0) qu[hs::kQueueWords - 1] = 0;
                }
                if (have_ox) {
                    int acc[hp::kRows], acch[hp::kRows]; // two dependency chains per row
#pragma unroll
                    for (int rr = 0; rr < hp::kRows; ++rr) { acc[rr] = acc0; acch[rr] = 0; }
#pragma unroll
                    for (int m = 0; m < MAXP / 2; ++m) { // taps beyond ax.pairs are zero: no per-step test
                        const uint4 w = stg[woff + m];
                        acc[0] = dp2a_lo_su(taps[2 * m], w.x, acc[0]);
                        acc[1] = dp2a_lo_su(taps[2 * m], w.y, acc[1]);
                        acc[2] = dp2a_lo_su(taps[2 * m], w.z, acc[2]);
                        acc[3] = dp2a_lo_su(taps[2 * m], w.w, acc[3]);
                        acch[0] = dp2a_hi_su(taps[2 * m + 1], w.x, acch[0]);
                        acch[1] = dp2a_hi_su(taps[2 * m + 1], w.y, acch[1]);
                        acch[2] = dp2a_hi_su(taps[2 * m + 1], w.z, acch[2]);
                        acch[3] = dp2a_hi_su(taps[2 * m + 1], w.w, acch[3]);
                    }
#pragma unroll
                    for (int rr = 0; rr < hp::kRows; ++rr) acc[rr] += acch[rr];
                    uint8_t* const o = reinterpret_cast<uint8_t*>(a.temp) + (size_t)(g - a.row0) * a.ax.out_size + ox;
#pragma unroll
                    for (int rr = 0; rr < hp::kRows; ++rr)
                        if (g + rr < rb.y) {
                            int v = acc[rr] >> prec;
                            v = v < 0 ? 0 : (v > 255 ? 255 : v);
                            o[(size_t)rr * a.ax.out_size] = (uint8_t)v;
                        }
                }
                __threadfence_block();
                bar_arrive_n(bar_free + k3, hs::kSub);
            }
        }
    }
    if (CLAHE && a.minmax) {
        if (mn2 != 0xffffffffu) { // fast-path extrema: biased by 512 and not yet clamped
            const int lo = (int)min(mn2 & 0xffffu, mn2 >> 16) - 512, hi = (int)max(mx2 & 0xffffu, mx2 >> 16) - 512;
            mn_e = min(mn_e, (uint32_t)(lo < 0 ? 0 : (lo > 255 ? 255 : lo)));
            mx_e = max(mx_e, (uint32_t)(hi < 0 ? 0 : (hi > 255 ? 255 : hi)));
        }
        const uint32_t mnw = warp_reduce_min(mn_e), mxw = warp_reduce_max(mx_e);
        if (lane == 0 && mnw != 0xffffffffu) {
            atomicMin(&a.minmax[0], mnw);
            atomicMax(&a.minmax[1], mxw);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------
// Strips: the widest output strip whose source span fits max_vec 8-sample vectors (<= 256, and <= one CLAHE
// tile width so that a strip meets at most one cell boundary).
cudaError_t hpipe_build_strips(const uint32_t* start_h, const uint32_t* size_h, uint32_t out_size, uint32_t in_size,
                               uint32_t window, uint32_t max_vec, uint32_t max_w, uint32_t* strip_w_out,
                               std::vector<HStrip>* strips, uint32_t* rbw_words) {
    if (out_size == 0 || max_vec < 4) return cudaErrorInvalidConfiguration;
    max_vec = std::min(max_vec, 256u);
    auto span_ok = [&](uint32_t w) {
        for (uint32_t ox0 = 0; ox0 < out_size; ox0 += w) {
            const uint32_t ox1 = std::min(out_size, ox0 + w);
            uint32_t lo = ox0 == 0 ? 0 : start_h[ox0], hi = 0;
            for (uint32_t x = ox0; x < ox1; ++x) {
                lo = std::min(lo, start_h[x]);
                hi = std::max(hi, start_h[x] + size_h[x]);
            }
            if (ox1 == out_size) hi = in_size;
            if ((hi - (lo & ~7u) + 7) / 8 > max_vec) return false;
        }
        return true;
    };
    uint32_t w = std::min(std::min(256u, max_w), out_size);
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
    *rbw_words = max_nvec * 2 + (window + 16) / 4 + 2 + 4; // staged words + zero-tap overrun (packed table, MAXP padding)
    return cudaSuccess;
}

bool hpipe_supported(uint32_t pairs) { return pairs >= 2 && pairs <= 48; }

// Cuts the (strip, row) space into contiguous equal-weight runs, one per CTA. `cuts` are the row positions where
// a piece must end (vertical CLAHE cell boundaries; first = 0, last = rows). A piece's weight is rows x nvec.
// Pieces are multiples of `unit` rows from their segment start, so that the sub-blocks get whole groups.
void hpipe_build_pieces(const std::vector<HStrip>& strips, const std::vector<uint64_t>& cuts, uint32_t n_ctas, uint32_t unit,
                        std::vector<uint32_t>* pieces_flat, std::vector<uint32_t>* cta_first, uint32_t* max_rows) {
    struct Seg { uint32_t strip, r0, r1; uint64_t w; };
    std::vector<Seg> segs;
    uint64_t total = 0;
    for (uint32_t s = 0; s < strips.size(); ++s)
        for (size_t i = 0; i + 1 < cuts.size(); ++i) {
            if (cuts[i + 1] <= cuts[i]) continue;
            const uint64_t units = (cuts[i + 1] - cuts[i] + unit - 1) / unit;
            segs.push_back(Seg{s, (uint32_t)cuts[i], (uint32_t)cuts[i + 1], units * std::max(1u, strips[s].nvec)});
            total += segs.back().w;
        }
    n_ctas = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_ctas, total / std::max<uint64_t>(1, 64)) ); // >= 64 vector-units per CTA
    pieces_flat->clear();
    cta_first->assign(1, 0);
    *max_rows = 0;
    uint64_t done = 0; // weight handed out so far
    size_t si = 0;
    uint32_t r = segs.empty() ? 0 : segs[0].r0;
    for (uint32_t b = 0; b < n_ctas; ++b) {
        const uint64_t target = total * (b + 1) / n_ctas;
        while (si < segs.size() && (done < target || b + 1 == n_ctas)) {
            const Seg& sg = segs[si];
            const uint64_t nv = std::max(1u, strips[sg.strip].nvec);
            const uint64_t units_left = (sg.r1 - r + unit - 1) / unit;
            uint64_t take = b + 1 == n_ctas ? units_left : std::min<uint64_t>(units_left, (target - done + nv - 1) / nv);
            if (take == 0) break;
            const uint32_t r1 = (uint32_t)std::min<uint64_t>(sg.r1, (uint64_t)r + take * unit);
            pieces_flat->push_back(sg.strip);
            pieces_flat->push_back(r);
            pieces_flat->push_back(r1);
            pieces_flat->push_back(0);
            *max_rows = std::max(*max_rows, r1 - r);
            done += take * nv;
            r = r1;
            if (r >= sg.r1) {
                ++si;
                if (si < segs.size()) r = segs[si].r0;
            }
        }
        cta_first->push_back((uint32_t)(pieces_flat->size() / 4));
    }
}

template <bool CLAHE, int MAXP, int NSUB>
static cudaError_t launch_hpipe_t(const HResizeArgs& a, const HPipeParams& pp, uint32_t n_ctas, size_t smem, cudaStream_t stream) {
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_hpipe<CLAHE, MAXP, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    k_hpipe<CLAHE, MAXP, NSUB><<<n_ctas, NSUB * 256, smem, stream>>>(a, pp);
    return cudaGetLastError();
}

template <bool CLAHE, int NSUB>
static cudaError_t launch_hpipe_s(const HResizeArgs& a, const HPipeParams& pp, uint32_t n_ctas, size_t smem, cudaStream_t stream) {
    const uint32_t p = a.ax.pairs;
#define SARPRO_HP(P) return launch_hpipe_t<CLAHE, P, NSUB>(a, pp, n_ctas, smem, stream)
    if (p <= 8) SARPRO_HP(8);
    if (p <= 16) SARPRO_HP(16);
    if (p <= 24) SARPRO_HP(24);
    if (p <= 32) SARPRO_HP(32);
    if (p <= 40) SARPRO_HP(40);
    SARPRO_HP(48);
#undef SARPRO_HP
}

template <bool CLAHE>
static cudaError_t launch_hspec_s(const HResizeArgs& a, const HSpecParams& pp, uint32_t n_ctas, cudaStream_t stream) {
    const uint32_t p = a.ax.pairs;
    const size_t smem = pp.L.total;
#define SARPRO_HS(P)                                                                                                     \
    do {                                                                                                                 \
        static size_t configured = 0;                                                                                    \
        if (smem > configured) {                                                                                         \
            cudaError_t e = cudaFuncSetAttribute(k_hspec<CLAHE, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                              \
            configured = smem;                                                                                           \
        }                                                                                                                \
        k_hspec<CLAHE, P><<<n_ctas, dim3(hs::kSub, hs::kSubs), smem, stream>>>(a, pp);                                   \
        return cudaGetLastError();                                                                                       \
    } while (0)
    if (p <= 8) SARPRO_HS(8);
    if (p <= 16) SARPRO_HS(16);
    if (p <= 24) SARPRO_HS(24);
    if (p <= 32) SARPRO_HS(32);
    if (p <= 40) SARPRO_HS(40);
    SARPRO_HS(48);
#undef SARPRO_HS
}

uint32_t hpipe_lut_shift(uint32_t hot) { return hot <= 1000 ? 6u : 5u; } // 16 replicas when they fit 64 KB, else 8

size_t hpipe_smem_bytes(int src_kind, int nsub, uint32_t hot, uint32_t max_rows, uint32_t rbw_words) {
    if (nsub == 12) // warp-specialised kernel: rbw_words = staged words + overrun, so (rbw_words - overrun) / 2 >= nvec
        return hspec_layout(src_kind == HSRC_DN_CLAHE, hot << hpipe_lut_shift(hot), max_rows, rbw_words / 2, rbw_words).total;
    return hpipe_layout(src_kind == HSRC_DN_CLAHE, (uint32_t)nsub, hot << hpipe_lut_shift(hot), max_rows, rbw_words).total;
}

cudaError_t launch_hpipe(const HResizeArgs& a, int src_kind, int nsub, const HStrip* strips_dev, const uint32_t* pieces_dev,
                         const uint32_t* cta_first_dev, uint32_t n_ctas, uint32_t strip_w, uint32_t hot, uint32_t max_rows,
                         cudaStream_t stream) {
    if (a.n_rows == 0 || a.ax.out_size == 0 || n_ctas == 0) return cudaSuccess;
    const bool clahe = src_kind == HSRC_DN_CLAHE;
    const size_t smem = hpipe_smem_bytes(src_kind, nsub, hot, max_rows, a.rbw_words);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    HPipeParams pp;
    pp.strips = strips_dev;
    pp.pieces = reinterpret_cast<const HPiece*>(pieces_dev);
    pp.cta_first = cta_first_dev;
    pp.strip_w = strip_w;
    pp.hot = hot;
    pp.lut_shift = hpipe_lut_shift(hot);
    pp.max_rows = max_rows;
    if (nsub == 12) { // warp-specialised: 2 sub-blocks x (8 producer + 4 consumer warps)
        HSpecParams sp;
        sp.strips = strips_dev;
        sp.pieces = reinterpret_cast<const HPiece*>(pieces_dev);
        sp.cta_first = cta_first_dev;
        sp.strip_w = strip_w;
        sp.hot = hot;
        sp.lut_shift = hpipe_lut_shift(hot);
        sp.lut_mul = 1u << sp.lut_shift;
        sp.cap2 = (hot - 1u) * 0x10001u;
        sp.rbw_words = a.rbw_words;
        sp.L = hspec_layout(clahe, hot << sp.lut_shift, max_rows, a.rbw_words / 2, a.rbw_words);
        if (clahe) return launch_hspec_s<true>(a, sp, n_ctas, stream);
        return launch_hspec_s<false>(a, sp, n_ctas, stream);
    }
    if (nsub == 3) {
        if (clahe) return launch_hpipe_s<true, 3>(a, pp, n_ctas, smem, stream);
        return launch_hpipe_s<false, 3>(a, pp, n_ctas, smem, stream);
    }
    if (clahe) return launch_hpipe_s<true, 2>(a, pp, n_ctas, smem, stream);
    return launch_hpipe_s<false, 2>(a, pp, n_ctas, smem, stream);
}

} // namespace sarpro

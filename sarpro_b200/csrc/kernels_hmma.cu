// kernels_hmma.cu — pass B for u8 samples, third generation: the per-pixel stage (DN -> sample) fused with the
// horizontal Lanczos pass, with the taps on the tensor-core integer path.
//
// The horizontal pass (resize.rs:39-50 -> fast_image_resize, i16 taps, i32 accumulate) is a banded matrix product
//     temp[r][ox] = clamp((half + sum_c S[r][c] * W[c][ox]) >> precision),   S u8 samples, W i16 taps.
// An i16 tap is 256*hi + lo (hi signed byte, lo unsigned byte), so the sum is 256 * (S x Whi) + (S x Wlo): two exact
// u8 x s8 / u8 x u8 -> s32 products, which is what mma.sync.m16n8k32 (IMMA.16832) computes. Integer, hence bit-exact
// whatever the summation order.
//
// A warp owns 16 source rows and walks 64-column blocks (two k-steps of 32 columns) along a strip. Per k-step, lane
// (g = lane/4, q = lane%4) loads the 8 DNs of columns 32*ks + 8q .. +7 for rows g and g+8 (128-bit streaming loads,
// prefetched one block ahead into the registers just consumed), turns them into samples, and the packed samples ARE
// the A fragments: the k index of an MMA is only summed over, so the host lays the tap bytes of the B fragments out in
// the same permuted order (column 32*ks + 8q + 4r + i  <->  register r, byte i of lane q). No sample ever goes through
// shared memory; the strip's B fragments are staged there once per piece. An output n-tile (8 output columns) is live
// while the walk crosses its window (three accumulator slots, rotated), then it is scaled, clamped and stored.
//
// Per-pixel stage. LUT strategies: one shared-memory gather (R-way lane-interleaved table).
// CLAHE (autoscale.rs:307-330, :602): DN -> address of the pixel's bin entry in an 8-way replicated float4 table
// holding the bilinear form of the four tile CDFs of the cell, u = A + B*dx + (C + D*dx)*dy in sample units, three
// FFMAs in plain fp32 (|u| < 512, so its rounding error is ~2e-5). A carries +S where S bounds the total error of
// the evaluation (computed per table entry when the piece's tables are built), so the true value lies in
// [u - 2S, u]. One FADD.RD against 1.5*2^(23-F) truncates u to F fraction bits in the mantissa: if those bits are
// not all zero, u >= n + 2^-F > n + 2S and floor(true) == floor(u) == n. Otherwise (2^-F of the pixels, F = 13
// normally) the 8-pixel vector is flagged; the warp enumerates its flagged vectors and recomputes them four at a time,
// one pixel per lane, with the reference's exact f64 operation order, and patches the fragment registers before the
// MMA. Entries whose four CDF values are identical are resolved when the table is built; bins whose CDFs are all 1.0
// use a marker entry in the cells where the reference's f64 roundings decide between 254 and 255.
#include <algorithm>
#include <vector>

#include "clahe_exact.cuh"
#include "common.cuh"
#include "kernels.h"

namespace sarpro {

namespace hm {
#ifndef SARPRO_HMMA_CLAHE_THREADS
#define SARPRO_HMMA_CLAHE_THREADS 384
#endif
constexpr uint32_t kThreadsLut = 512, kThreadsClahe = SARPRO_HMMA_CLAHE_THREADS; // CLAHE: fewer warps, ~158 registers each for the gather pipeline
constexpr uint32_t kQuadEntries = 257;                       // 256 bins + the invalid-pixel entry
constexpr uint32_t kQuadCellBytes = kQuadEntries * 8 * 16;   // one cell, 8 replicas
constexpr int kSlots = 3;                                    // n-tiles in flight per warp
constexpr uint32_t kMaxStripB = 80 * 1024;                   // B fragments of one strip (shared memory)
constexpr uint32_t kMaxStripVecs = 1024;                     // 8-column vectors of one strip (128 blocks)
constexpr float kBigC = 12582912.0f + 512.0f;                // u + kBigC (RD): low 16 bits = floor(u) + 512
constexpr float kMarker = 480.0f;                            // floor of the marker entries (regular entries stay below 460)
constexpr uint32_t kMarkerLess2 = (480u + 512u - 1u) * 0x10001u;
} // namespace hm

__device__ __forceinline__ uint32_t hm_lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t hm_lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <uint32_t OFF>
__device__ __forceinline__ float4 hm_lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
// float4 at shared address a + off; off is warp-uniform, so the add folds into the LDS operand (R + UR)
__device__ __forceinline__ float4 hm_lds_f4u(uint32_t a, uint32_t off) {
    float4 v;
    asm volatile("{ .reg .u32 t; add.u32 t, %4, %5; ld.shared.v4.f32 {%0,%1,%2,%3}, [t]; }"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(a), "r"(off));
    return v;
}
__device__ __forceinline__ uint32_t hm_keep(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
// A loop invariant ptxas must hold in a register: it sees through hm_keep's mov and re-derives such values inside the
// loop (the packed table base as an extra IMAD by 0x10001 per pixel pair, the clamp bias as a MOV per VIADDMNMX); a
// value that went through a shuffle cannot be re-derived.
__device__ __forceinline__ uint32_t hm_pin(uint32_t v, uint32_t lane) {
    return __shfl_sync(0xffffffffu, v, (int)lane);
}
// a * b + c as one IMAD (left to the compiler, the multiply and the add end up in different basic blocks)
__device__ __forceinline__ uint32_t hm_mad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint4 hm_lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 hm_ldg_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// 128-bit streaming load, no L1 allocation, 256-byte L2 fetch granularity (the next block of the walk)
__device__ __forceinline__ uint4 hm_ld_dn(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void mma_u8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_u8u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct HMmaSmem {
    uint32_t lut, quad, cdf, ctrl, nt, cm, bfrag, total;
};
__host__ __device__ inline HMmaSmem hmma_layout(bool clahe, uint32_t lut_bytes, uint32_t b_bytes) {
    HMmaSmem L;
    L.lut = 0;
    uint32_t o = (lut_bytes + 15) & ~15u;
    L.quad = o;
    if (clahe) o += 2 * hm::kQuadCellBytes;
    L.cdf = o;
    if (clahe) o += 6 * 256 * 8;
    L.ctrl = o;
    o += 64;
    L.nt = o;   // n-tile table of the strip: {first k-step, last k-step, byte offset of its fragments - first*512, 0}
    o += 32 * 16;
    L.cm = o;   // per 8-column vector of the strip: columns with fl(omdx+dx) == 1.0 | (== 1 - 2^-53) << 8
    if (clahe) o += hm::kMaxStripVecs * 2;
    L.bfrag = o;
    o += b_bytes;
    L.total = o;
    return L;
}

struct HMmaParams {
    const uint4* btab;     // B fragments: [koff + (kstep - first kstep)][lane] = {hi r0, hi r1, lo r0, lo r1}
    const int4* ntile;     // per n-tile: {first k-step, last k-step, koff, 0}; a k-step is 32 source columns
    const uint4* strips;   // per strip: {first n-tile, end n-tile, first block, end block}; a block is two k-steps
    const HPiece* pieces;
    const uint32_t* cta_first;
    uint32_t b_bytes; // the largest strip's B fragments (staged in shared memory)
};

// 32 lane-private replicas (conflict-free gathers) when the table still fits the 16-bit address range, else 16 (<= 64 KB) or 8
__host__ __device__ inline uint32_t hmma_lut_shift(uint32_t hot) { return hot <= 500 ? 7u : (hot <= 1000 ? 6u : 5u); }

// Error bound of the fp32 evaluation of one table entry (see the header), in sample units. X, Y bound |dx|, |dy|.
__device__ __forceinline__ double hm_half_ulp(double x) {
    const double ax = fabs(x);
    if (ax < 1e-30) return 0.0;
    int e;
    frexp(ax, &e); // ax = m * 2^e, m in [0.5, 1)
    return ldexp(1.0, e - 25); // half an ulp of the binade of ax (fp32: 24 significant bits)
}
__device__ __forceinline__ double hm_entry_error(double A, double B, double C, double D) {
    const double X = 1.0, Y = 1.0;          // |dx| <= 1, |dy| <= 1 (autoscale.rs:308-318: d in [-0.5, 1))
    const double ex = 2.4e-7, ey = 3.1e-8;  // |dxf - dx|, |dyf - dy| (see dx_of / the row geometry below)
    const double aB = fabs(B), aC = fabs(C), aD = fabs(D), aA = fabs(A);
    double e = hm_half_ulp(A) + hm_half_ulp(B) * X + hm_half_ulp(C) * Y + hm_half_ulp(D) * X * Y; // table roundings
    e += (aB + aD * Y) * ex + (aC + aD * X) * ey;                                                 // geometry
    e += hm_half_ulp(aC + aD * X) * Y;                                                            // t1 = fl(D*dx + C)
    e += hm_half_ulp(aA + aB * X);                                                                // t2 = fl(B*dx + A)
    e += hm_half_ulp(aA + aB * X + (aC + aD * X) * Y);                                            // u  = fl(t1*dy + t2)
    return e * 1.05 + 1e-9; // second-order terms; the reference's own f64 roundings (< 1e-12)
}

template <bool CLAHE>
__global__ void __launch_bounds__(CLAHE ? hm::kThreadsClahe : hm::kThreadsLut, 1) k_hmma(HResizeArgs a, HMmaParams pp) {
    extern __shared__ uint4 smem4[];
    unsigned char* const smem = reinterpret_cast<unsigned char*>(smem4);
    constexpr uint32_t NT = CLAHE ? hm::kThreadsClahe : hm::kThreadsLut;
    constexpr uint32_t FULL = 0xffffffffu;
    if (a.skip && *a.skip) return;
    // the table range comes from the band's plan in device memory (device or host planner): no host round trip before this launch
    const uint32_t hot = a.plan->hot, hot_top = a.plan->hot_top;
    if (hot == 0 || a.plan->use_generic) return; // the generic exact kernel (launched behind this one) takes the band
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t lut_shift = hmma_lut_shift(hot); // table word of DN idx, replica r: byte (idx << lut_shift) + 4r
    const HMmaSmem L = hmma_layout(CLAHE, hot << lut_shift, pp.b_bytes);
    const uint32_t cols = a.src_cols;                          // row pitch and width the kernel walks (a multiple of 8)
    const uint32_t width = a.src_width ? a.src_width : cols;  // true raster width (re-pitched rasters: cols - width < 8)
    const uint32_t tid = threadIdx.x, lane = hm_keep(tid & 31u), g = hm_keep(lane >> 2), q = hm_keep(lane & 3u);
    if (sbase + (hot << lut_shift) > 65536u) __trap(); // 16-bit table addresses (dynamic smem starts low on sm_100)
    uint32_t* const s_ctrl = reinterpret_cast<uint32_t*>(smem + L.ctrl);

    {   // once per CTA: DN -> table word, R lane-interleaved replicas
    uint4* s_lut4 = reinterpret_cast<uint4*>(smem + L.lut);
    const uint32_t per = 1u << (lut_shift - 4); // uint4 per table entry
    for (uint32_t i = tid; i < hot * per; i += NT) {
        const uint32_t idx = i >> (lut_shift - 4), r0 = (i & (per - 1)) * 4u;
        const uint32_t e = idx + 1 == hot ? hot_top : (a.lut[idx] & 255u);
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
    const uint32_t cap2 = hm_keep((hot - 1u) * 0x10001u);
    const uint32_t lut_mul = hm_keep(1u << lut_shift);
    const uint32_t cj = hm_pin((sbase + L.lut + (lane & (lut_mul / 4u - 1u)) * 4u) * 0x10001u, lane);
    const int prec = a.ax.precision;
    const int acc0 = (int)hm_keep(prec > 0 ? (1u << (prec - 1)) : 0u);
    const uint16_t* const src = reinterpret_cast<const uint16_t*>(a.src);
    // scale_u16_to_u8 (autoscale.rs:348-364) takes min/max over ALL samples, invalid pixels (written as 0) included
    uint32_t mn2 = 0xffffffffu, mx2 = 0u;   // fast path: u16x2 running min / max of floor(u) + 512 (not yet clamped)
    uint32_t mn_e = 0xffffffffu, mx_e = 0;  // exact-path samples
    uint32_t staged_strip = 0xffffffffu;
    const uint32_t sb_lane = hm_keep(sbase + L.bfrag + lane * 16u);
    // 8-byte stores of the resized rows: whole n-tiles only and 8-byte aligned rows
    const bool wide_store = (a.ax.out_size & 7u) == 0 && (reinterpret_cast<uintptr_t>(a.temp) & 7u) == 0;
    const uint32_t cols8 = hm_pin(cols - 8u, lane);
    const uint32_t cm_base = hm_pin(sbase + L.cm + q * 2u, lane);
    const float inv2tw = __uint_as_float(hm_pin(__float_as_uint(CLAHE ? a.clahe.inv2tw : 0.f), lane));
    const float dstep = __uint_as_float(hm_pin(__float_as_uint(__fmul_rn(2.0f, inv2tw)), lane));
    const uint32_t relu_c = hm_pin(0xFE00FE00u ^ lane, lane) ^ lane; // -512 per half (a shuffled constant would be folded)

    for (uint32_t pi = pp.cta_first[blockIdx.x]; pi < pp.cta_first[blockIdx.x + 1]; ++pi) {
        const HPiece pc = pp.pieces[pi];
        const uint4 st = pp.strips[pc.strip]; // {j0, j1, cb0, cb1}
        __syncthreads(); // the previous piece is done with the tables
        if (tid == 0) { s_ctrl[0] = 0; s_ctrl[1] = 0xffffffffu; s_ctrl[2] = 0; }
        const uint32_t koff0 = (uint32_t)pp.ntile[st.x].z;
        if (staged_strip != pc.strip) { // the strip's B fragments
            const int4 ml = pp.ntile[st.y - 1u];
            const uint32_t n16 = ((uint32_t)ml.z + (uint32_t)(ml.y - ml.x + 1) - koff0) * 32u;
            uint4* s_b = reinterpret_cast<uint4*>(smem + L.bfrag);
            const uint4* gb = pp.btab + (size_t)koff0 * 32u;
            for (uint32_t i = tid; i < n16; i += NT) s_b[i] = hm_ldg_u4(gb + i);
            int4* s_nt = reinterpret_cast<int4*>(smem + L.nt);
            for (uint32_t i = tid; i < st.y - st.x; i += NT) {
                const int4 m = pp.ntile[st.x + i];
                s_nt[i] = make_int4(m.x, m.y, (int)(((uint32_t)m.z - koff0 - (uint32_t)m.x) * 512u), 0);
            }
            if (CLAHE) {
                uint16_t* s_cm = reinterpret_cast<uint16_t*>(smem + L.cm);
                for (uint32_t i = tid; i < (st.w - st.z) * 8u; i += NT) {
                    const uint32_t c = min(st.z * 64u + i * 8u, cols - 8u);
                    const uint4 ct = *reinterpret_cast<const uint4*>(a.clahe.col_t + c);
                    const uint32_t cw[4] = {ct.x, ct.y, ct.z, ct.w};
                    uint32_t m0 = 0, m1 = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        m0 |= (((cw[j] >> 7) & 1u) | ((cw[j] >> 22) & 2u)) << (2 * j);
                        m1 |= (((cw[j] >> 6) & 1u) | ((cw[j] >> 21) & 2u)) << (2 * j);
                    }
                    s_cm[i] = (uint16_t)(m0 | (m1 << 8));
                }
            }
            staged_strip = pc.strip;
        }

        // ---- per-piece tables --------------------------------------------------------------------
        uint32_t cellA = 0, bcol = 0xffffffffu, fbits = 13;
        float magic = 1536.0f;
        bool fixA = false, fixB = false; // saturated-entry markers in use for cell A / B
        if (CLAHE) {
            const ClaheDev& cl = a.clahe;
            const uint32_t ty = cl.row_t[pc.r0]; // the piece lies inside one vertical bilinear cell
            const uint32_t ty0 = ty & 7u, ty1 = (ty >> 8) & 7u;
            const uint32_t first_col = min(st.z * 64u, cols - 1u), last_col = min(st.w * 64u, cols) - 1u;
            cellA = cl.col_t[first_col] & 7u;
            const uint32_t cellB = cl.col_t[last_col] & 7u; // == cellA or cellA + 1 (strip span <= tile width)
            __syncthreads();
            // first column of cell B; saturated-bin shortcut (all four CDFs exactly 1.0 -> sample 255) is valid for a
            // cell only when fl(omdx+dx) == 1 for all of its columns in the strip and fl(omdy+dy) == 1 for all rows
            int okA = 1, okB = 1, okR = 1;
            for (uint32_t c = first_col + tid; c <= last_col; c += NT) {
                const uint32_t ct = cl.col_t[c];
                if ((ct & 7u) != cellA) atomicMin(&s_ctrl[1], c);
                if (!(ct & 0x80u)) { if ((ct & 7u) == cellA) okA = 0; else okB = 0; }
            }
            for (uint32_t r = pc.r0 + tid; r < pc.r1; r += NT) if (cl.row_sat[r] != 255u) okR = 0;
            okA = __syncthreads_and(okA);
            okB = __syncthreads_and(okB);
            okR = __syncthreads_and(okR);
            bcol = s_ctrl[1];
            const bool sat_ok[2] = {okA && okR, okB && okR};
            fixA = !sat_ok[0];
            fixB = !sat_ok[1];
            // f64 CDFs of the (up to) 3 x 2 tiles of the piece, for the exact path
            double* s_cdf = reinterpret_cast<double*>(smem + L.cdf);
            for (uint32_t i = tid; i < 6 * 256; i += NT) {
                const uint32_t t = i >> 8, bin = i & 255u;
                const uint32_t tyy = t >= 3 ? ty1 : ty0, txx = min(cellA + (t % 3u), 7u);
                s_cdf[i] = cl.cdf[((size_t)tyy * 8 + txx) * 256 + bin];
            }
            // bilinear-form table of both cells: pass 0 finds the largest evaluation error, pass 1 writes the entries
            float4* s_quad = reinterpret_cast<float4*>(smem + L.quad);
            double shift = 0.0;
            for (int pass = 0; pass < 2; ++pass) {
                float emax = 0.f;
                for (uint32_t i = tid; i < 2 * hm::kQuadEntries; i += NT) {
                    const uint32_t cs = i / hm::kQuadEntries, bin = i % hm::kQuadEntries;
                    const uint32_t pcx = cs ? cellB : cellA;
                    float4 qv = make_float4(0.5f, 0.f, 0.f, 0.f); // invalid pixel: sample 0 (autoscale.rs:604)
                    if (bin != 256) {
                        const uint32_t p1 = pcx + 1 < 8 ? pcx + 1 : 7;
                        const double c00 = cl.cdf[((size_t)ty0 * 8 + pcx) * 256 + bin], c01 = cl.cdf[((size_t)ty0 * 8 + p1) * 256 + bin];
                        const double c10 = cl.cdf[((size_t)ty1 * 8 + pcx) * 256 + bin], c11 = cl.cdf[((size_t)ty1 * 8 + p1) * 256 + bin];
                        if (c00 == 0.0 && c01 == 0.0 && c10 == 0.0 && c11 == 0.0) {
                            qv = make_float4(0.5f, 0.f, 0.f, 0.f);   // 0*x + 0*y == 0 exactly
                        } else if (c00 == 1.0 && c01 == 1.0 && c10 == 1.0 && c11 == 1.0) {
                            // v_ref = fl(fl(sx*omdy) + fl(sx*dy)), sx = fl(omdx + dx): 255 wherever both sums are exactly 1.0
                            // (sat_ok); elsewhere the entry is a marker (floor 480) that the block fix-up resolves per pixel
                            qv = make_float4(sat_ok[cs] ? 255.5f : hm::kMarker + 0.5f, 0.f, 0.f, 0.f);
                        } else if (c00 == c01 && c00 == c10 && c00 == c11) {
                            // Four identical CDF values c (the clamped corner cell, or tiles that agree at this bin): the
                            // reference's c*omdx + c*dx, ... differs from c by a few ulps only, |255*v_ref - 255*c| < 3e-13 for
                            // every pixel, so the sample is floor(255*c) everywhere unless 255*c sits that close to an integer.
                            // Without this, a bin whose fp32 u = A happened to have zero fraction bits would flag every one of
                            // its pixels in the cell.
                            const double U = 255.0 * c00, fl = floor(U);
                            const bool clear = U - fl > 2e-12 && fl + 1.0 - U > 2e-12;
                            qv = clear ? make_float4((float)(fl + 0.5), 0.f, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
                        } else {
                            const double A = 255.0 * c00, B = 255.0 * (c01 - c00), C = 255.0 * (c10 - c00);
                            const double D = 255.0 * ((c11 - c10) - (c01 - c00));
                            const double err = hm_entry_error(A + 1e-3, B, C, D);
                            // |u| must stay below 512: the binade of the magic constants and the 16-bit biased floor
                            const bool in_range = fabs(A) + fabs(B) + fabs(C) + fabs(D) < 460.0; // and below the marker
                            if (pass == 0) {
                                if (in_range) emax = fmaxf(emax, (float)err * 1.0001f);
                            } else if (err <= shift && in_range) {
                                qv = make_float4((float)(A + shift), (float)B, (float)C, (float)D);
                            } else {
                                qv = make_float4(0.f, 0.f, 0.f, 0.f); // fraction bits all zero: always the exact path
                            }
                        }
                    }
                    if (pass == 1) {
#pragma unroll
                        for (int r = 0; r < 8; ++r) s_quad[(cs * hm::kQuadEntries + bin) * 8 + r] = qv;
                    }
                }
                if (pass == 0) {
                    atomicMax(&s_ctrl[2], __float_as_uint(emax));
                    __syncthreads();
                    const float em = __uint_as_float(s_ctrl[2]);
                    // F fraction bits: the guard 2^-F must exceed 2*shift, shift >= every entry's error
                    fbits = 10u;
                    for (uint32_t f = 13u; f > 10u; --f)
                        if ((double)em <= ldexp(0.45, -(int)f)) { fbits = f; break; }
                    shift = ldexp(0.45, -(int)fbits);
                    magic = (float)ldexp(1.5, 23 - (int)fbits);
                }
            }
        }
        __syncthreads();
        const uint32_t fmask = hm_keep((1u << fbits) - 1u);
        const int twA = (int)hm_keep(a.clahe.tile_w * (2u * cellA + 1u)), twB = (int)hm_keep(a.clahe.tile_w * (2u * cellA + 3u));
        const float magic_k = __uint_as_float(hm_keep(__float_as_uint(magic)));

        // exact u8 sample of pixel (r, c) (local row, column) with the reference's f64 operation order
        auto exact_px = [&](uint32_t r, uint32_t c) -> uint32_t {
            if (c >= width) return 0u; // padding column of a re-pitched raster: zero taps, and it must not enter the min / max
            const uint32_t d = src[(size_t)r * cols + c];
            const uint32_t di = min(d, hot - 1u);
            const uint32_t word = reinterpret_cast<const uint32_t*>(smem + L.lut)[(size_t)di << (lut_shift - 2)];
            if (!CLAHE) return word;
            uint32_t o = 0;
            if (d) {
                const ClaheDev& cl = a.clahe;
                const uint32_t tx = cl.col_t[c];
                const uint32_t bin = (word - (sbase + L.quad)) >> 7;
                const double* s_cdf = reinterpret_cast<const double*>(smem + L.cdf);
                const uint32_t x0 = ((tx & 7u) - cellA) * 256u + bin, x1 = (((tx >> 8) & 7u) - cellA) * 256u + bin;
                double v = clahe_blend_exact_rn(s_cdf[x0], s_cdf[x1], s_cdf[768 + x0], s_cdf[768 + x1], cl.col_dx[c], cl.col_omdx[c],
                                                cl.row_dy[r], cl.row_omdy[r]);
                v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
                o = (uint32_t)__dmul_rn(v, 255.0);
            }
            mn_e = min(mn_e, o);
            mx_e = max(mx_e, o);
            return o;
        };

        const int4* const s_nt = reinterpret_cast<const int4*>(smem + L.nt);
        // ---- 16-row groups of the piece, handed out to the warps --------------------------------------
        const uint32_t n_groups = (pc.r1 - pc.r0 + 15u) / 16u;
        for (;;) {
            uint32_t grp = 0;
            if (lane == 0) grp = atomicAdd(&s_ctrl[0], 1u);
            grp = __shfl_sync(FULL, grp, 0);
            if (grp >= n_groups) break;
            const uint32_t rbase = pc.r0 + grp * 16u;
            // rows beyond the piece repeat its last row (never stored; duplicates do not disturb the min / max)
            const uint32_t rA = min(rbase + g, pc.r1 - 1u), rB = min(rbase + g + 8u, pc.r1 - 1u);
            const uint32_t oA = hm_pin(rA * cols, lane), oB = hm_pin(rB * cols, lane); // element offsets (< 2^32 for any raster in HBM)
            const bool okA_row = rbase + g < pc.r1, okB_row = rbase + g + 8u < pc.r1;
            uint8_t* const tA = reinterpret_cast<uint8_t*>(a.temp) + (size_t)(rbase + g - a.row0) * a.ax.out_size;
            uint8_t* const tB = tA + (size_t)8 * a.ax.out_size;
            float dyA = 0.f, dyB = 0.f;
            // saturated pixels of rows A / B: byte masks (0 / 0xff) "sample is 254" and "neither 254 nor 255" for the two
            // classes of columns (fl(omdx+dx) == 1.0 / == 1 - 2^-53); only read by the marker fix-up
            uint32_t rowcls = 0; // bit rw*4 + cls*2: sample is 254; bit rw*4 + cls*2 + 1: neither 254 nor 255
            if (CLAHE) {
                dyA = (float)a.clahe.row_dy[rA];
                dyB = (float)a.clahe.row_dy[rB];
                if (fixA || fixB) {
#pragma unroll
                    for (int rw = 0; rw < 2; ++rw) {
                        const uint32_t r = rw ? rB : rA;
                        const uint32_t s0 = a.clahe.row_sat[r], s1 = a.clahe.row_sat1[r];
                        rowcls |= ((s0 == 254u ? 1u : 0u) | ((s0 != 254u && s0 != 255u) ? 2u : 0u) | (s1 == 254u ? 4u : 0u) |
                                   ((s1 != 254u && s1 != 255u) ? 8u : 0u)) << (4 * rw);
                    }
                }
            }

            int acc[hm::kSlots][4]; // sum of sample * tap: the hi-byte products are folded in (<< 8) block by block
            uint32_t sj[hm::kSlots], sfb[hm::kSlots], slb[hm::kSlots], sbo[hm::kSlots];
#pragma unroll
            for (int s = 0; s < hm::kSlots; ++s) {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[s][i] = 0;
                sj[s] = st.x + s;
                sfb[s] = 0xffffffffu; slb[s] = 0; sbo[s] = 0;
                if (sj[s] < st.y) { const int4 m = s_nt[sj[s] - st.x]; sfb[s] = m.x; slb[s] = m.y; sbo[s] = m.z; }
            }
            // columns past the raster repeat its last 8-sample vector (they carry zero taps)
            auto vec_col = [&](uint32_t cb, uint32_t h) { return min(cb * 64u + h * 32u + q * 8u, cols8); };
            uint4 d[4]; // [0] row A k-step 0 (cols 8q..8q+7 of the block), [1] row A k-step 1 (cols 32+8q..), [2], [3]: row B
            {
                const uint32_t c0 = vec_col(st.z, 0), c1 = vec_col(st.z, 1);
                d[0] = hm_ld_dn(src + (oA + c0)); d[1] = hm_ld_dn(src + (oA + c1));
                d[2] = hm_ld_dn(src + (oB + c0)); d[3] = hm_ld_dn(src + (oB + c1));
            }
            for (uint32_t cb = st.z; cb < st.w; ++cb) {
                const bool more = cb + 1 < st.w;
                uint32_t w[4][2]; // packed samples: [vector][px 0..3 / 4..7]
                uint32_t riskmask = 0;
                // block entirely in cell A / entirely in cell B / holds the boundary (per-pixel select)
                const uint32_t tag = !CLAHE ? 0u : ((cb * 64u + 64u <= bcol) ? 0u : (cb * 64u >= bcol ? 1u : 2u));
                const bool fix = CLAHE && (tag == 0 ? fixA : (tag == 1 ? fixB : (fixA || fixB)));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t c0 = vec_col(cb, h);
                    const uint32_t cn = vec_col(cb + 1, h);
                    float dx[8];
                    uint32_t cm0 = 0, cm1 = 0; // columns of the vector with fl(omdx+dx) == 1.0 / == 1 - 2^-53 (bit per column)
                    if (CLAHE) {
                        // dx = m / (2*tile_w), m = 2c - tile_w*(2t+1) (k_clahe_axis), t = the cell of the vector's first column;
                        // fp32: |dxf - dx| < 2.4e-7
                        const int m0 = 2 * (int)c0 - (c0 >= bcol ? twB : twA);
                        const float dx0 = __fmul_rn((float)m0, inv2tw);
#pragma unroll
                        for (int k = 0; k < 8; ++k) dx[k] = __fmaf_rn((float)k, dstep, dx0);
                        if (fix) {
                            const uint32_t cmw = hm_lds_u16(cm_base + ((cb - st.z) * 8u + h * 4u) * 2u); // (beyond the raster: the clamped vector's)
                            cm0 = cmw & 255u;
                            cm1 = cmw >> 8;
                        }
                    }
                    if (!CLAHE) {
#pragma unroll
                        for (int rw = 0; rw < 2; ++rw) {
                            const int v = rw * 2 + h;
                            const uint32_t wv[4] = {d[v].x, d[v].y, d[v].z, d[v].w};
                            uint32_t a2[4], pr[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) a2[j] = hm_mad(__vminu2(wv[j], cap2), lut_mul, cj);
                            if (more) d[v] = hm_ld_dn(src + ((rw ? oB : oA) + cn)); // the DNs are consumed: prefetch in place
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t e0 = hm_lds_u32(a2[j] & 0xffffu), e1 = hm_lds_u32(a2[j] >> 16);
                                pr[j] = __byte_perm(e0, e1, 0x5410);
                            }
                            w[v][0] = __byte_perm(pr[0], pr[1], 0x6420);
                            w[v][1] = __byte_perm(pr[2], pr[3], 0x6420);
                        }
                    } else {
                        // both rows of the k-step together: 16 table words first, then the bin entries in groups of four
                        // (the shared loads are issued in source order, so the order below is the software pipeline)
                        uint32_t e[2][8];
#pragma unroll
                        for (int rw = 0; rw < 2; ++rw) {
                            const int v = rw * 2 + h;
                            const uint32_t wv[4] = {d[v].x, d[v].y, d[v].z, d[v].w};
                            uint32_t a2[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t vm = __vminu2(wv[j], cap2);
                                a2[j] = hm_mad(vm, lut_mul, cj);
                            }
                            if (more) d[v] = hm_ld_dn(src + ((rw ? oB : oA) + cn)); // the DNs are consumed: prefetch in place
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                e[rw][2 * j] = hm_lds_u32(a2[j] & 0xffffu);
                                e[rw][2 * j + 1] = hm_lds_u32(a2[j] >> 16);
                            }
                        }
                        uint32_t celloff = tag == 1 ? hm::kQuadCellBytes : 0u; // warp-uniform
                        if (tag == 2) { // the pixel's own cell: cell B columns sit one tile further (dx - 1)
                            const bool vb = c0 >= bcol; // dx[] was built for the vector's first column
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                if (c0 + k >= bcol) {
                                    e[0][k] += hm::kQuadCellBytes;
                                    e[1][k] += hm::kQuadCellBytes;
                                    if (!vb) dx[k] = __fsub_rn(dx[k], 1.0f);
                                }
                            }
                        }
                        uint32_t prr[2][4], racc[2] = {0, 0};
                        float4 qa[4], qb[4];
                        auto load4 = [&](float4 (&qq)[4], int rw, int k0) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) qq[i] = hm_lds_f4u(e[rw][k0 + i], celloff);
                        };
                        auto comp4 = [&](const float4 (&qq)[4], int rw, int k0) {
                            const float dy = rw ? dyB : dyA;
                            uint32_t lo[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float x = dx[k0 + i];
                                const float u = __fmaf_rn(__fmaf_rn(qq[i].w, x, qq[i].z), dy, __fmaf_rn(qq[i].y, x, qq[i].x));
                                const uint32_t m = __float_as_uint(__fadd_rd(u, magic_k));
                                racc[rw] |= (m - 1u) ^ m; // bit F set <=> the F fraction bits are all zero
                                lo[i] = __float_as_uint(__fadd_rd(u, hm::kBigC));
                            }
                            prr[rw][k0 / 2] = __byte_perm(lo[0], lo[1], 0x5410);
                            prr[rw][k0 / 2 + 1] = __byte_perm(lo[2], lo[3], 0x5410);
                        };
                        load4(qa, 0, 0);
                        load4(qb, 0, 4);
                        comp4(qa, 0, 0);
                        load4(qa, 1, 0);
                        comp4(qb, 0, 4);
                        load4(qb, 1, 4);
                        comp4(qa, 1, 0);
                        comp4(qb, 1, 4);
#pragma unroll
                        for (int rw = 0; rw < 2; ++rw) {
                            const int v = rw * 2 + h;
                            const uint32_t (&pr)[4] = prr[rw];
                            // (the one vector per row that holds padding columns of a re-pitched raster: their position-dependent
                            // samples exist in no real pixel, so the vector takes the exact path, which skips them)
                            bool risky = racc[rw] > fmask || c0 + 8u > width;
                            const uint32_t k0 = __viaddmin_s16x2_relu(pr[0], relu_c, 0x00FF00FFu);
                            const uint32_t k1 = __viaddmin_s16x2_relu(pr[1], relu_c, 0x00FF00FFu);
                            const uint32_t k2 = __viaddmin_s16x2_relu(pr[2], relu_c, 0x00FF00FFu);
                            const uint32_t k3 = __viaddmin_s16x2_relu(pr[3], relu_c, 0x00FF00FFu);
                            w[v][0] = __byte_perm(k0, k1, 0x6420);
                            w[v][1] = __byte_perm(k2, k3, 0x6420);
                            bool marked = false;
                            // marker pixels (saturated entries of a cell without the closed form; clamped to 255 above): 254
                            // where the row / column class says so (k_clahe_axis), the exact path for classes it does not cover
                            if (fix && (__vimax3_u16x2(__vimax3_u16x2(pr[0], pr[1], pr[2]), pr[3], hm::kMarkerLess2) != hm::kMarkerLess2)) {
                                uint32_t mk = 0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const uint32_t dd = __vminu2(pr[j], hm::kMarkerLess2) ^ pr[j];
                                    mk |= (((dd & 0xffffu) ? 1u : 0u) | ((dd >> 16) ? 2u : 0u)) << (2 * j);
                                }
                                marked = true;
                                const uint32_t rc = rowcls >> (4 * rw);
                                const uint32_t n254 = ((rc & 1u) ? cm0 : 0u) | ((rc & 4u) ? cm1 : 0u);
                                const uint32_t odd = (~(cm0 | cm1) & 0xffu) | ((rc & 2u) ? cm0 : 0u) | ((rc & 8u) ? cm1 : 0u);
                                if (mk & odd) risky = true;
                                const uint32_t fm = mk & n254;
                                w[v][0] -= ((fm & 15u) * 0x00204081u) & 0x01010101u;
                                w[v][1] -= ((fm >> 4) * 0x00204081u) & 0x01010101u;
                            }
                            if (risky) {
                                riskmask |= 1u << v;
                            } else if (!marked) {
                                mn2 = __vimin3_u16x2(mn2, __vimin3_u16x2(pr[0], pr[1], pr[2]), pr[3]);
                                mx2 = __vimax3_u16x2(mx2, __vimax3_u16x2(pr[0], pr[1], pr[2]), pr[3]);
                            } else { // samples as stored
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const uint32_t o = (w[v][k >> 2] >> (8 * (k & 3))) & 255u;
                                    mn_e = min(mn_e, o);
                                    mx_e = max(mx_e, o);
                                }
                            }
                        }
                    }
                }
                // tap fragments of the k-steps of this block that lie inside the windows of the n-tiles in flight (an
                // exhausted slot has sfb = ~0); in flight while the fix-up below runs
                const uint32_t ks0 = cb * 2u;
                uint4 bf[hm::kSlots][2];
                bool a0[hm::kSlots], a1[hm::kSlots];
#pragma unroll
                for (int s = 0; s < hm::kSlots; ++s) {
                    a0[s] = ks0 >= sfb[s] && ks0 <= slb[s];
                    a1[s] = ks0 + 1u >= sfb[s] && ks0 + 1u <= slb[s];
                    if (a0[s]) bf[s][0] = hm_lds_u4(sb_lane + sbo[s] + ks0 * 512u);
                    if (a1[s]) bf[s][1] = hm_lds_u4(sb_lane + sbo[s] + ks0 * 512u + 512u);
                }
                if (CLAHE) {
                    // exact fix-up: flagged vectors (8 pixels of one lane) are enumerated over the warp, ordered by (vector, lane),
                    // and recomputed four at a time, one pixel per lane; each group of 8 lanes serves one flagged vector
                    uint32_t bal[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) bal[v] = __ballot_sync(FULL, (riskmask >> v) & 1u);
                    if (bal[0] | bal[1] | bal[2] | bal[3]) {
                        const uint32_t c0n = __popc(bal[0]), c1n = c0n + __popc(bal[1]), c2n = c1n + __popc(bal[2]), total = c2n + __popc(bal[3]);
                        for (uint32_t base = 0; base < total; base += 4u) {
                            const uint32_t idx = base + (lane >> 3);
                            const bool on = idx < total;
                            // vector number and rank of the flagged lane among the lanes flagged for that vector
                            const uint32_t v = !on ? 0u : (idx < c0n ? 0u : (idx < c1n ? 1u : (idx < c2n ? 2u : 3u)));
                            const uint32_t n = idx - (v == 0 ? 0u : (v == 1 ? c0n : (v == 2 ? c1n : c2n)));
                            const uint32_t bv = v == 0 ? bal[0] : (v == 1 ? bal[1] : (v == 2 ? bal[2] : bal[3]));
                            const uint32_t sl = on ? __fns(bv, 0, n + 1) : 0u; // the flagged lane (n-th set bit)
                            const uint32_t sg = sl >> 2, sq = sl & 3u, k = lane & 7u;
                            const uint32_t r = min(rbase + sg + ((v & 2u) ? 8u : 0u), pc.r1 - 1u);
                            const uint32_t c = min(cb * 64u + (v & 1u) * 32u + sq * 8u, cols - 8u) + k;
                            uint32_t b = on ? exact_px(r, c) << (8u * (lane & 3u)) : 0u;
                            b |= __shfl_xor_sync(FULL, b, 1);
                            b |= __shfl_xor_sync(FULL, b, 2); // lanes 8s..8s+3 hold samples 0..3, lanes 8s+4..8s+7 samples 4..7
#pragma unroll
                            for (int sidx = 0; sidx < 4; ++sidx) { // hand slot sidx's two words to its flagged lane
                                const uint32_t w0 = __shfl_sync(FULL, b, 8 * sidx), w1 = __shfl_sync(FULL, b, 8 * sidx + 4);
                                const uint32_t tl = __shfl_sync(FULL, on ? (sl | (v << 8)) : 0xffffu, 8 * sidx);
                                if ((tl & 0xffu) == lane && tl != 0xffffu) {
                                    const uint32_t tv = tl >> 8;
                                    if (tv == 0) { w[0][0] = w0; w[0][1] = w1; }
                                    else if (tv == 1) { w[1][0] = w0; w[1][1] = w1; }
                                    else if (tv == 2) { w[2][0] = w0; w[2][1] = w1; }
                                    else { w[3][0] = w0; w[3][1] = w1; }
                                }
                            }
                        }
                    }
                }
                // ---- the taps: two k-steps per block ------------------------------------------------------
                const uint32_t ka[4] = {w[0][0], w[2][0], w[0][1], w[2][1]};
                const uint32_t kb[4] = {w[1][0], w[3][0], w[1][1], w[3][1]};
#pragma unroll
                for (int s = 0; s < hm::kSlots; ++s) {
                    if (a0[s] || a1[s]) {
                        int th[4] = {0, 0, 0, 0};
                        if (a0[s]) {
                            mma_u8s8(th, ka, bf[s][0].x, bf[s][0].y);
                            mma_u8u8(acc[s], ka, bf[s][0].z, bf[s][0].w);
                        }
                        if (a1[s]) {
                            mma_u8s8(th, kb, bf[s][1].x, bf[s][1].y);
                            mma_u8u8(acc[s], kb, bf[s][1].z, bf[s][1].w);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[s][i] += th[i] << 8;
                        if (ks0 + 1u >= slb[s]) { // the n-tile is complete: scale, clamp, store; the slot takes the next n-tile
                            const uint32_t ox = sj[s] * 8u + q * 2u;
                            if (wide_store) { // (warp-uniform: the shuffles below are executed by all lanes)
                                uint32_t x = 0; // bytes: row g col ox, row g col ox+1, row g+8 col ox, row g+8 col ox+1
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    int vv = (acc0 + acc[s][i]) >> prec;
                                    vv = vv < 0 ? 0 : (vv > 255 ? 255 : vv);
                                    x |= (uint32_t)vv << (8 * i);
                                    acc[s][i] = 0;
                                }
                                const uint32_t y1 = __shfl_down_sync(FULL, x, 1), y2 = __shfl_down_sync(FULL, x, 2), y3 = __shfl_down_sync(FULL, x, 3);
                                if (q == 0) { // columns sj*8 .. sj*8+7 of both rows (out_size is a multiple of 8: the tile is whole)
                                    if (okA_row) *reinterpret_cast<uint2*>(tA + ox) = make_uint2(__byte_perm(x, y1, 0x5410), __byte_perm(y2, y3, 0x5410));
                                    if (okB_row) *reinterpret_cast<uint2*>(tB + ox) = make_uint2(__byte_perm(x, y1, 0x7632), __byte_perm(y2, y3, 0x7632));
                                }
                            } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                int vv = (acc0 + acc[s][i]) >> prec;
                                vv = vv < 0 ? 0 : (vv > 255 ? 255 : vv);
                                if (((i & 2) ? okB_row : okA_row) && ox + (i & 1) < a.ax.out_size) ((i & 2) ? tB : tA)[ox + (i & 1)] = (uint8_t)vv;
                                acc[s][i] = 0;
                            }
                            }
                            sj[s] += hm::kSlots;
                            sfb[s] = 0xffffffffu;
                            if (sj[s] < st.y) { const int4 m = s_nt[sj[s] - st.x]; sfb[s] = m.x; slb[s] = m.y; sbo[s] = m.z; }
                        }
                    }
                }
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
// n-tiles (8 output columns), their 64-column source blocks, the permuted tap bytes, and strips of n-tiles whose
// source span is at most max_span columns (CLAHE: one tile width, so that a strip meets at most one cell boundary).
bool hmma_build_plan(const uint32_t* start_h, const uint32_t* size_h, const int32_t* coef_h, uint32_t window, uint32_t out_size,
                     uint32_t in_size, uint32_t max_span, HMmaPlanHost* plan, uint32_t max_strip_ntiles) {
    plan->btab.clear();
    plan->ntile.clear();
    plan->strips.clear();
    plan->weights.clear();
    plan->b_bytes = 0;
    if (out_size == 0 || in_size < 8 || (in_size % 8) != 0 || in_size >= (1u << 22)) return false; // (2c - tile_w*(2t+1) exact in fp32)
    const uint32_t n_nt = (out_size + 7) / 8;
    uint32_t koff = 0;
    for (uint32_t j = 0; j < n_nt; ++j) {
        uint32_t ws = 0xffffffffu, we = 0;
        for (uint32_t ox = j * 8; ox < std::min(out_size, j * 8 + 8); ++ox) {
            if (size_h[ox] == 0) continue;
            ws = std::min(ws, start_h[ox]);
            we = std::max(we, start_h[ox] + size_h[ox]);
        }
        if (we == 0) { ws = 0; we = 1; }
        const uint32_t fk = ws / 32, lk = (we - 1) / 32;
        if (j && (fk < (uint32_t)plan->ntile.back().x || lk < (uint32_t)plan->ntile.back().y)) return false; // windows advance
        plan->ntile.push_back(make_int4((int)fk, (int)lk, (int)koff, 0));
        for (uint32_t ks = fk; ks <= lk; ++ks)
            for (uint32_t lane = 0; lane < 32; ++lane) {
                const uint32_t n = lane >> 2, qq = lane & 3u, ox = j * 8 + n;
                uint32_t reg[4] = {0, 0, 0, 0}; // hi r0, hi r1, lo r0, lo r1
                for (uint32_t r = 0; r < 2; ++r)
                    for (uint32_t i = 0; i < 4; ++i) {
                        // k-step ks holds columns 32*ks + 8*q + 4*r + i (the order the samples are packed in)
                        const uint32_t c = ks * 32 + qq * 8 + r * 4 + i;
                        int32_t tap = 0;
                        if (ox < out_size && c >= start_h[ox] && c < start_h[ox] + size_h[ox])
                            tap = coef_h[(size_t)ox * window + (c - start_h[ox])];
                        if (tap < -32768 || tap > 32767) return false;
                        const uint32_t lo = (uint32_t)tap & 255u, hi = (uint32_t)(tap >> 8) & 255u;
                        reg[r] |= hi << (8 * i);
                        reg[2 + r] |= lo << (8 * i);
                    }
                plan->btab.push_back(make_uint4(reg[0], reg[1], reg[2], reg[3]));
            }
        koff += lk - fk + 1;
    }
    // strips
    uint32_t j0 = 0;
    while (j0 < n_nt) {
        uint32_t j1 = j0 + 1;
        auto span = [&](uint32_t e) { return ((uint32_t)plan->ntile[e - 1].y / 2 + 1 - (uint32_t)plan->ntile[j0].x / 2) * 64u; };
        auto bbytes = [&](uint32_t e) {
            return ((uint32_t)plan->ntile[e - 1].z + (uint32_t)(plan->ntile[e - 1].y - plan->ntile[e - 1].x + 1) - (uint32_t)plan->ntile[j0].z) * 512u;
        };
        if ((max_span && span(j1) > max_span) || bbytes(j1) > hm::kMaxStripB || span(j1) > hm::kMaxStripVecs * 8u) return false;
        while (j1 < n_nt && j1 - j0 < std::min(32u, std::max(1u, max_strip_ntiles)) && (!max_span || span(j1 + 1) <= max_span) && bbytes(j1 + 1) <= hm::kMaxStripB &&
               span(j1 + 1) <= hm::kMaxStripVecs * 8u)
            ++j1;
        const uint32_t cb0 = (uint32_t)plan->ntile[j0].x / 2, cb1 = (uint32_t)plan->ntile[j1 - 1].y / 2 + 1;
        // at most kSlots n-tiles of the strip meet any block, so n-tile j + kSlots starts after n-tile j has ended
        for (uint32_t cb = cb0; cb < cb1; ++cb) {
            uint32_t n = 0;
            for (uint32_t j = j0; j < j1; ++j)
                if ((uint32_t)plan->ntile[j].x / 2 <= cb && cb <= (uint32_t)plan->ntile[j].y / 2) ++n;
            if (n > (uint32_t)hm::kSlots) return false;
        }
        plan->strips.push_back(make_uint4(j0, j1, cb0, cb1));
        plan->weights.push_back(HStrip{cb0 * 64, (cb1 - cb0) * 8});
        plan->b_bytes = std::max(plan->b_bytes, bbytes(j1));
        j0 = j1;
    }
    return true;
}

// Cuts the (strip, row) space into contiguous equal-weight runs, one per CTA. `cuts` are the row positions where
// a piece must end (vertical CLAHE cell boundaries; first = 0, last = rows). A piece's weight is rows x nvec.
// Pieces are multiples of `unit` rows from their segment start, so that every warp of a CTA gets a whole 16-row group per round.
void hmma_build_pieces(const std::vector<HStrip>& strips, const std::vector<uint64_t>& cuts, uint32_t n_ctas, uint32_t unit,
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


// Host replay of the kernel's walk for one row of u8 samples (test hook): strips, slot rotation, k-step windows, the
// permuted tap bytes (hi * 256 + lo) and the final shift / clamp, in the order the device follows. out has out_size bytes.
bool hmma_replay_row(const HMmaPlanHost& plan, const uint8_t* samples, uint32_t in_size, uint32_t out_size, int precision, uint8_t* out) {
    const int acc0 = precision > 0 ? (1 << (precision - 1)) : 0;
    for (const uint4& st : plan.strips) {
        const uint32_t koff0 = (uint32_t)plan.ntile[st.x].z;
        int acc[hm::kSlots][8];
        uint32_t sj[hm::kSlots], sfb[hm::kSlots], slb[hm::kSlots], sbo[hm::kSlots];
        auto load_slot = [&](int s) {
            sfb[s] = 0xffffffffu; slb[s] = 0; sbo[s] = 0;
            if (sj[s] < st.y) { const int4 m = plan.ntile[sj[s]]; sfb[s] = (uint32_t)m.x; slb[s] = (uint32_t)m.y; sbo[s] = (uint32_t)m.z - koff0 - (uint32_t)m.x; }
        };
        for (int s = 0; s < hm::kSlots; ++s) {
            for (int i = 0; i < 8; ++i) acc[s][i] = 0;
            sj[s] = st.x + s;
            load_slot(s);
        }
        for (uint32_t cb = st.z; cb < st.w; ++cb)
            for (int s = 0; s < hm::kSlots; ++s) {
                bool any = false;
                for (uint32_t h = 0; h < 2; ++h) {
                    const uint32_t ks = cb * 2 + h;
                    if (!(ks >= sfb[s] && ks <= slb[s])) continue;
                    any = true;
                    const uint4* b = plan.btab.data() + (size_t)(koff0 + sbo[s] + ks) * 32u; // fragments of this k-step, 32 lanes
                    for (uint32_t lane = 0; lane < 32; ++lane) {
                        const uint32_t n = lane >> 2, q = lane & 3u;
                        const uint32_t reg[4] = {b[lane].x, b[lane].y, b[lane].z, b[lane].w};
                        for (uint32_t r = 0; r < 2; ++r)
                            for (uint32_t i = 0; i < 4; ++i) {
                                const uint32_t c = std::min(ks * 32 + q * 8, in_size - 8) + r * 4 + i; // the column the kernel loads
                                const int hi = (int)(int8_t)((reg[r] >> (8 * i)) & 255u), lo = (int)((reg[2 + r] >> (8 * i)) & 255u);
                                acc[s][n] += (int)samples[c] * (hi * 256 + lo);
                            }
                    }
                }
                if (any && cb * 2 + 1 >= slb[s]) {
                    for (uint32_t n = 0; n < 8; ++n) {
                        const uint32_t ox = sj[s] * 8 + n;
                        int v = (acc0 + acc[s][n]) >> precision;
                        v = v < 0 ? 0 : (v > 255 ? 255 : v);
                        if (ox < out_size) out[ox] = (uint8_t)v;
                        acc[s][n] = 0;
                    }
                    sj[s] += hm::kSlots;
                    load_slot(s);
                }
            }
    }
    return true;
}

uint32_t hmma_warps(bool clahe) { return (clahe ? hm::kThreadsClahe : hm::kThreadsLut) / 32u; }

// Shared memory for the largest table a plan may ask for (the launch does not know `hot`: it is in device memory). The three
// table shapes top out at 500 << 7 = 1000 << 6 = 2000 << 5 = 64,000 bytes, so the bound is the same for every hot.
size_t hmma_smem_bytes(int src_kind, uint32_t b_bytes) {
    return hmma_layout(src_kind == HSRC_DN_CLAHE, kHmmaMaxHot << hmma_lut_shift(kHmmaMaxHot), b_bytes).total;
}

cudaError_t launch_hmma(const HResizeArgs& a, int src_kind, const uint4* btab_dev, const int4* ntile_dev, const uint4* strips_dev,
                        const uint32_t* pieces_dev, const uint32_t* cta_first_dev, uint32_t n_ctas, uint32_t b_bytes,
                        cudaStream_t stream) {
    if (a.n_rows == 0 || a.ax.out_size == 0 || n_ctas == 0) return cudaSuccess;
    if (!a.plan) return cudaErrorInvalidValue;
    const bool clahe = src_kind == HSRC_DN_CLAHE;
    const size_t smem = hmma_smem_bytes(src_kind, b_bytes);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    HMmaParams pp;
    pp.btab = btab_dev;
    pp.ntile = ntile_dev;
    pp.strips = strips_dev;
    pp.pieces = reinterpret_cast<const HPiece*>(pieces_dev);
    pp.cta_first = cta_first_dev;
    pp.b_bytes = b_bytes;
    if (clahe) {
        if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_hmma<true>), smem)) return e;
        k_hmma<true><<<n_ctas, hm::kThreadsClahe, smem, stream>>>(a, pp);
    } else {
        if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_hmma<false>), smem)) return e;
        k_hmma<false><<<n_ctas, hm::kThreadsLut, smem, stream>>>(a, pp);
    }
    return cudaGetLastError();
}

} // namespace sarpro

// kernels_hrow.cu — pass B for u8 samples, fourth generation: the per-pixel stage (DN -> sample) fused with the horizontal
// Lanczos pass; DN tiles and tap fragments staged by TMA, the CLAHE blend folded per row, taps on the integer tensor cores.
//
// What changed against kernels_hmma.cu (profiles/r02_*): that kernel gave every warp its own 16 rows of a strip, which forces
// the CLAHE blend u = A + B*dx + (C + D*dx)*dy to fetch four coefficients per pixel (a 128-bit shared gather, 4 wavefronts per
// warp instruction) and three FFMAs, keeps 16 registers of prefetched DNs and the slot state of three n-tiles alive per thread,
// and leaves warps idle when a piece has fewer 16-row groups than the CTA has warps (a rank's band of a sharded scene).
// Here the whole CTA works on ONE 16-row group of a strip at a time:
//   * the blend is folded over dy once per group: P = A + C*dy, Q = B + D*dy for the 16 rows (f64, rounded to fp32), stored as
//     [bin][row] pairs, 128 B per bin. A lane owns one row (lane & 15), so the 16 lanes of a half-warp read 16 different bank
//     pairs whatever their bins are: a conflict-free 64-bit gather (2 wavefronts) and ONE FFMA per pixel, u = P + Q*dx;
//   * DN tiles (16 rows x 64 columns, 2 KB) arrive through a ring of TMA stages (cp.async.bulk.tensor.2d with the 128-byte
//     swizzle, mbarrier full / empty pairs, one producer warp): no global-load address arithmetic or prefetch registers in
//     the consumers, and the DN reads are conflict-free LDS.128;
//   * stage 1: warps pull 64-column blocks of the group from a shared counter (any warp takes any block) and leave the u8
//     samples in a shared tile of the group (16 rows x the strip's columns; 8-byte chunks rotated per row so that both the
//     writers and the fragment readers are conflict-free). Stage 2: a warp owns whole n-tiles (8 output columns): A fragments
//     of mma.sync.m16n8k32 straight from the sample tile, tap fragments from L2 (1.3 MB for the C3 axis, read once per group),
//     accumulators in registers, 8-byte stores. (A first version accumulated per-block partial sums with red.shared.add:
//     shared atomics on 32 different words cost 64 LSU cycles per warp instruction, 4x slower than the whole third-generation
//     kernel.) Stage 2 of a group overlaps the fold of the next group's table; two CTA barriers per group.
// Exactness is argued exactly as in kernels_hmma.cu: the fp32 value carries a shift S >= its error bound (now smaller: dy
// enters in f64 at fold time), a truncation-bit test flags the 2^-F of the pixels whose floor could differ, and flagged
// 8-pixel vectors are recomputed with the reference's f64 operation order (autoscale.rs:320-329). Integer taps: bit-exact.
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "clahe_exact.cuh"
#include "common.cuh"
#include "kernels.h"

namespace sarpro {

namespace hr {
constexpr uint32_t kConsumerWarps = 12;
constexpr uint32_t kConsumers = kConsumerWarps * 32;
constexpr uint32_t kThreads = kConsumers + 32; // + the producer warp
constexpr uint32_t kBins = 257;                // 256 bins + the invalid-pixel entry
constexpr uint32_t kCellBytes = kBins * 128;   // [bin][16 rows] float2
constexpr uint32_t kMaxNt = 36;                // n-tiles per strip (three per consumer warp)
constexpr uint32_t kStages = 12;               // ring of 2 KB DN tiles: one per consumer warp, so no warp waits on another's stage
constexpr uint32_t kMaxBlocks = 50;            // 64-column blocks per strip (the sample tile holds 16 rows x 3200 columns)
constexpr float kBigC = 12582912.0f + 512.0f;  // u + kBigC (RD): low 16 bits = floor(u) + 512
constexpr float kMarker = 480.0f;
constexpr uint32_t kMarkerLess2 = (480u + 512u - 1u) * 0x10001u;
} // namespace hr

// ---- small PTX helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hr_lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t hr_lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 hr_lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
// float2 at shared address a + off; when off is warp-uniform the add folds into the LDS operand (R + UR)
__device__ __forceinline__ float2 hr_lds_f2o(uint32_t a, uint32_t off) {
    float2 v;
    asm volatile("{ .reg .u32 t; add.u32 t, %2, %3; ld.shared.v2.f32 {%0,%1}, [t]; }" : "=f"(v.x), "=f"(v.y) : "r"(a), "r"(off));
    return v;
}
__device__ __forceinline__ uint32_t hr_mad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint2 hr_lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 hr_lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void hr_sts_u2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void hr_red_add(uint32_t addr, int v) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void hr_bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(hr::kConsumers) : "memory"); }
__device__ __forceinline__ void hr_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void hr_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hr_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void hr_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra HR_DONE_%=;\n"
        "bra HR_WAIT_%=;\n"
        "HR_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void hr_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void hr_bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void hr_mma_u8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void hr_mma_u8u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- shared-memory layout ------------------------------------------------------------------------------------------------
struct HRowSmem {
    uint32_t ring, lut, pq, cdf, smp, cm, ctrl, bars, total;
};
__host__ __device__ inline uint32_t hrow_lut_shift(uint32_t hot) { return hot <= 500 ? 7u : (hot <= 1000 ? 6u : 5u); }
__host__ __device__ inline HRowSmem hrow_layout(bool clahe, uint32_t lut_bytes) {
    HRowSmem L;
    uint32_t o = 0;
    L.lut = o;  // first: the packed address arithmetic of stage 1 needs the table below 64 KB of the shared window
    o += (lut_bytes + 1023) & ~1023u;
    L.ring = o; // 1024-byte aligned 2 KB stages (128-byte swizzle atoms)
    o += hr::kStages * 2048u;
    L.smp = o;  // samples of the group: 16 rows x kMaxBlocks * 64 bytes, 8-byte chunks rotated per row inside 128-byte lines
    o += 16u * hr::kMaxBlocks * 64u;
    L.pq = o;
    if (clahe) o += 2 * hr::kCellBytes;
    L.cdf = o;  // f64 {A, B, C, D} of both cells (the exact path reads the CDFs from global memory: 2^-F of the pixels)
    if (clahe) o += 2 * hr::kBins * 32;
    L.cm = o;
    if (clahe) o += hr::kMaxBlocks * 8u * 2u;
    L.ctrl = o;
    o += 64;
    L.bars = o;
    o += 2 * 16 * 8 + 16 * 4; // full / empty mbarriers, then the block number each stage was last armed for
    L.total = o;
    return L;
}

struct HRowParams {
    const uint4* btab;       // tap fragments: [koff + (k-step - first k-step)][lane] = {hi r0, hi r1, lo r0, lo r1}
    const int4* ntile;       // per n-tile: {first k-step, last k-step, koff, 0}; a k-step is 32 source columns
    const uint4* strips;     // per strip: {first n-tile, end n-tile, first block, end block}; a block is 64 source columns
    const HPiece* pieces;
    const uint32_t* cta_first;
};

__device__ __forceinline__ double hr_half_ulp(double x) {
    const double ax = fabs(x);
    if (ax < 1e-30) return 0.0;
    int e;
    frexp(ax, &e);
    return ldexp(1.0, e - 25); // half an ulp of the fp32 binade of ax
}
// Error bound of u = fl32(fl32(Q) * dxf + fl32(P)), P = A + C*dy, Q = B + D*dy (f64, |dy| <= 1), against the reference's f64 value.
__device__ __forceinline__ double hr_entry_error(double A, double B, double C, double D) {
    const double X = 1.0;          // |dx| <= 1 (autoscale.rs:308-318: d in [-0.5, 1))
    const double ex = 2.4e-7;      // |dxf - dx|
    const double aP = fabs(A) + fabs(C), aQ = fabs(B) + fabs(D);
    double e = hr_half_ulp(aP) + hr_half_ulp(aQ) * X; // P, Q rounded to fp32
    e += aQ * ex;                                      // geometry
    e += hr_half_ulp(aP + aQ * X);                     // the FFMA
    return e * 1.05 + 1e-9;                            // second-order terms; the reference's own f64 roundings (< 1e-12)
}

// 8-byte chunk c8 (64 per KB... 16 per 128-byte line) of row r of the sample tile: lines of 128 bytes, the chunk rotated by
// f(r) = 4*(r&3) + (r>>2) inside its line. Writers (16 lanes = 16 rows, one chunk) and readers (4 rows x 4 consecutive chunks)
// both touch 16 different bank pairs.
__device__ __forceinline__ uint32_t hr_smp_addr(uint32_t base, uint32_t pitch, uint32_t r, uint32_t fr, uint32_t c8) {
    return base + r * pitch + (c8 >> 4) * 128u + (((c8 + fr) & 15u) << 3);
}

template <bool CLAHE>
__global__ void __launch_bounds__(hr::kThreads, 1) k_hrow(HResizeArgs a, HRowParams pp, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr uint32_t FULL = 0xffffffffu;
    if (a.skip && *a.skip) return;
    const uint32_t hot = a.plan->hot, hot_top = a.plan->hot_top;
    if (hot == 0 || a.plan->use_generic) return; // the generic exact kernel (queued behind this one) takes the band
    // the TMA stages need 1024-byte alignment (128-byte swizzle atoms); the launch reserves the slack
    unsigned char* const smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t lut_shift = hrow_lut_shift(hot);
    const HRowSmem L = hrow_layout(CLAHE, hot << lut_shift);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t cols = a.src_cols;
    uint32_t* const s_ctrl = reinterpret_cast<uint32_t*>(smem + L.ctrl);
    constexpr uint32_t n_stages = hr::kStages;
    const uint32_t bar_full = sbase + L.bars, bar_empty = sbase + L.bars + 16 * 8;
    // Block number a stage was last armed for: a consumer first waits until the producer has armed the stage for ITS block,
    // then on the barrier (a parity wait alone cannot tell this fill from the one a whole trip round the ring earlier).
    volatile uint32_t* const s_issue = reinterpret_cast<volatile uint32_t*>(smem + L.bars + 2 * 16 * 8);

    if (tid == 0) {
        for (uint32_t s = 0; s < n_stages; ++s) {
            hr_mbar_init(bar_full + 8 * s, 1);
            hr_mbar_init(bar_empty + 8 * s, 1);
            s_issue[s] = 0xffffffffu;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // =====================================================================================================================
    // producer warp: walks the same (piece, group, block) sequence as the consumers and keeps the ring of DN tiles full
    // =====================================================================================================================
    if (wid == hr::kConsumerWarps) {
        if (lane == 0) {
            uint32_t seq = 0;
            for (uint32_t pi = pp.cta_first[blockIdx.x]; pi < pp.cta_first[blockIdx.x + 1]; ++pi) {
                const HPiece pc = pp.pieces[pi];
                const uint4 st = pp.strips[pc.strip];
                const uint32_t n_groups = (pc.r1 - pc.r0 + 15u) / 16u;
                for (uint32_t grp = 0; grp < n_groups; ++grp) {
                    const int row = (int)(pc.r0 + grp * 16u); // local row of the raster the tensor map describes
                    for (uint32_t cb = st.z; cb < st.w; ++cb, ++seq) {
                        const uint32_t stg = seq % n_stages, round = seq / n_stages;
                        if (round) hr_mbar_wait(bar_empty + 8 * stg, (round - 1u) & 1u);
                        hr_mbar_expect_tx(bar_full + 8 * stg, 2048u);
                        hr_tma_2d(sbase + L.ring + stg * 2048u, &tmap, bar_full + 8 * stg, (int)(cb * 64u), row);
                        s_issue[stg] = seq;
                    }
                }
            }
        }
        return;
    }

    // =====================================================================================================================
    // consumers
    // =====================================================================================================================
    const uint32_t rr = lane & 15u, hh = lane >> 4;   // stage 1: the lane's row of the group and its 32-column half of a block
    const uint32_t g = lane >> 2, q = lane & 3u;      // stage 2: fragment coordinates
    {   // once per CTA: DN -> table word, R lane-interleaved replicas (replica = lane % R; R >= 16 carries the lane's row offset)
        uint4* s_lut4 = reinterpret_cast<uint4*>(smem + L.lut);
        const uint32_t per = 1u << (lut_shift - 4); // uint4 per table entry
        for (uint32_t i = tid; i < hot * per; i += hr::kConsumers) {
            const uint32_t idx = i >> (lut_shift - 4), r0 = (i & (per - 1)) * 4u;
            const uint32_t e = idx + 1 == hot ? hot_top : (a.lut[idx] & 255u);
            uint32_t v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (CLAHE) {
                    const uint32_t bin = idx ? e : 256u; // DN 0 is the only invalid DN (pipeline.rs:22)
                    v[j] = sbase + L.pq + bin * 128u + ((r0 + j) & 15u) * 8u;
                } else {
                    v[j] = e;
                }
            }
            s_lut4[i] = make_uint4(v[0], v[1], v[2], v[3]);
        }
    }
    const uint32_t n_rep = 1u << (lut_shift - 2);                          // replicas of the table
    const uint32_t lut_lane = sbase + L.lut + (lane & (n_rep - 1u)) * 4u;
    const uint32_t row_fix = (CLAHE && n_rep < 16u) ? (rr & 8u) * 8u : 0u; // 8 replicas only carry (row & 7)
    const uint32_t lut_mul = 1u << lut_shift;
    const int prec = a.ax.precision;
    const int acc0 = prec > 0 ? (1 << (prec - 1)) : 0;
    uint32_t mn2 = 0xffffffffu, mx2 = 0u;   // fast path: u16x2 running min / max of floor(u) + 512 (not yet clamped)
    uint32_t mn_e = 0xffffffffu, mx_e = 0;  // exact-path samples
    uint32_t staged_strip = 0xffffffffu;
    const float inv2tw = CLAHE ? a.clahe.inv2tw : 0.f;
    const float dstep = __fmul_rn(2.0f, inv2tw);
    const uint32_t f_own = 4u * (rr & 3u) + (rr >> 2);
    const uint32_t f_ga = 4u * (g & 3u) + (g >> 2), f_gb = 4u * (g & 3u) + ((g + 8u) >> 2);
    double4* const s_abcd = reinterpret_cast<double4*>(smem + L.cdf);
    if (sbase + L.lut + (hot << lut_shift) > 65536u) __trap(); // 16-bit table addresses (dynamic shared memory starts low on sm_100)
    const uint32_t cap2 = (hot - 1u) * 0x10001u;
    const uint32_t cj = lut_lane * 0x10001u;
    const bool wide_store = (a.ax.out_size & 7u) == 0 && (reinterpret_cast<uintptr_t>(a.temp) & 7u) == 0;
    uint32_t seq_base = 0; // blocks the CTA has consumed before the current group

    for (uint32_t pi = pp.cta_first[blockIdx.x]; pi < pp.cta_first[blockIdx.x + 1]; ++pi) {
        const HPiece pc = pp.pieces[pi];
        const uint4 st = pp.strips[pc.strip]; // {j0, j1, cb0, cb1}
        const uint32_t n_nt = st.y - st.x, n_blocks = st.w - st.z;
        const uint32_t pitch = ((n_blocks + 1u) & ~1u) * 64u; // bytes per row of the sample tile: whole 128-byte lines
        hr_bar_consumers(); // the previous piece is done with the tables
        if (tid == 0) { s_ctrl[1] = 0xffffffffu; s_ctrl[2] = 0; s_ctrl[3] = 7u; }
        if (CLAHE && staged_strip != pc.strip) {
            uint16_t* s_cm = reinterpret_cast<uint16_t*>(smem + L.cm);
            for (uint32_t i = tid; i < n_blocks * 8u; i += hr::kConsumers) {
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
            staged_strip = pc.strip;
        }

        // ---- per-piece tables --------------------------------------------------------------------------------
        uint32_t cellA = 0, bcol = 0xffffffffu, fbits = 13;
        float magic = 1536.0f;
        bool fixA = false, fixB = false;
        if (CLAHE) {
            const ClaheDev& cl = a.clahe;
            const uint32_t ty = cl.row_t[pc.r0]; // the piece lies inside one vertical bilinear cell
            const uint32_t ty0 = ty & 7u, ty1 = (ty >> 8) & 7u;
            const uint32_t first_col = min(st.z * 64u, cols - 1u), last_col = min(st.w * 64u, cols) - 1u;
            cellA = cl.col_t[first_col] & 7u;
            const uint32_t cellB = cl.col_t[last_col] & 7u; // == cellA or cellA + 1 (strip span <= tile width)
            hr_bar_consumers();
            // first column of cell B; saturated-bin shortcut (all four CDFs exactly 1.0 -> sample 255) is valid for a
            // cell only when fl(omdx+dx) == 1 for all of its columns in the strip and fl(omdy+dy) == 1 for all rows
            uint32_t bad = 0; // bit 0: cell A not ok, bit 1: cell B, bit 2: rows
            for (uint32_t c = first_col + tid; c <= last_col; c += hr::kConsumers) {
                const uint32_t ct = cl.col_t[c];
                if ((ct & 7u) != cellA) atomicMin(&s_ctrl[1], c);
                if (!(ct & 0x80u)) bad |= ((ct & 7u) == cellA) ? 1u : 2u;
            }
            for (uint32_t r = pc.r0 + tid; r < pc.r1; r += hr::kConsumers) if (cl.row_sat[r] != 255u) bad |= 4u;
            if (bad) atomicAnd(&s_ctrl[3], ~bad);
            hr_bar_consumers();
            bcol = s_ctrl[1];
            const uint32_t okm = s_ctrl[3];
            const bool sat_ok[2] = {(okm & 5u) == 5u, (okm & 6u) == 6u};
            fixA = !sat_ok[0];
            fixB = !sat_ok[1];
            // bilinear-form coefficients of both cells in f64: pass 0 finds the largest evaluation error, pass 1 writes them
            double shift = 0.0;
            for (int pass = 0; pass < 2; ++pass) {
                float emax = 0.f;
                for (uint32_t i = tid; i < 2 * hr::kBins; i += hr::kConsumers) {
                    const uint32_t cs = i / hr::kBins, bin = i % hr::kBins;
                    const uint32_t pcx = cs ? cellB : cellA;
                    double4 qv = make_double4(0.5, 0.0, 0.0, 0.0); // invalid pixel: sample 0 (autoscale.rs:604)
                    if (bin != 256) {
                        const uint32_t p1 = pcx + 1 < 8 ? pcx + 1 : 7;
                        const double c00 = cl.cdf[((size_t)ty0 * 8 + pcx) * 256 + bin], c01 = cl.cdf[((size_t)ty0 * 8 + p1) * 256 + bin];
                        const double c10 = cl.cdf[((size_t)ty1 * 8 + pcx) * 256 + bin], c11 = cl.cdf[((size_t)ty1 * 8 + p1) * 256 + bin];
                        if (c00 == 0.0 && c01 == 0.0 && c10 == 0.0 && c11 == 0.0) {
                            qv = make_double4(0.5, 0.0, 0.0, 0.0);   // 0*x + 0*y == 0 exactly
                        } else if (c00 == 1.0 && c01 == 1.0 && c10 == 1.0 && c11 == 1.0) {
                            // v_ref = fl(fl(sx*omdy) + fl(sx*dy)), sx = fl(omdx + dx): 255 wherever both sums are exactly 1.0
                            // (sat_ok); elsewhere the entry is a marker (floor 480) that the fix-up resolves per pixel
                            qv = make_double4(sat_ok[cs] ? 255.5 : (double)hr::kMarker + 0.5, 0.0, 0.0, 0.0);
                        } else if (c00 == c01 && c00 == c10 && c00 == c11) {
                            // four identical CDF values c: |255*v_ref - 255*c| < 3e-13 for every pixel (see kernels_hmma.cu)
                            const double U = 255.0 * c00, fl = floor(U);
                            const bool clear = U - fl > 2e-12 && fl + 1.0 - U > 2e-12;
                            qv = clear ? make_double4(fl + 0.5, 0.0, 0.0, 0.0) : make_double4(0.0, 0.0, 0.0, 0.0);
                        } else {
                            const double A = 255.0 * c00, B = 255.0 * (c01 - c00), C = 255.0 * (c10 - c00);
                            const double D = 255.0 * ((c11 - c10) - (c01 - c00));
                            const double err = hr_entry_error(A + 1e-3, B, C, D);
                            // |u| must stay below 512: the binade of the magic constants and the 16-bit biased floor
                            const bool in_range = fabs(A) + fabs(B) + fabs(C) + fabs(D) < 460.0; // and below the marker
                            if (pass == 0) {
                                if (in_range) emax = fmaxf(emax, (float)err * 1.0001f);
                            } else if (err <= shift && in_range) {
                                qv = make_double4(A + shift, B, C, D);
                            } else {
                                qv = make_double4(0.0, 0.0, 0.0, 0.0); // fraction bits all zero: always the exact path
                            }
                        }
                    }
                    if (pass == 1) s_abcd[i] = qv;
                }
                if (pass == 0) {
                    atomicMax(&s_ctrl[2], __float_as_uint(emax));
                    hr_bar_consumers();
                    const float em = __uint_as_float(s_ctrl[2]);
                    // F fraction bits: the guard 2^-F must exceed 2*shift, shift >= every entry's error
                    fbits = 10u;
                    for (uint32_t f = 13u; f > 10u; --f)
                        if ((double)em <= ldexp(0.45, -(int)f)) { fbits = f; break; }
                    shift = ldexp(0.45, -(int)fbits);
                    magic = (float)ldexp(1.5, 23 - (int)fbits);
                }
            }
            hr_bar_consumers(); // the coefficient tables are complete before the first fold reads them
        }
        const uint32_t fmask = (1u << fbits) - 1u;
        const int twA = (int)(a.clahe.tile_w * (2u * cellA + 1u)), twB = (int)(a.clahe.tile_w * (2u * cellA + 3u));

        // fold the blend over dy for the 16 rows of a group: [cell][bin][row] = {P, Q}
        auto fold = [&](uint32_t rbase) {
            if (!CLAHE) return;
            float2* s_pq = reinterpret_cast<float2*>(smem + L.pq);
            // entry i = (cell * 257 + bin) * 16 + row; the stride (384 threads) is a multiple of 16: a thread keeps its row
            const double dy = a.clahe.row_dy[min(rbase + (tid & 15u), pc.r1 - 1u)];
            for (uint32_t i = tid; i < 2 * hr::kBins * 16u; i += hr::kConsumers) {
                const double4 c = s_abcd[i >> 4];
                s_pq[i] = make_float2((float)__dadd_rn(c.x, __dmul_rn(c.z, dy)), (float)__dadd_rn(c.y, __dmul_rn(c.w, dy)));
            }
        };

        const uint32_t n_groups = (pc.r1 - pc.r0 + 15u) / 16u;
        fold(pc.r0);
        if (tid == 0) s_ctrl[0] = 0;
        hr_bar_consumers();
        for (uint32_t grp = 0; grp < n_groups; ++grp) {
            const uint32_t rbase = pc.r0 + grp * 16u;
            const uint32_t n_valid_rows = min(16u, pc.r1 - rbase);
            const bool row_ok = rr < n_valid_rows;
            const uint32_t my_row = min(rbase + rr, pc.r1 - 1u);
            uint32_t rowcls = 0; // bit cls*2: sample is 254; bit cls*2 + 1: neither 254 nor 255 (column classes 0 / 1)
            if (CLAHE && (fixA || fixB)) {
                const uint32_t s0 = a.clahe.row_sat[my_row], s1 = a.clahe.row_sat1[my_row];
                rowcls = (s0 == 254u ? 1u : 0u) | ((s0 != 254u && s0 != 255u) ? 2u : 0u) | (s1 == 254u ? 4u : 0u) |
                         ((s1 != 254u && s1 != 255u) ? 8u : 0u);
            }

            // ---- stage 1: samples of the group's blocks, pulled from the shared counter -------------------------------------
            for (;;) {
                uint32_t bi = 0;
                if (lane == 0) bi = atomicAdd(&s_ctrl[0], 1u);
                bi = __shfl_sync(FULL, bi, 0);
                if (bi >= n_blocks) break;
                const uint32_t cb = st.z + bi;
                const uint32_t seq = seq_base + bi;
                const uint32_t stg = seq % n_stages;
                const uint32_t sdn = sbase + L.ring + stg * 2048u;
                while (s_issue[stg] != seq) {}
                hr_mbar_wait(bar_full + 8 * stg, (seq / n_stages) & 1u);
                uint4 dq[4];
#pragma unroll
                for (uint32_t v = 0; v < 4; ++v) // 128-byte swizzle of the TMA box: chunk ^ (row & 7)
                    dq[v] = hr_lds_u4(sdn + rr * 128u + (((hh * 4u + v) ^ (rr & 7u)) << 4));
                __syncwarp();
                if (lane == 0) hr_mbar_arrive(bar_empty + 8 * stg); // the tile is in registers: the stage can be refilled

                // block entirely in cell A / entirely in cell B / holds the boundary (per-pixel select)
                const uint32_t tag = !CLAHE ? 0u : ((cb * 64u + 64u <= bcol) ? 0u : (cb * 64u >= bcol ? 1u : 2u));
                const bool fix = CLAHE && (tag == 0 ? fixA : (tag == 1 ? fixB : (fixA || fixB)));
#pragma unroll
                for (uint32_t v = 0; v < 4; ++v) {
                    const uint32_t chunk = hh * 4u + v;
                    const uint32_t c0 = cb * 64u + chunk * 8u; // first column of the lane's vector
                    const uint32_t dw[4] = {dq[v].x, dq[v].y, dq[v].z, dq[v].w};
                    const bool in_raster = c0 < cols; // (cols % 8 == 0: a vector is inside or outside as a whole)
                    uint32_t w0, w1;
                    if (!CLAHE) {
                        uint32_t pr[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t a2 = hr_mad(__vminu2(dw[j], cap2), lut_mul, cj);
                            pr[j] = __byte_perm(hr_lds_u32(a2 & 0xffffu), hr_lds_u32(a2 >> 16), 0x5410);
                        }
                        w0 = __byte_perm(pr[0], pr[1], 0x6420);
                        w1 = __byte_perm(pr[2], pr[3], 0x6420);
                    } else {
                        uint32_t e[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { // two DNs per word: clamp, scale and rebase both halves at once (16-bit addresses)
                            const uint32_t a2 = hr_mad(__vminu2(dw[j], cap2), lut_mul, cj);
                            e[2 * j] = hr_lds_u32(a2 & 0xffffu);
                            e[2 * j + 1] = hr_lds_u32(a2 >> 16);
                        }
                        // dx = m / (2*tile_w), m = 2c - tile_w*(2t+1) (k_clahe_axis), t = the cell of the vector's first column;
                        // fp32: |dxf - dx| < 2.4e-7
                        float dx[8];
                        const bool vb = c0 >= bcol;
                        const int m0 = 2 * (int)c0 - (vb ? twB : twA);
                        const float dx0 = __fmul_rn((float)m0, inv2tw);
#pragma unroll
                        for (int k = 0; k < 8; ++k) dx[k] = __fmaf_rn((float)k, dstep, dx0);
                        const uint32_t celloff = (tag == 1 ? hr::kCellBytes : 0u) + row_fix;
                        if (tag == 2) { // the pixel's own cell: cell B columns sit one tile further (dx - 1)
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                if (c0 + k >= bcol) {
                                    e[k] += hr::kCellBytes;
                                    if (!vb) dx[k] = __fsub_rn(dx[k], 1.0f);
                                }
                            }
                        }
                        uint32_t lo[8], racc = 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float2 pq = hr_lds_f2o(e[k], celloff);
                            const float u = __fmaf_rn(pq.y, dx[k], pq.x);
                            const uint32_t m = __float_as_uint(__fadd_rd(u, magic));
                            racc |= (m - 1u) ^ m; // bit F set <=> the F fraction bits are all zero
                            lo[k] = __float_as_uint(__fadd_rd(u, hr::kBigC));
                        }
                        uint32_t pr[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) pr[j] = __byte_perm(lo[2 * j], lo[2 * j + 1], 0x5410);
                        bool risky = racc > fmask;
                        const uint32_t relu_c = 0xFE00FE00u; // -512 per half
                        const uint32_t k0 = __viaddmin_s16x2_relu(pr[0], relu_c, 0x00FF00FFu);
                        const uint32_t k1 = __viaddmin_s16x2_relu(pr[1], relu_c, 0x00FF00FFu);
                        const uint32_t k2 = __viaddmin_s16x2_relu(pr[2], relu_c, 0x00FF00FFu);
                        const uint32_t k3 = __viaddmin_s16x2_relu(pr[3], relu_c, 0x00FF00FFu);
                        w0 = __byte_perm(k0, k1, 0x6420);
                        w1 = __byte_perm(k2, k3, 0x6420);
                        bool marked = false;
                        // marker pixels (saturated entries of a cell without the closed form; clamped to 255 above): 254 where
                        // the row / column class says so (k_clahe_axis), the exact path for classes it does not cover
                        if (fix && (__vimax3_u16x2(__vimax3_u16x2(pr[0], pr[1], pr[2]), pr[3], hr::kMarkerLess2) != hr::kMarkerLess2)) {
                            const uint32_t cmw = reinterpret_cast<const uint16_t*>(smem + L.cm)[bi * 8u + chunk];
                            const uint32_t cm0 = cmw & 255u, cm1 = cmw >> 8;
                            uint32_t mk = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t dd = __vminu2(pr[j], hr::kMarkerLess2) ^ pr[j];
                                mk |= (((dd & 0xffffu) ? 1u : 0u) | ((dd >> 16) ? 2u : 0u)) << (2 * j);
                            }
                            marked = true;
                            const uint32_t n254 = ((rowcls & 1u) ? cm0 : 0u) | ((rowcls & 4u) ? cm1 : 0u);
                            const uint32_t odd = (~(cm0 | cm1) & 0xffu) | ((rowcls & 2u) ? cm0 : 0u) | ((rowcls & 8u) ? cm1 : 0u);
                            if (mk & odd) risky = true;
                            const uint32_t fm = mk & n254;
                            w0 -= ((fm & 15u) * 0x00204081u) & 0x01010101u;
                            w1 -= ((fm >> 4) * 0x00204081u) & 0x01010101u;
                        }
                        if (risky && in_raster) {
                            // exact path: the lane recomputes its 8 pixels with the reference's f64 operation order
                            // (autoscale.rs:320-329, :602-606); 2^-F of the pixels get here
                            const ClaheDev& cl = a.clahe;
                            const double dyr = cl.row_dy[my_row], omdyr = cl.row_omdy[my_row];
                            const uint32_t tyr = cl.row_t[my_row];
                            const double* t0 = cl.cdf + (size_t)(tyr & 7u) * 8u * 256u;
                            const double* t1 = cl.cdf + (size_t)((tyr >> 8) & 7u) * 8u * 256u;
                            uint32_t o[8];
#pragma unroll 1
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t d = (dw[k >> 1] >> (16 * (k & 1))) & 0xffffu;
                                uint32_t ov = 0;
                                if (d) {
                                    const uint32_t c = c0 + k;
                                    const uint32_t word = hr_lds_u32(lut_lane + min(d, hot - 1u) * lut_mul);
                                    const uint32_t bin = (word - (sbase + L.pq)) >> 7;
                                    const uint32_t tx = cl.col_t[c];
                                    const uint32_t x0 = (tx & 7u) * 256u + bin, x1 = ((tx >> 8) & 7u) * 256u + bin;
                                    double vv = clahe_blend_exact_rn(t0[x0], t0[x1], t1[x0], t1[x1], cl.col_dx[c], cl.col_omdx[c], dyr, omdyr);
                                    vv = vv < 0.0 ? 0.0 : (vv > 1.0 ? 1.0 : vv);
                                    ov = (uint32_t)__dmul_rn(vv, 255.0);
                                }
                                o[k] = ov;
                                if (row_ok) { mn_e = min(mn_e, ov); mx_e = max(mx_e, ov); }
                            }
                            w0 = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
                            w1 = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
                        } else if (row_ok && in_raster) {
                            if (!marked) {
                                mn2 = __vimin3_u16x2(mn2, __vimin3_u16x2(pr[0], pr[1], pr[2]), pr[3]);
                                mx2 = __vimax3_u16x2(mx2, __vimax3_u16x2(pr[0], pr[1], pr[2]), pr[3]);
                            } else { // samples as stored
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const uint32_t ov = ((k < 4 ? w0 : w1) >> (8 * (k & 3))) & 255u;
                                    mn_e = min(mn_e, ov);
                                    mx_e = max(mx_e, ov);
                                }
                            }
                        }
                    }
                    hr_sts_u2(hr_smp_addr(sbase + L.smp, pitch, rr, f_own, bi * 8u + chunk), w0, w1);
                }
            }
            seq_base += n_blocks;
            hr_bar_consumers(); // the samples of the group are complete

            // ---- stage 2: the taps. A warp owns whole n-tiles: A fragments from the sample tile, tap fragments from L2 ------
            for (uint32_t nt = wid; nt < n_nt; nt += hr::kConsumerWarps) {
                const int4 m = pp.ntile[st.x + nt]; // {first k-step, last k-step, koff}
                int ac[4] = {0, 0, 0, 0}, th[4] = {0, 0, 0, 0};
                const uint4* bt = pp.btab + (size_t)m.z * 32u + lane;
                const int nk = m.y - m.x + 1;
                // the tap fragments come from L2 (several hundred cycles): four k-steps in flight
                constexpr int kDepth = 4;
                uint4 bf[kDepth];
#pragma unroll
                for (int i = 0; i < kDepth; ++i)
                    if (i < nk) bf[i] = __ldg(bt + (size_t)i * 32u);
                // the lane's 8-column chunk of k-step kk (relative to the strip) is c8 = 4*kk + q: line kk >> 2, rotated position
                // (c8 + f) & 15, which advances by 4 per k-step
                uint32_t kk = (uint32_t)m.x - 2u * st.z;
                uint32_t pa = (4u * kk + q + f_ga) & 15u, pb = (4u * kk + q + f_gb) & 15u;
                uint32_t la = sbase + L.smp + g * pitch + (kk >> 2) * 128u, lb = la + 8u * pitch;
                for (int k0 = 0; k0 < nk; k0 += kDepth) {
#pragma unroll
                    for (int i = 0; i < kDepth; ++i) {
                        if (k0 + i < nk) {
                            const uint4 bcur = bf[i];
                            if (k0 + i + kDepth < nk) bf[i] = __ldg(bt + (size_t)(k0 + i + kDepth) * 32u);
                            const uint2 ra = hr_lds_u2(la + (pa << 3)); // row g
                            const uint2 rb = hr_lds_u2(lb + (pb << 3)); // row g + 8
                            pa = (pa + 4u) & 15u;
                            pb = (pb + 4u) & 15u;
                            ++kk;
                            if ((kk & 3u) == 0) { la += 128u; lb += 128u; }
                            const uint32_t fa[4] = {ra.x, rb.x, ra.y, rb.y};
                            hr_mma_u8s8(th, fa, bcur.x, bcur.y);
                            hr_mma_u8u8(ac, fa, bcur.z, bcur.w);
                        }
                    }
                }
                // scale, clamp, store: lane (g, q) holds columns 2q, 2q+1 of rows g and g+8 of the n-tile
                uint32_t x = 0; // bytes: row g col 2q, row g col 2q+1, row g+8 col 2q, row g+8 col 2q+1
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int vv = (acc0 + ac[i] + (th[i] << 8)) >> prec;
                    vv = vv < 0 ? 0 : (vv > 255 ? 255 : vv);
                    x |= (uint32_t)vv << (8 * i);
                }
                const uint32_t ox = (st.x + nt) * 8u;
                uint8_t* const tA = reinterpret_cast<uint8_t*>(a.temp) + (size_t)(rbase + g - a.row0) * a.ax.out_size + ox;
                uint8_t* const tB = tA + (size_t)8 * a.ax.out_size;
                if (wide_store) { // whole n-tiles and 8-byte aligned rows: the quad's bytes go out as one 8-byte store per row
                    const uint32_t y1 = __shfl_down_sync(FULL, x, 1), y2 = __shfl_down_sync(FULL, x, 2), y3 = __shfl_down_sync(FULL, x, 3);
                    if (q == 0) {
                        if (g < n_valid_rows) *reinterpret_cast<uint2*>(tA) = make_uint2(__byte_perm(x, y1, 0x5410), __byte_perm(y2, y3, 0x5410));
                        if (g + 8u < n_valid_rows) *reinterpret_cast<uint2*>(tB) = make_uint2(__byte_perm(x, y1, 0x7632), __byte_perm(y2, y3, 0x7632));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t col = ox + q * 2u + (i & 1);
                        const uint32_t row = g + ((i & 2) ? 8u : 0u);
                        if (row < n_valid_rows && col < a.ax.out_size) ((i & 2) ? tB : tA)[q * 2u + (i & 1)] = (uint8_t)(x >> (8 * i));
                    }
                }
            }
            if (grp + 1 < n_groups) fold(rbase + 16u); // the next group's table (stage 2 does not read it)
            if (tid == 0) s_ctrl[0] = 0;
            hr_bar_consumers(); // stage 2 is done with the samples; the table and the block counter are ready
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

// ---- host side ---------------------------------------------------------------------------------------------------------
// n-tiles (8 output columns) with their k-step windows and permuted tap bytes, and strips of at most kMaxNt n-tiles whose
// source span is at most min(max_span, kMaxBlocks * 64) columns (CLAHE: max_span = one tile width, so that a strip meets at
// most one cell boundary). Any scale factor fits (no limit on the n-tiles a block touches).
bool hrow_build_plan(const uint32_t* start_h, const uint32_t* size_h, const int32_t* coef_h, uint32_t window, uint32_t out_size,
                     uint32_t in_size, uint32_t max_span, HRowPlanHost* plan) {
    *plan = HRowPlanHost();
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
                        const uint32_t c = ks * 32 + qq * 8 + r * 4 + i; // the order the samples are packed in
                        int32_t tap = 0;
                        if (ox < out_size && c >= start_h[ox] && c < start_h[ox] + size_h[ox]) tap = coef_h[(size_t)ox * window + (c - start_h[ox])];
                        if (tap < -32768 || tap > 32767) return false;
                        const uint32_t lo = (uint32_t)tap & 255u, hi = (uint32_t)(tap >> 8) & 255u;
                        reg[r] |= hi << (8 * i);
                        reg[2 + r] |= lo << (8 * i);
                    }
                plan->btab.push_back(make_uint4(reg[0], reg[1], reg[2], reg[3]));
            }
        koff += lk - fk + 1;
    }
    const uint32_t span_cap = std::min(max_span ? max_span : 0xffffffffu, hr::kMaxBlocks * 64u);
    uint32_t j0 = 0;
    while (j0 < n_nt) {
        uint32_t j1 = j0 + 1;
        auto span = [&](uint32_t e) { return ((uint32_t)plan->ntile[e - 1].y / 2 + 1 - (uint32_t)plan->ntile[j0].x / 2) * 64u; };
        if (span(j1) > span_cap) return false;
        while (j1 < n_nt && j1 - j0 < hr::kMaxNt && span(j1 + 1) <= span_cap) ++j1;
        const uint32_t cb0 = (uint32_t)plan->ntile[j0].x / 2, cb1 = (uint32_t)plan->ntile[j1 - 1].y / 2 + 1;
        plan->strips.push_back(make_uint4(j0, j1, cb0, cb1));
        plan->weights.push_back(HStrip{cb0 * 64, (cb1 - cb0) * 8});
        j0 = j1;
    }
    return true;
}

// Host replay of the kernel's tap stage for one row of u8 samples (test hook): per n-tile its k-steps with the permuted tap
// bytes (hi * 256 + lo), then the final shift / clamp. out has out_size bytes.
bool hrow_replay_row(const HRowPlanHost& plan, const uint8_t* samples, uint32_t in_size, uint32_t out_size, int precision, uint8_t* out) {
    const int acc0 = precision > 0 ? (1 << (precision - 1)) : 0;
    for (const uint4& st : plan.strips)
        for (uint32_t j = st.x; j < st.y; ++j) {
            const int4 m = plan.ntile[j];
            if ((uint32_t)m.x / 2 < st.z || (uint32_t)m.y / 2 >= st.w) return false; // the n-tile's window lies inside the strip's blocks
            int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int ks = m.x; ks <= m.y; ++ks) {
                const uint4* b = plan.btab.data() + (size_t)(m.z + ks - m.x) * 32u;
                for (uint32_t lane = 0; lane < 32; ++lane) {
                    const uint32_t n = lane >> 2, q = lane & 3u;
                    const uint32_t reg[4] = {b[lane].x, b[lane].y, b[lane].z, b[lane].w};
                    for (uint32_t r = 0; r < 2; ++r)
                        for (uint32_t i = 0; i < 4; ++i) {
                            const uint32_t c = (uint32_t)ks * 32 + q * 8 + r * 4 + i;
                            const int smp = c < in_size ? (int)samples[c] : 0; // (the TMA box is zero-filled beyond the raster)
                            const int hi = (int)(int8_t)((reg[r] >> (8 * i)) & 255u), lo = (int)((reg[2 + r] >> (8 * i)) & 255u);
                            acc[n] += smp * (hi * 256 + lo);
                        }
                }
            }
            for (uint32_t n = 0; n < 8; ++n) {
                const uint32_t ox = j * 8 + n;
                int v = (acc0 + acc[n]) >> precision;
                v = v < 0 ? 0 : (v > 255 ? 255 : v);
                if (ox < out_size) out[ox] = (uint8_t)v;
            }
        }
    return true;
}

uint32_t hrow_warps() { return hr::kConsumerWarps; }
size_t hrow_scratch_bytes(uint32_t n_ctas) { return (size_t)n_ctas * 2u * hr::kBins * sizeof(double4); }

// Shared memory for the largest table a plan may ask for (the launch does not know `hot`: it is in device memory).
size_t hrow_smem_bytes(int src_kind) {
    return hrow_layout(src_kind == HSRC_DN_CLAHE, kHmmaMaxHot << hrow_lut_shift(kHmmaMaxHot)).total + 1024; // + alignment slack
}

// The DN raster as a 2-D tensor map: boxes of 64 columns x 16 rows, 128-byte swizzle, zero fill beyond the raster.
cudaError_t hrow_make_tensor_map(void* map128, const void* base, uint64_t cols, uint64_t rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
        if (e != cudaSuccess) return e;
        if (!fn || qr != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {64, 16};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(reinterpret_cast<CUtensorMap*>(map128), CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_hrow(const HResizeArgs& a, int src_kind, const HRowPlanDev& p, const uint32_t* pieces_dev, const uint32_t* cta_first_dev,
                        uint32_t n_ctas, void* scratch, cudaStream_t stream) {
    if (a.n_rows == 0 || a.ax.out_size == 0 || n_ctas == 0) return cudaSuccess;
    if (!a.plan) return cudaErrorInvalidValue;
    const bool clahe = src_kind == HSRC_DN_CLAHE;
    (void)scratch;
    const size_t smem = hrow_smem_bytes(src_kind);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    alignas(64) CUtensorMap tmap;
    if (cudaError_t e = hrow_make_tensor_map(&tmap, a.src, a.src_cols, a.src_rows)) return e;
    HRowParams pp;
    pp.btab = p.btab;
    pp.ntile = p.ntile;
    pp.strips = p.strips;
    pp.pieces = reinterpret_cast<const HPiece*>(pieces_dev);
    pp.cta_first = cta_first_dev;
    if (clahe) {
        if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_hrow<true>), smem)) return e;
        k_hrow<true><<<n_ctas, hr::kThreads, smem, stream>>>(a, pp, tmap);
    } else {
        if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_hrow<false>), smem)) return e;
        k_hrow<false><<<n_ctas, hr::kThreads, smem, stream>>>(a, pp, tmap);
    }
    return cudaGetLastError();
}

} // namespace sarpro

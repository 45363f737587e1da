// kernels_resize.cu — separable fixed-point Lanczos3 (resize.rs:32-89 -> fast_image_resize 5.x
// ResizeAlg::Convolution(FilterType::Lanczos3)), horizontal pass first into a same-type temporary,
// then vertical (see plan.cpp / DESIGN.md for the restated third-party arithmetic).
//
// The horizontal pass is where the full-resolution raster is consumed, so it is fused with the
// per-pixel stage: its loader is a template over the pixel source
//   HSRC_IMAGE     a u8/u16 image (stand-alone resize_image_data_with_meta)
//   HSRC_DN_LUT    u16 DN -> LUT (autoscale + scale_u16_to_u8 folded by the planner)
//   HSRC_DN_CLAHE  u16 DN -> CLAHE bin LUT -> exact bilinear CDF blend -> quantise -> remap
// Each CTA owns a strip of output columns; the source span of the strip (plus the Lanczos halo)
// is staged row by row in shared memory as final u8/u16 samples, then every thread accumulates
// one output column with dp2a (two i16 taps x two u8 samples per instruction).
#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace sarpro {

constexpr uint32_t kHLutHot = 8192;
constexpr int kRowsPerGroup = 4;

__device__ __forceinline__ double blend_exact(double c00, double c01, double c10, double c11, double dx, double omdx,
                                              double dy, double omdy) {
    const double top = __dadd_rn(__dmul_rn(c00, omdx), __dmul_rn(c01, dx));
    const double bottom = __dadd_rn(__dmul_rn(c10, omdx), __dmul_rn(c11, dx));
    return __dadd_rn(__dmul_rn(top, omdy), __dmul_rn(bottom, dy));
}

// Loads 8 consecutive source samples (row r, columns c..c+7) as final pixel values.
template <int SRC, bool PIX16>
struct Loader {
    const HResizeArgs& a;
    const void* s_lut;     // shared: u8[kHLutHot] or u16[kHLutHot]
    const uint8_t* s_remap; // shared 256 or nullptr
    bool aligned;
    uint32_t mn, mx;

    __device__ __forceinline__ uint32_t look(uint32_t d) const {
        if (PIX16) return d < kHLutHot ? (uint32_t) reinterpret_cast<const uint16_t*>(s_lut)[d] : (uint32_t)__ldg(&a.lut[d]);
        return d < kHLutHot ? (uint32_t) reinterpret_cast<const uint8_t*>(s_lut)[d] : (uint32_t)(__ldg(&a.lut[d]) & 255u);
    }

    __device__ __forceinline__ void load_dn(uint32_t r, uint32_t c, uint32_t d[8]) const {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(a.src) + (size_t)r * a.src_cols + c;
        if (aligned && c + 8 <= a.src_cols) {
            const uint4 q = ld_stream_u4(p);
            d[0] = q.x & 0xffffu; d[1] = q.x >> 16; d[2] = q.y & 0xffffu; d[3] = q.y >> 16;
            d[4] = q.z & 0xffffu; d[5] = q.z >> 16; d[6] = q.w & 0xffffu; d[7] = q.w >> 16;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] = (c + k < a.src_cols) ? (uint32_t)p[k] : 0u;
        }
    }
    // true raster width: src_cols is the row pitch, which exceeds it by 1..7 padding columns on a re-pitched raster
    __device__ __forceinline__ uint32_t width() const { return a.src_width ? a.src_width : a.src_cols; }

    __device__ __forceinline__ void get(uint32_t r, uint32_t c, uint32_t o[8]) {
        if (SRC == HSRC_IMAGE) {
            if (PIX16) {
                load_dn(r, c, o);
            } else {
                const uint8_t* p = reinterpret_cast<const uint8_t*>(a.src) + (size_t)r * a.src_cols + c;
                if (aligned && c + 8 <= a.src_cols) {
                    const uint2 q = ld_stream_u2(p);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { o[k] = (q.x >> (8 * k)) & 255u; o[4 + k] = (q.y >> (8 * k)) & 255u; }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[k] = (c + k < a.src_cols) ? (uint32_t)p[k] : 0u;
                }
            }
        } else if (SRC == HSRC_DN_LUT) {
            uint32_t d[8];
            load_dn(r, c, d);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = look(d[k]);
        } else {
            uint32_t d[8];
            load_dn(r, c, d);
            const ClaheDev& cl = a.clahe;
            const double dy = cl.row_dy[r], omdy = cl.row_omdy[r];
            const uint32_t ty = cl.row_t[r];
            const double* cdf_t0 = cl.cdf + (size_t)(ty & 7u) * cl.tiles_x * 256u;
            const double* cdf_t1 = cl.cdf + (size_t)((ty >> 8) & 7u) * cl.tiles_x * 256u;
            const double max_val = PIX16 ? 65535.0 : 255.0;
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                uint32_t v = 0;
                const uint32_t cc = c + k;
                if (cc < width()) { // (padding columns carry no tap and must not enter the min / max)
                    if (d[k] != 0) {
                        const uint32_t bin = look(d[k]) & 255u;
                        const uint32_t tx = cl.col_t[cc];
                        const uint32_t x0 = (tx & 7u) * 256u + bin, x1 = ((tx >> 8) & 7u) * 256u + bin;
                        double bl = blend_exact(cdf_t0[x0], cdf_t0[x1], cdf_t1[x0], cdf_t1[x1], cl.col_dx[cc],
                                                cl.col_omdx[cc], dy, omdy);
                        bl = bl < 0.0 ? 0.0 : (bl > 1.0 ? 1.0 : bl);
                        v = (uint32_t)__dmul_rn(bl, max_val);
                    }
                    mn = min(mn, v);
                    mx = max(mx, v);
                    if (!PIX16 && s_remap) v = s_remap[v];
                }
                o[k] = v;
            }
        }
    }
};

template <int SRC, bool PIX16>
__global__ void __launch_bounds__(256) k_hresize(HResizeArgs a, const HStrip* __restrict__ strips, uint32_t rbw,
                                                 uint32_t rows_per_block) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (a.skip && *a.skip) return;
    if (a.run_if && !*a.run_if) return;
    using Pix = typename std::conditional<PIX16, uint16_t, uint8_t>::type;
    const uint32_t tid = threadIdx.x, oxb = blockDim.x;
    const uint32_t ox = blockIdx.x * oxb + tid;
    const bool have_ox = ox < a.ax.out_size;
    const HStrip st = strips[blockIdx.x];
    const uint32_t ntab = PIX16 ? a.ax.window : a.ax.pairs;

    // shared carve-up: taps [ntab][oxb] u32 | row buffers [R][rbw] bytes | lut | remap
    uint32_t* s_tap = reinterpret_cast<uint32_t*>(smem);
    unsigned char* s_rows = smem + (size_t)ntab * oxb * 4;
    unsigned char* s_lut = s_rows + (size_t)kRowsPerGroup * rbw;
    uint8_t* s_remap = s_lut + (SRC == HSRC_IMAGE ? 0 : kHLutHot * sizeof(Pix));

    for (uint32_t i = tid; i < ntab * oxb; i += oxb) {
        const uint32_t t = i / oxb, x = i % oxb;
        const uint32_t gx = blockIdx.x * oxb + x;
        uint32_t v = 0;
        if (gx < a.ax.out_size) v = PIX16 ? (uint32_t)a.ax.coef[(size_t)gx * a.ax.window + t] : a.ax.packed[(size_t)gx * a.ax.pairs + t];
        s_tap[t * oxb + x] = v;
    }
    if (SRC != HSRC_IMAGE) {
        for (uint32_t i = tid; i < kHLutHot; i += oxb) reinterpret_cast<Pix*>(s_lut)[i] = (Pix)a.lut[i];
        if (SRC == HSRC_DN_CLAHE && !PIX16 && a.remap)
            for (uint32_t i = tid; i < 256; i += oxb) s_remap[i] = a.remap[i];
    }
    // zero the row buffers once: the tail padding is read (with zero taps) but never staged
    for (uint32_t i = tid; i < kRowsPerGroup * rbw / 4; i += oxb) reinterpret_cast<uint32_t*>(s_rows)[i] = 0;
    __syncthreads();

    Loader<SRC, PIX16> ld{a, s_lut, (SRC == HSRC_DN_CLAHE && !PIX16 && a.remap) ? s_remap : nullptr, false, 0xffffffffu, 0};
    {
        const uintptr_t p = reinterpret_cast<uintptr_t>(a.src);
        const uint32_t esz = (SRC == HSRC_IMAGE && !PIX16) ? 1 : 2;
        ld.aligned = (p % (8 * esz) == 0) && (a.src_cols % 8 == 0);
    }
    const uint32_t start = have_ox ? a.ax.start[ox] : 0;
    const uint32_t nsize = have_ox ? a.ax.size[ox] : 0;
    // u8: tap 0 of the packed table sits at source (start & ~3); word offset into the row buffer
    const uint32_t woff = have_ox ? (((start & ~3u) - st.sc0) >> 2) : 0;
    const uint32_t poff = have_ox ? (start - st.sc0) : 0;
    const int prec = a.ax.precision;

    const uint32_t rb0 = blockIdx.y * rows_per_block;
    const uint32_t rb1 = min(rb0 + rows_per_block, a.n_rows);
    for (uint32_t g = rb0; g < rb1; g += kRowsPerGroup) {
        const uint32_t nr = min((uint32_t)kRowsPerGroup, rb1 - g);
        // ---- stage nr rows of final samples -------------------------------------------------
        const uint32_t nv = nr * st.nvec;
        for (uint32_t v = tid; v < nv; v += oxb) {
            const uint32_t rr = v / st.nvec, vv = v % st.nvec;
            uint32_t o[8];
            ld.get(a.row0 + g + rr, st.sc0 + vv * 8, o);
            if (PIX16) {
                uint4 p;
                p.x = o[0] | (o[1] << 16); p.y = o[2] | (o[3] << 16); p.z = o[4] | (o[5] << 16); p.w = o[6] | (o[7] << 16);
                *reinterpret_cast<uint4*>(s_rows + (size_t)rr * rbw + vv * 16) = p;
            } else {
                uint2 p;
                p.x = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
                p.y = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
                *reinterpret_cast<uint2*>(s_rows + (size_t)rr * rbw + vv * 8) = p;
            }
        }
        __syncthreads();
        // ---- one output column per thread ---------------------------------------------------
        if (have_ox) {
            if (!PIX16) {
                int acc[kRowsPerGroup];
#pragma unroll
                for (int rr = 0; rr < kRowsPerGroup; ++rr) acc[rr] = prec > 0 ? (1 << (prec - 1)) : 0;
                const uint32_t nwords = a.ax.pairs >> 1;
                for (uint32_t m = 0; m < nwords; ++m) {
                    const int t0 = (int)s_tap[(2 * m) * oxb + tid], t1 = (int)s_tap[(2 * m + 1) * oxb + tid];
#pragma unroll
                    for (int rr = 0; rr < kRowsPerGroup; ++rr) {
                        const uint32_t w = reinterpret_cast<const uint32_t*>(s_rows + (size_t)rr * rbw)[woff + m];
                        acc[rr] = dp2a_lo_su(t0, w, acc[rr]);
                        acc[rr] = dp2a_hi_su(t1, w, acc[rr]);
                    }
                }
#pragma unroll
                for (int rr = 0; rr < kRowsPerGroup; ++rr)
                    if ((uint32_t)rr < nr) {
                        int v = acc[rr] >> prec;
                        v = v < 0 ? 0 : (v > 255 ? 255 : v);
                        reinterpret_cast<uint8_t*>(a.temp)[(size_t)(g + rr) * a.ax.out_size + ox] = (uint8_t)v;
                    }
            } else {
                long long acc[kRowsPerGroup];
#pragma unroll
                for (int rr = 0; rr < kRowsPerGroup; ++rr) acc[rr] = prec > 0 ? (1ll << (prec - 1)) : 0ll;
                for (uint32_t k = 0; k < nsize; ++k) {
                    const long long t = (long long)(int)s_tap[k * oxb + tid];
#pragma unroll
                    for (int rr = 0; rr < kRowsPerGroup; ++rr) {
                        const uint32_t px = reinterpret_cast<const uint16_t*>(s_rows + (size_t)rr * rbw)[poff + k];
                        acc[rr] += t * (long long)px;
                    }
                }
#pragma unroll
                for (int rr = 0; rr < kRowsPerGroup; ++rr)
                    if ((uint32_t)rr < nr) {
                        long long v = acc[rr] >> prec;
                        v = v < 0 ? 0 : (v > 65535 ? 65535 : v);
                        reinterpret_cast<uint16_t*>(a.temp)[(size_t)(g + rr) * a.ax.out_size + ox] = (uint16_t)v;
                    }
            }
        }
        __syncthreads();
    }
    if (SRC == HSRC_DN_CLAHE && a.minmax) {
        const uint32_t mn = warp_reduce_min(ld.mn), mx = warp_reduce_max(ld.mx);
        if ((tid & 31) == 0 && mn != 0xffffffffu) {
            atomicMin(&a.minmax[0], mn);
            atomicMax(&a.minmax[1], mx);
        }
    }
}


// Builds the strip table for an axis (host arrays needed): exported for the context.
cudaError_t hresize_build_strips(const uint32_t* start_h, const uint32_t* size_h, uint32_t out_size, uint32_t in_size,
                                 uint32_t window, uint32_t pairs, int pix16, int src_kind, uint32_t* oxb_out,
                                 std::vector<HStrip>* strips, uint32_t* rbw_out, uint32_t* smem_out) {
    const uint32_t esz = pix16 ? 2 : 1;
    const uint32_t ntab = pix16 ? window : pairs;
    uint32_t oxb = 256;
    for (;;) {
        strips->clear();
        uint32_t max_nvec = 0;
        const uint32_t n_strips = (out_size + oxb - 1) / oxb;
        for (uint32_t s = 0; s < n_strips; ++s) {
            const uint32_t ox0 = s * oxb, ox1 = std::min(out_size, ox0 + oxb);
            uint32_t lo = start_h[ox0], hi = 0;
            for (uint32_t x = ox0; x < ox1; ++x) {
                lo = std::min(lo, start_h[x]);
                hi = std::max(hi, start_h[x] + size_h[x]);
            }
            if (s == 0) lo = 0;                 // every source column is staged by some strip
            if (s + 1 == n_strips) hi = in_size; // (CLAHE min/max must see the whole raster)
            HStrip st;
            st.sc0 = lo & ~7u;
            st.nvec = (hi - st.sc0 + 7) / 8;
            max_nvec = std::max(max_nvec, st.nvec);
            strips->push_back(st);
        }
        // row buffer: staged samples + room for the zero-tap overrun of the packed table (<= window+8 samples)
        uint32_t rbw = (max_nvec * 8 + window + 16) * esz;
        rbw = (rbw + 15) & ~15u;
        uint32_t smem = ntab * oxb * 4 + kRowsPerGroup * rbw;
        if (src_kind != HSRC_IMAGE) smem += kHLutHot * esz + 256;
        if (smem <= 200 * 1024 || oxb == 32) {
            *oxb_out = oxb;
            *rbw_out = rbw;
            *smem_out = smem;
            return smem <= 227 * 1024 ? cudaSuccess : cudaErrorInvalidConfiguration;
        }
        oxb >>= 1;
    }
}

template <int SRC, bool PIX16>
static cudaError_t launch_hresize_t(const HResizeArgs& a, const HStrip* strips_dev, uint32_t n_strips, uint32_t oxb,
                                    uint32_t rbw, uint32_t smem, int sm_count, cudaStream_t stream) {
    if (cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(&k_hresize<SRC, PIX16>), smem)) return e;
    // rows per block: enough blocks to fill the machine a few times over, multiple of the row group
    uint32_t target_blocks = (uint32_t)sm_count * 8;
    uint32_t yblocks = std::max(1u, target_blocks / std::max(1u, n_strips));
    uint32_t rpb = (a.n_rows + yblocks - 1) / yblocks;
    rpb = ((rpb + kRowsPerGroup - 1) / kRowsPerGroup) * kRowsPerGroup;
    if (rpb < (uint32_t)kRowsPerGroup * 4) rpb = kRowsPerGroup * 4;
    yblocks = (a.n_rows + rpb - 1) / rpb;
    k_hresize<SRC, PIX16><<<dim3(n_strips, yblocks), oxb, smem, stream>>>(a, strips_dev, rbw, rpb);
    return cudaGetLastError();
}

cudaError_t launch_hresize_planned(const HResizeArgs& a, int src_kind, int pix16, const HStrip* strips_dev,
                                   uint32_t n_strips, uint32_t oxb, uint32_t rbw, uint32_t smem, int sm_count,
                                   cudaStream_t stream) {
    if (a.n_rows == 0 || a.ax.out_size == 0) return cudaSuccess;
#define SARPRO_H(S, P) return launch_hresize_t<S, P>(a, strips_dev, n_strips, oxb, rbw, smem, sm_count, stream)
    if (!pix16) {
        if (src_kind == HSRC_IMAGE) SARPRO_H(HSRC_IMAGE, false);
        if (src_kind == HSRC_DN_LUT) SARPRO_H(HSRC_DN_LUT, false);
        SARPRO_H(HSRC_DN_CLAHE, false);
    } else {
        if (src_kind == HSRC_IMAGE) SARPRO_H(HSRC_IMAGE, true);
        if (src_kind == HSRC_DN_LUT) SARPRO_H(HSRC_DN_LUT, true);
        SARPRO_H(HSRC_DN_CLAHE, true);
    }
#undef SARPRO_H
}

// ---------------------------------------------------------------------------------------------
// Vertical pass: temp [rows][width] -> out rows [oy0, oy1)
// ---------------------------------------------------------------------------------------------
template <bool PIX16>
__global__ void __launch_bounds__(256) k_vresize(const void* __restrict__ temp, uint32_t temp_row0, uint32_t width,
                                                 AxisDev ax, uint32_t oy0, void* __restrict__ out, uint32_t out_pitch,
                                                 uint32_t out_x0, const uint32_t* __restrict__ skip,
                                                 const uint32_t* __restrict__ run_if) {
    extern __shared__ int s_coef[];
    if (skip && *skip) return;
    if (run_if && !*run_if) return;
    const uint32_t oy = oy0 + blockIdx.y;
    const uint32_t n = ax.size[oy], start = ax.start[oy];
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) s_coef[k] = ax.coef[(size_t)oy * ax.window + k];
    __syncthreads();
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    const int prec = ax.precision;
    if (!PIX16) {
        const uint8_t* t = reinterpret_cast<const uint8_t*>(temp) + (size_t)(start - temp_row0) * width + x;
        int acc = prec > 0 ? (1 << (prec - 1)) : 0;
        for (uint32_t k = 0; k < n; ++k) acc += s_coef[k] * (int)t[(size_t)k * width];
        int v = acc >> prec;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        reinterpret_cast<uint8_t*>(out)[(size_t)oy * out_pitch + out_x0 + x] = (uint8_t)v;
    } else {
        const uint16_t* t = reinterpret_cast<const uint16_t*>(temp) + (size_t)(start - temp_row0) * width + x;
        long long acc = prec > 0 ? (1ll << (prec - 1)) : 0ll;
        for (uint32_t k = 0; k < n; ++k) acc += (long long)s_coef[k] * (long long)t[(size_t)k * width];
        long long v = acc >> prec;
        v = v < 0 ? 0 : (v > 65535 ? 65535 : v);
        reinterpret_cast<uint16_t*>(out)[(size_t)oy * out_pitch + out_x0 + x] = (uint16_t)v;
    }
}
// u8 vertical pass, 4 adjacent columns per thread: two source rows are byte-interleaved (PRMT) so that one
// dp2a consumes (tap k, tap k+1) x (row k, row k+1) of a column.
__global__ void __launch_bounds__(128) k_vresize8x4(const uint8_t* __restrict__ temp, uint32_t temp_row0, uint32_t width,
                                                    AxisDev ax, uint32_t oy0, uint8_t* __restrict__ out, uint32_t out_pitch,
                                                    uint32_t out_x0, const uint32_t* __restrict__ skip,
                                                    const uint32_t* __restrict__ run_if) {
    extern __shared__ int s_pair[]; // packed (tap k | tap k+1 << 16)
    if (skip && *skip) return;
    if (run_if && !*run_if) return;
    const uint32_t oy = oy0 + blockIdx.y;
    const uint32_t n = ax.size[oy], start = ax.start[oy];
    const uint32_t np = (n + 1) / 2;
    for (uint32_t k = threadIdx.x; k < np; k += blockDim.x) {
        const int c0 = ax.coef[(size_t)oy * ax.window + 2 * k];
        const int c1 = 2 * k + 1 < n ? ax.coef[(size_t)oy * ax.window + 2 * k + 1] : 0;
        s_pair[k] = (int)(((uint32_t)c0 & 0xffffu) | ((uint32_t)c1 << 16));
    }
    __syncthreads();
    const uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x >= width) return;
    const int prec = ax.precision;
    const int init = prec > 0 ? (1 << (prec - 1)) : 0;
    int a0 = init, a1 = init, a2 = init, a3 = init;
    const uint8_t* t = temp + (size_t)(start - temp_row0) * width + x;
    const uint32_t last = n - 1;
    for (uint32_t k = 0; k < np; ++k) {
        const uint32_t r0 = 2 * k, r1 = min(2 * k + 1, last); // odd tail: tap is 0, row index stays in range
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(t + (size_t)r0 * width);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(t + (size_t)r1 * width);
        const uint32_t lo = __byte_perm(w0, w1, 0x5140); // c0r0 c0r1 c1r0 c1r1
        const uint32_t hi = __byte_perm(w0, w1, 0x7362); // c2r0 c2r1 c3r0 c3r1
        const int p = s_pair[k];
        a0 = dp2a_lo_su(p, lo, a0);
        a1 = dp2a_hi_su(p, lo, a1);
        a2 = dp2a_lo_su(p, hi, a2);
        a3 = dp2a_hi_su(p, hi, a3);
    }
    auto clip = [&](int a) { int v = a >> prec; return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
    uint8_t* o = out + (size_t)oy * out_pitch + out_x0 + x;
    const uint32_t packed = clip(a0) | (clip(a1) << 8) | (clip(a2) << 16) | (clip(a3) << 24);
    if (x + 4 <= width && (reinterpret_cast<uintptr_t>(o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = packed;
    else
        for (uint32_t j = 0; j < 4 && x + j < width; ++j) o[j] = (uint8_t)(packed >> (8 * j));
}

cudaError_t launch_vresize(const void* temp, uint32_t temp_row0, uint32_t width, AxisDev ax, uint32_t oy0, uint32_t oy1,
                           void* out, uint32_t out_pitch, uint32_t out_x0, int pix16, cudaStream_t stream,
                           const uint32_t* skip, const uint32_t* run_if) {
    if (oy1 <= oy0 || width == 0) return cudaSuccess;
    if (!pix16 && width % 4 == 0 && (reinterpret_cast<uintptr_t>(temp) & 3) == 0) {
        const dim3 grid((width / 4 + 127) / 128, oy1 - oy0);
        const size_t smem = (size_t)(ax.window + 1) / 2 * sizeof(int) + 16;
        k_vresize8x4<<<grid, 128, smem, stream>>>((const uint8_t*)temp, temp_row0, width, ax, oy0, (uint8_t*)out, out_pitch, out_x0, skip, run_if);
        return cudaGetLastError();
    }
    const dim3 grid((width + 255) / 256, oy1 - oy0);
    const size_t smem = (size_t)ax.window * sizeof(int);
    if (pix16) k_vresize<true><<<grid, 256, smem, stream>>>(temp, temp_row0, width, ax, oy0, out, out_pitch, out_x0, skip, run_if);
    else k_vresize<false><<<grid, 256, smem, stream>>>(temp, temp_row0, width, ax, oy0, out, out_pitch, out_x0, skip, run_if);
    return cudaGetLastError();
}

} // namespace sarpro

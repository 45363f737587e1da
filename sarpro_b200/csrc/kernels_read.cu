// kernels_read.cu — downsample-on-read: the raw raster (u16 DN, or f32) resampled to the reader's target shape before autoscale
// (sentinel1.rs:1074-1109 -> gdal.rs:145-177: GDAL RasterIO with ResampleAlg::Average for reductions >= 4, Lanczos below).
// The arithmetic is the published one of GDAL >= 3.3 (gcore/overview.cpp), restated in plan_read.cpp / oracle_read.cpp;
// parity unpinned. f64 accumulation in the reference's order (rows outer, columns inner), no contraction
// (--fmad=false), so the result is bit-identical to the oracle's.
//
// k_read_average: one warp per (output row, 32 output columns). For every source row of the output row's span the warp
// stages the contiguous source segment its 32 output pixels cover (<= kSpanMax samples) in shared memory with coalesced
// loads, then each lane accumulates its own column span from there: the raster is read from HBM exactly once (2 B per pixel),
// whatever the reduction factor.
#include "common.cuh"
#include "kernels.h"

namespace sarpro {

namespace rd {
constexpr uint32_t kWarps = 8;       // per CTA
constexpr uint32_t kSpanMax = 1536;  // source samples a warp stages per row (32 output columns x reduction <= 47)
}

// exact u16 -> f64 on the FP64 add pipe (2^52 + v is exact; the subtraction removes the bias) instead of the quarter-rate I2F
__device__ __forceinline__ double rd_to_double(uint16_t v) { return __dadd_rn(__hiloint2double(0x43300000, (int)v), -4503599627370496.0); }
__device__ __forceinline__ double rd_to_double(float v) { return (double)v; }

template <typename T>
__global__ void __launch_bounds__(rd::kWarps * 32) k_read_average(const T* __restrict__ src, uint32_t rows, uint32_t cols, ReadAvgAxis ax,
                                                                 ReadAvgAxis ay, float* __restrict__ out, uint32_t out_rows, uint32_t out_cols) {
    __shared__ __align__(16) T s_row[rd::kWarps][rd::kSpanMax];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t oy = blockIdx.y * rd::kWarps + warp;
    if (oy >= out_rows) return;
    const uint32_t ox0 = blockIdx.x * 32u, ox = ox0 + lane;
    const bool live = ox < out_cols;
    const uint32_t oxl = min(ox0 + 31u, out_cols - 1u);
    // the segment starts on a 16-byte boundary of the row when the raster allows 128-bit loads (rows 16-byte aligned)
    constexpr int kVec = 16 / (int)sizeof(T);
    const bool vec = (cols % kVec) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    const int seg0 = vec ? (ax.start[ox0] & ~(kVec - 1)) : ax.start[ox0], seg1 = ax.end[oxl];
    const int xs = live ? ax.start[ox] : seg0, xe = live ? ax.end[ox] : seg0;
    const double wxf = live ? ax.w_first[ox] : 1.0, wxl = live ? ax.w_last[ox] : 1.0;
    const int ys = ay.start[oy], ye = ay.end[oy];
    const double wyf = ay.w_first[oy], wyl = ay.w_last[oy];
    const bool staged = (uint32_t)(seg1 - seg0) + (uint32_t)kVec <= rd::kSpanMax;
    double total = 0.0, wsum = 0.0;
    T* const s = s_row[warp];
    for (int y = ys; y < ye; ++y) {
        const double wy = y == ys ? wyf : (y + 1 == ye ? wyl : 1.0);
        const T* row = src + (size_t)y * cols;
        if (staged) {
            if (vec) { // whole 16-byte vectors (the row length is a multiple of kVec, so the last one stays inside the row)
                const uint4* rv = reinterpret_cast<const uint4*>(row + seg0);
                uint4* sv = reinterpret_cast<uint4*>(s);
                const int nv = (seg1 - seg0 + kVec - 1) / kVec;
                for (int i = (int)lane; i < nv; i += 32) sv[i] = __ldg(rv + i);
            } else {
                for (int x = seg0 + (int)lane; x < seg1; x += 32) s[x - seg0] = row[x];
            }
            __syncwarp();
        }
        // Same operations in the same order as the reference loop (w = wy * wx; total += v * w; wsum += w), with the factors
        // that are exactly 1.0 elided: wy * 1.0 == wy and v * 1.0 == v bit for bit, so interior columns cost one
        // multiplication less and interior rows none at all.
        const T* px = staged ? s - seg0 : row;
        if (xs < xe) {
            {   // first column of the span (also the only one of a one-column span)
                const double w = __dmul_rn(wy, wxf);
                total = __dadd_rn(total, __dmul_rn(rd_to_double(px[xs]), w));
                wsum = __dadd_rn(wsum, w);
            }
            if (wy == 1.0) { // warp-uniform: an interior row
                for (int x = xs + 1; x + 1 < xe; ++x) {
                    total = __dadd_rn(total, rd_to_double(px[x]));
                    wsum = __dadd_rn(wsum, 1.0);
                }
            } else {
                for (int x = xs + 1; x + 1 < xe; ++x) {
                    total = __dadd_rn(total, __dmul_rn(rd_to_double(px[x]), wy));
                    wsum = __dadd_rn(wsum, wy);
                }
            }
            if (xe - xs >= 2) { // last column
                const double w = __dmul_rn(wy, wxl);
                total = __dadd_rn(total, __dmul_rn(rd_to_double(px[xe - 1]), w));
                wsum = __dadd_rn(wsum, w);
            }
        }
        if (staged) __syncwarp();
    }
    if (live) out[(size_t)oy * out_cols + ox] = (float)__ddiv_rn(total, wsum);
}

template <typename T>
static cudaError_t launch_read_average_t(const T* src, uint32_t rows, uint32_t cols, const ReadAvgAxis& ax, const ReadAvgAxis& ay, float* out,
                                         uint32_t out_rows, uint32_t out_cols, cudaStream_t stream) {
    const dim3 grid((out_cols + 31) / 32, (out_rows + rd::kWarps - 1) / rd::kWarps);
    k_read_average<T><<<grid, rd::kWarps * 32, 0, stream>>>(src, rows, cols, ax, ay, out, out_rows, out_cols);
    return cudaGetLastError();
}
cudaError_t launch_read_average(const void* src, int src_u16, uint32_t rows, uint32_t cols, const ReadAvgAxis& ax, const ReadAvgAxis& ay,
                                float* out, uint32_t out_rows, uint32_t out_cols, cudaStream_t stream) {
    if (out_rows == 0 || out_cols == 0) return cudaSuccess;
    return src_u16 ? launch_read_average_t((const uint16_t*)src, rows, cols, ax, ay, out, out_rows, out_cols, stream)
                   : launch_read_average_t((const float*)src, rows, cols, ax, ay, out, out_rows, out_cols, stream);
}

// ---- Lanczos (mild reductions): separable, horizontal first into an f64 intermediate ---------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_read_conv_h(const T* __restrict__ src, uint32_t rows, uint32_t cols, ReadConvAxis ax,
                                                     double* __restrict__ tmp, uint32_t out_cols) {
    const uint32_t dx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (dx >= out_cols) return;
    const T* row = src + (size_t)y * cols + ax.start[dx];
    const double* w = ax.w + (size_t)dx * ax.window;
    double v = 0.0;
    for (int k = 0; k < ax.count[dx]; ++k) v = __dadd_rn(v, __dmul_rn((double)row[k], w[k]));
    tmp[(size_t)y * out_cols + dx] = v;
}
__global__ void __launch_bounds__(256) k_read_conv_v(const double* __restrict__ tmp, ReadConvAxis ay, float* __restrict__ out,
                                                     uint32_t out_rows, uint32_t out_cols) {
    const uint32_t dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= out_cols) return;
    const double* w = ay.w + (size_t)dy * ay.window;
    const double* col = tmp + (size_t)ay.start[dy] * out_cols + dx;
    double v = 0.0;
    for (int k = 0; k < ay.count[dy]; ++k) v = __dadd_rn(v, __dmul_rn(col[(size_t)k * out_cols], w[k]));
    out[(size_t)dy * out_cols + dx] = (float)v;
}
cudaError_t launch_read_lanczos(const void* src, int src_u16, uint32_t rows, uint32_t cols, const ReadConvAxis& ax, const ReadConvAxis& ay,
                                double* tmp, float* out, uint32_t out_rows, uint32_t out_cols, cudaStream_t stream) {
    if (out_rows == 0 || out_cols == 0) return cudaSuccess;
    const dim3 gh((out_cols + 255) / 256, rows), gv((out_cols + 255) / 256, out_rows);
    if (src_u16) k_read_conv_h<uint16_t><<<gh, 256, 0, stream>>>((const uint16_t*)src, rows, cols, ax, tmp, out_cols);
    else k_read_conv_h<float><<<gh, 256, 0, stream>>>((const float*)src, rows, cols, ax, tmp, out_cols);
    if (cudaError_t e = cudaGetLastError()) return e;
    k_read_conv_v<<<gv, 256, 0, stream>>>(tmp, ay, out, out_rows, out_cols);
    return cudaGetLastError();
}

} // namespace sarpro

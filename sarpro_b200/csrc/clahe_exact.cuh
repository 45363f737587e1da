// clahe_exact.cuh — the CLAHE sample of one pixel with the reference's exact f64 operation order
// (autoscale.rs:320-329 then :602-606), shared by the pass-B kernels for their exact fix-up paths.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace sarpro {

__device__ __forceinline__ double clahe_blend_exact_rn(double c00, double c01, double c10, double c11, double dx,
                                                       double omdx, double dy, double omdy) {
    const double top = __dadd_rn(__dmul_rn(c00, omdx), __dmul_rn(c01, dx));      // autoscale.rs:327
    const double bottom = __dadd_rn(__dmul_rn(c10, omdx), __dmul_rn(c11, dx));   // :328
    return __dadd_rn(__dmul_rn(top, omdy), __dmul_rn(bottom, dy));               // :329
}

// u8 sample of pixel (r, c) (local row, column) with DN d != 0. Arguments by value so that the kernel-parameter
// struct never has to be materialised in local memory.
static __device__ __noinline__ uint32_t clahe_exact_sample_impl(const uint16_t* __restrict__ lut, const double* __restrict__ cdf,
                                                         const double* __restrict__ col_dx, const double* __restrict__ col_omdx,
                                                         const uint16_t* __restrict__ col_t, const double* __restrict__ row_dy,
                                                         const double* __restrict__ row_omdy, const uint16_t* __restrict__ row_t,
                                                         uint32_t r, uint32_t c, uint32_t d) {
    const uint32_t bin = lut[d] & 255u;
    const uint32_t ty = row_t[r], tx = col_t[c];
    const double* t0 = cdf + (size_t)(ty & 7u) * 8u * 256u;
    const double* t1 = cdf + (size_t)((ty >> 8) & 7u) * 8u * 256u;
    const uint32_t x0 = (tx & 7u) * 256u + bin, x1 = ((tx >> 8) & 7u) * 256u + bin;
    double v = clahe_blend_exact_rn(t0[x0], t0[x1], t1[x0], t1[x1], col_dx[c], col_omdx[c], row_dy[r], row_omdy[r]);
    v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    return (uint32_t)__dmul_rn(v, 255.0);
}
// Column geometry of autoscale.rs:308-318 for global column g, bit-identical to k_clahe_axis (same f64 operations):
// returns dx, 1-dx and the clamped tile pair (t0, t1).
__device__ __forceinline__ void clahe_axis_exact(uint32_t g, uint32_t tile_size, uint32_t n_tiles, double* d, double* omd,
                                                 uint32_t* t0, uint32_t* t1) {
    const double f = __dsub_rn(__ddiv_rn((double)g, (double)tile_size), 0.5);
    const long long t = (long long)fmax(floor(f), 0.0);
    const double dd = __dsub_rn(f, (double)t);
    const long long hi = (long long)n_tiles - 1;
    *d = dd;
    *omd = __dsub_rn(1.0, dd);
    *t0 = (uint32_t)(t < 0 ? 0 : (t > hi ? hi : t));
    *t1 = (uint32_t)((t + 1) < 0 ? 0 : ((t + 1) > hi ? hi : (t + 1)));
}

#define clahe_exact_sample(LUT, CL, R, C, D) \
    clahe_exact_sample_impl((LUT), (CL).cdf, (CL).col_dx, (CL).col_omdx, (CL).col_t, (CL).row_dy, (CL).row_omdy, (CL).row_t, (R), (C), (D))

} // namespace sarpro

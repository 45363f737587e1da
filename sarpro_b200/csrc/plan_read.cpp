// plan_read.cpp — host tables of the downsample-on-read kernels (kernels_read.cu): the per-axis source spans and weights of
// GDAL's RasterIO resampling as the reference calls it (gdal.rs:145-177 read_band_resampled; Average / Lanczos picked at
// sentinel1.rs:1092-1102). The library itself (system libgdal, version unpinned) is not part of the reference tree: this is
// the published algorithm of GDAL >= 3.3 (gcore/overview.cpp), parity unpinned.
#include <algorithm>
#include <cmath>

#include "plan.h"

namespace sarpro {

// sentinel1.rs:1083-1102: aspect-preserving output shape for a long-side target (never upscales) and the resampler:
// Average for a reduction of 4 or more, Lanczos below.
void read_dims_for_target(uint64_t cols, uint64_t rows, uint64_t target, uint64_t* out_cols, uint64_t* out_rows, int* alg) {
    const uint64_t long_side = std::max(cols, rows);
    const double scale = std::min((double)target / (double)long_side, 1.0);
    *out_cols = (uint64_t)std::max(std::round((double)cols * scale), 1.0);
    *out_rows = (uint64_t)std::max(std::round((double)rows * scale), 1.0);
    const double reduction = std::max((double)long_side / (double)target, 1.0);
    *alg = reduction >= 4.0 ? 0 : 1;
}

// GDALResampleChunk_AverageOrRMS: destination pixel d covers the source interval [d r, (d + 1) r), r = in / out; the span runs
// from (int)(d r + 1e-8) to ceil((d + 1) r - 1e-8), at least one sample, clipped to the raster; the first and the last sample
// count with the fraction of them the interval covers.
void build_read_average_axis(uint64_t in, uint64_t out, ReadAverageAxisHost* a) {
    a->start.resize(out);
    a->end.resize(out);
    a->w_first.resize(out);
    a->w_last.resize(out);
    const double r = (double)in / (double)out;
    for (uint64_t d = 0; d < out; ++d) {
        const double lo = (double)d * r, hi = (double)(d + 1) * r;
        int s = (int)(lo + 1e-8);
        int e = (int)std::ceil(hi - 1e-8);
        if (e == s) ++e;
        if (e > (int)in) e = (int)in;
        if (s >= e) s = e - 1;
        double wf = 1.0 - (lo - (double)s), wl = 1.0 - ((double)e - hi);
        if (!(wf > 0.0) || wf > 1.0) wf = 1.0;
        if (!(wl > 0.0) || wl > 1.0) wl = 1.0;
        a->start[d] = s;
        a->end[d] = e;
        a->w_first[d] = wf;
        a->w_last[d] = wl;
    }
}

namespace {
double lanczos3_kernel(double x) {
    if (x == 0.0) return 1.0;
    if (x <= -3.0 || x >= 3.0) return 0.0;
    const double px = M_PI * x;
    return (std::sin(px) / px) * (std::sin(px / 3.0) / (px / 3.0));
}
} // namespace

// GDALResampleChunk_Convolution with the Lanczos kernel (a = 3): centre (d + 0.5) r, radius 3 r when shrinking, span
// [floor(c - R + 0.5), (int)(c + R + 0.5)) clipped to the raster, weights normalised by their sum.
void build_read_lanczos_axis(uint64_t in, uint64_t out, ReadLanczosAxisHost* a) {
    const double r = (double)in / (double)out;
    const double sw = r > 1.0 ? 1.0 / r : 1.0;
    const double R = 3.0 / sw;
    a->window = (int)std::ceil(2.0 * R) + 2;
    a->start.resize(out);
    a->count.resize(out);
    a->w.assign(out * (size_t)a->window, 0.0);
    for (uint64_t d = 0; d < out; ++d) {
        const double c = ((double)d + 0.5) * r;
        int s = (int)std::floor(c - R + 0.5), e = (int)(c + R + 0.5);
        if (s < 0) s = 0;
        if (e > (int)in) e = (int)in;
        if (e <= s) { s = std::min<int>(std::max<int>((int)c, 0), (int)in - 1); e = s + 1; }
        double sum = 0.0;
        double* w = a->w.data() + d * (size_t)a->window;
        for (int j = s; j < e; ++j) {
            w[j - s] = lanczos3_kernel(((double)j + 0.5 - c) * sw);
            sum += w[j - s];
        }
        if (sum != 0.0)
            for (int j = s; j < e; ++j) w[j - s] /= sum;
        a->start[d] = s;
        a->count[d] = e - s;
    }
}

} // namespace sarpro

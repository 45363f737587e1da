// oracle_read.cpp — CPU restatement of the downsample-on-read flow the CLI takes with --size
// (src/io/sentinel1.rs:1074-1109 -> src/io/gdal.rs:145-177: RasterBand::read_as::<f32> with ResampleAlg::Average / Lanczos).
//
// TEST INFRASTRUCTURE ONLY (see oracle.h). PARITY UNPINNED: the arithmetic lives in the system libgdal behind the `gdal`
// crate (Cargo.toml:34-35, library version unpinned), which is not under /root/reference and not on this box. Restated from
// the published algorithm of GDAL >= 3.3 RasterIO resampling (gcore/overview.cpp):
//   Average  GDALResampleChunk_AverageOrRMS: destination pixel d covers source [d*r, (d+1)*r) per axis (r = src/dst), first
//            index (int)(d*r + 1e-8), end index ceil((d+1)*r - 1e-8) (at least one pixel, clipped to the raster), the first
//            and last pixel of the span weighted by their covered fraction, weight = wy * wx, sum and weight sum in f64,
//            rows outer / columns inner, result (float)(sum / weight_sum).
//   Lanczos  GDALResampleChunk_Convolution, a = 3: separable, horizontal first into an f64 intermediate; centre
//            (d + 0.5) * r, radius 3 * r (downsampling), span [floor(c - R + 0.5), (int)(c + R + 0.5)) clipped to the raster,
//            weights lanczos3((j + 0.5 - c) / r) normalised by their sum, f64 accumulate, (float) at the end.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "oracle.h"

namespace {

struct AvgAxis {
    std::vector<int> start, end;
    std::vector<double> w_first, w_last;
};
AvgAxis average_axis(size_t in, size_t out) {
    AvgAxis a;
    a.start.resize(out); a.end.resize(out); a.w_first.resize(out); a.w_last.resize(out);
    const double r = (double)in / (double)out;
    for (size_t d = 0; d < out; ++d) {
        const double off = (double)d * r, off2 = (double)(d + 1) * r;
        int s = (int)(off + 1e-8);
        int e = (int)std::ceil(off2 - 1e-8);
        if (e == s) ++e;
        if (e > (int)in) e = (int)in;
        if (s >= e) s = e - 1;
        a.start[d] = s;
        a.end[d] = e;
        double wf = 1.0 - (off - (double)s), wl = 1.0 - ((double)e - off2);
        if (!(wf > 0.0) || wf > 1.0) wf = 1.0;
        if (!(wl > 0.0) || wl > 1.0) wl = 1.0;
        a.w_first[d] = wf;
        a.w_last[d] = wl;
    }
    return a;
}

double lanczos3(double x) {
    if (x == 0.0) return 1.0;
    if (x <= -3.0 || x >= 3.0) return 0.0;
    const double pix = M_PI * x;
    return (std::sin(pix) / pix) * (std::sin(pix / 3.0) / (pix / 3.0));
}
struct ConvAxis {
    std::vector<int> start, count;
    std::vector<double> w; // [out][window]
    int window = 0;
};
ConvAxis lanczos_axis(size_t in, size_t out) {
    ConvAxis a;
    const double r = (double)in / (double)out;
    const double sw = r > 1.0 ? 1.0 / r : 1.0; // weight-space scale (GDAL dfXScaleWeight)
    const double R = 3.0 / sw;
    a.window = (int)std::ceil(2.0 * R) + 2;
    a.start.resize(out); a.count.resize(out); a.w.assign(out * (size_t)a.window, 0.0);
    for (size_t d = 0; d < out; ++d) {
        const double c = ((double)d + 0.5) * r;
        int s = (int)std::floor(c - R + 0.5), e = (int)(c + R + 0.5);
        if (s < 0) s = 0;
        if (e > (int)in) e = (int)in;
        if (e <= s) { s = std::min<int>(std::max<int>((int)c, 0), (int)in - 1); e = s + 1; }
        double sum = 0.0;
        for (int j = s; j < e; ++j) {
            const double wv = lanczos3(((double)j + 0.5 - c) * sw);
            a.w[d * (size_t)a.window + (size_t)(j - s)] = wv;
            sum += wv;
        }
        if (sum != 0.0)
            for (int j = s; j < e; ++j) a.w[d * (size_t)a.window + (size_t)(j - s)] /= sum;
        a.start[d] = s;
        a.count[d] = e - s;
    }
    return a;
}

template <typename T>
void read_resampled(const T* src, size_t rows, size_t cols, size_t out_cols, size_t out_rows, int alg, float* out) {
    if (alg == 0) { // Average
        const AvgAxis ax = average_axis(cols, out_cols), ay = average_axis(rows, out_rows);
#pragma omp parallel for schedule(static)
        for (long long dy = 0; dy < (long long)out_rows; ++dy)
            for (size_t dx = 0; dx < out_cols; ++dx) {
                double total = 0.0, wsum = 0.0;
                for (int y = ay.start[dy]; y < ay.end[dy]; ++y) {
                    const double wy = y == ay.start[dy] ? ay.w_first[dy] : (y + 1 == ay.end[dy] ? ay.w_last[dy] : 1.0);
                    for (int x = ax.start[dx]; x < ax.end[dx]; ++x) {
                        const double wx = x == ax.start[dx] ? ax.w_first[dx] : (x + 1 == ax.end[dx] ? ax.w_last[dx] : 1.0);
                        const double w = wy * wx;
                        total += (double)src[(size_t)y * cols + x] * w;
                        wsum += w;
                    }
                }
                out[(size_t)dy * out_cols + dx] = (float)(total / wsum);
            }
        return;
    }
    const ConvAxis ax = lanczos_axis(cols, out_cols), ay = lanczos_axis(rows, out_rows);
    std::vector<double> tmp(rows * out_cols);
#pragma omp parallel for schedule(static)
    for (long long y = 0; y < (long long)rows; ++y)
        for (size_t dx = 0; dx < out_cols; ++dx) {
            double v = 0.0;
            for (int k = 0; k < ax.count[dx]; ++k) v += (double)src[(size_t)y * cols + ax.start[dx] + k] * ax.w[dx * (size_t)ax.window + k];
            tmp[(size_t)y * out_cols + dx] = v;
        }
#pragma omp parallel for schedule(static)
    for (long long dy = 0; dy < (long long)out_rows; ++dy)
        for (size_t dx = 0; dx < out_cols; ++dx) {
            double v = 0.0;
            for (int k = 0; k < ay.count[dy]; ++k) v += tmp[(size_t)(ay.start[dy] + k) * out_cols + dx] * ay.w[dy * (size_t)ay.window + k];
            out[(size_t)dy * out_cols + dx] = (float)v;
        }
}

} // namespace

extern "C" {

// sentinel1.rs:1083-1102: output shape for a long-side target (no upscale) and the resampler the reader picks
void oracle_read_dims_for_target(size_t cols, size_t rows, size_t target, size_t* out_cols, size_t* out_rows, int* alg) {
    const size_t long_side = std::max(cols, rows);
    const double scale = std::min((double)target / (double)long_side, 1.0);
    *out_cols = (size_t)std::max(std::round((double)cols * scale), 1.0);
    *out_rows = (size_t)std::max(std::round((double)rows * scale), 1.0);
    const double reduction = std::max((double)long_side / (double)target, 1.0);
    *alg = reduction >= 4.0 ? 0 : 1;
}

// gdal.rs:145-177 read_band_resampled on a u16 (GRD DN) or f32 raster
void oracle_read_band_resampled_u16(const uint16_t* src, size_t rows, size_t cols, size_t out_cols, size_t out_rows, int alg, float* out) {
    read_resampled(src, rows, cols, out_cols, out_rows, alg, out);
}
void oracle_read_band_resampled_f32(const float* src, size_t rows, size_t cols, size_t out_cols, size_t out_rows, int alg, float* out) {
    read_resampled(src, rows, cols, out_cols, out_rows, alg, out);
}

} // extern "C"

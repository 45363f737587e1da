// oracle_resize.cpp — CPU ORACLE (test infrastructure only; see oracle.h).
//
// resize.rs / padding.rs restated, plus the third-party arithmetic they call:
//   fast_image_resize = "5.2.1" (Cargo.toml:33; Cargo.lock is not committed, so the patch
//   version is unpinned).  The crate is NOT under /root/reference and cannot be fetched; what
//   follows restates its published algorithm (a Rust port of Pillow-SIMD's ImagingResample):
//     * ResizeAlg::Convolution(FilterType::Lanczos3), call sites resize.rs:39-50 (U8), :62-81 (U16)
//     * separable: horizontal pass first into a (dst_w x used_src_rows) temporary of the SAME
//       pixel type (rounded + clamped), then the vertical pass
//     * coefficient windows: scale = in/out, filter_scale = max(scale,1), radius = 3*filter_scale,
//       in_center = (x+0.5)*scale, x_min = floor(in_center-radius) clamped to 0,
//       x_max = ceil(in_center+radius) clamped to in_size, w = lanczos3((x-(in_center-0.5))/filter_scale),
//       leading/trailing zero weights trimmed from the bound, weights normalised to sum 1 (f64)
//     * U8 : coefficients quantised to i16 at an adaptive precision p (largest p < 22 reached
//            before round(max_w * 2^(p+1)) >= 2^15), i32 accumulator seeded with 1<<(p-1),
//            result = clamp(acc >> p, 0, 255)
//     * U16: coefficients quantised to i32 (p < 46, bound 2^31), i64 accumulator seeded with
//            1<<(p-1), result = clamp(acc >> p, 0, 65535)
// PARITY UNPINNED: no reference test pins this stage. tests/ cross-check the u8
// path against Pillow's LANCZOS (same window formulation, different fixed-point precision) as a
// sanity bound only.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

int g_resize_threads = 1;

struct Bound { uint32_t start, size; };
struct Coefficients {
    std::vector<double> values;
    size_t window_size = 0;
    std::vector<Bound> bounds;
};

inline double sinc_filter(double x) {
    if (x == 0.0) return 1.0;
    x *= M_PI;
    return std::sin(x) / x;
}
inline double lanczos3(double x) {
    if (x >= -3.0 && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3.0);
    return 0.0;
}

Coefficients precompute_coefficients(uint32_t in_size, double in0, double in1, uint32_t out_size) {
    Coefficients co;
    if (in_size == 0 || out_size == 0) return co;
    const double scale = (in1 - in0) / (double)out_size;
    if (scale <= 0.0) return co;
    const double filter_support = 3.0;
    const double filter_scale = std::fmax(scale, 1.0);
    const double filter_radius = filter_support * filter_scale;
    const size_t window_size = (size_t)std::ceil(filter_radius) * 2 + 1;
    const double recip_filter_scale = 1.0 / filter_scale;
    co.window_size = window_size;
    co.values.reserve(window_size * out_size);
    co.bounds.reserve(out_size);
    for (uint32_t out_x = 0; out_x < out_size; ++out_x) {
        const double in_center = in0 + ((double)out_x + 0.5) * scale;
        const uint32_t x_min = (uint32_t)std::fmax(std::floor(in_center - filter_radius), 0.0);
        const uint32_t x_max = (uint32_t)std::fmin(std::ceil(in_center + filter_radius), (double)in_size);
        const size_t cur_index = co.values.size();
        double ww = 0.0;
        const double center = in_center - 0.5;
        uint32_t bound_start = x_min, bound_end = x_max;
        for (uint32_t x = x_min; x < x_max; ++x) {
            const double w = lanczos3(((double)x - center) * recip_filter_scale);
            if (x == bound_start && w == 0.0) {
                bound_start += 1; // skip zero coefficients at the start of the bound
            } else {
                co.values.push_back(w);
                ww += w;
            }
        }
        for (size_t k = co.values.size(); k > cur_index; --k) {
            if (bound_end <= bound_start || co.values[k - 1] != 0.0) break;
            bound_end -= 1; // skip zero coefficients at the end of the bound
        }
        if (ww != 0.0)
            for (size_t k = cur_index; k < co.values.size(); ++k) co.values[k] /= ww;
        co.values.resize(cur_index + window_size, 0.0);
        co.bounds.push_back(Bound{bound_start, bound_end - bound_start});
    }
    return co;
}

double max_weight(const Coefficients& co) {
    double m = 0.0;
    bool first = true;
    for (double v : co.values) {
        if (first || v > m) { m = v; first = false; }
    }
    return m;
}

struct Norm16 { // U8 pixels: i16 coefficients
    int precision = 0;
    std::vector<int16_t> k; // window_size per output
    size_t window = 0;
    std::vector<Bound> bounds;
};
Norm16 make_norm16(const Coefficients& co) {
    Norm16 n;
    n.window = co.window_size;
    n.bounds = co.bounds;
    const double mw = max_weight(co);
    int precision = 0;
    for (int cur = 0; cur < 22; ++cur) {
        precision = cur;
        const double nv = std::round(mw * (double)(1 << (precision + 1)));
        if (nv >= (double)(1 << 15)) break;
    }
    n.precision = precision;
    const double scale = (double)(1 << precision);
    n.k.resize(co.values.size());
    for (size_t i = 0; i < co.values.size(); ++i) n.k[i] = (int16_t)std::round(co.values[i] * scale);
    return n;
}
struct Norm32 { // U16 pixels: i32 coefficients
    int precision = 0;
    std::vector<int32_t> k;
    size_t window = 0;
    std::vector<Bound> bounds;
};
Norm32 make_norm32(const Coefficients& co) {
    Norm32 n;
    n.window = co.window_size;
    n.bounds = co.bounds;
    const double mw = max_weight(co);
    int precision = 0;
    for (int cur = 0; cur < 46; ++cur) {
        precision = cur;
        const double nv = std::round(mw * (double)((int64_t)1 << (precision + 1)));
        if (nv >= (double)((int64_t)1 << 31)) break;
    }
    n.precision = precision;
    const double scale = (double)((int64_t)1 << precision);
    n.k.resize(co.values.size());
    for (size_t i = 0; i < co.values.size(); ++i) n.k[i] = (int32_t)std::round(co.values[i] * scale);
    return n;
}

inline uint8_t clip8(int32_t acc, int p) {
    const int32_t v = acc >> p;
    return (uint8_t)std::min(std::max(v, 0), 255);
}
inline uint16_t clip16(int64_t acc, int p) {
    const int64_t v = acc >> p;
    return (uint16_t)std::min<int64_t>(std::max<int64_t>(v, 0), 65535);
}

template <typename T> struct Traits;
template <> struct Traits<uint8_t> {
    using Norm = Norm16; using Acc = int32_t;
    static Norm make(const Coefficients& c) { return make_norm16(c); }
    static uint8_t clip(Acc a, int p) { return clip8(a, p); }
};
template <> struct Traits<uint16_t> {
    using Norm = Norm32; using Acc = int64_t;
    static Norm make(const Coefficients& c) { return make_norm32(c); }
    static uint16_t clip(Acc a, int p) { return clip16(a, p); }
};

template <typename T>
void horiz_pass(const T* src, size_t src_w, size_t y_first, size_t n_rows, T* dst, size_t dst_w,
                const typename Traits<T>::Norm& nrm) {
    using Acc = typename Traits<T>::Acc;
    const int p = nrm.precision;
    const Acc initial = p > 0 ? (Acc)1 << (p - 1) : 0;
#pragma omp parallel for num_threads(g_resize_threads) schedule(static) if (g_resize_threads > 1)
    for (long long y = 0; y < (long long)n_rows; ++y) {
        const T* srow = src + (y_first + (size_t)y) * src_w;
        T* drow = dst + (size_t)y * dst_w;
        for (size_t x = 0; x < dst_w; ++x) {
            const Bound b = nrm.bounds[x];
            const auto* k = &nrm.k[x * nrm.window];
            Acc ss = initial;
            for (uint32_t i = 0; i < b.size; ++i) ss += (Acc)srow[b.start + i] * (Acc)k[i];
            drow[x] = Traits<T>::clip(ss, p);
        }
    }
}
template <typename T>
void vert_pass(const T* src, size_t w, T* dst, size_t dst_h, const typename Traits<T>::Norm& nrm,
               size_t y_shift) {
    using Acc = typename Traits<T>::Acc;
    const int p = nrm.precision;
    const Acc initial = p > 0 ? (Acc)1 << (p - 1) : 0;
#pragma omp parallel for num_threads(g_resize_threads) schedule(static) if (g_resize_threads > 1)
    for (long long y = 0; y < (long long)dst_h; ++y) {
        const Bound b = nrm.bounds[(size_t)y];
        const auto* k = &nrm.k[(size_t)y * nrm.window];
        T* drow = dst + (size_t)y * w;
        for (size_t x = 0; x < w; ++x) {
            Acc ss = initial;
            for (uint32_t i = 0; i < b.size; ++i) ss += (Acc)src[(b.start - y_shift + i) * w + x] * (Acc)k[i];
            drow[x] = Traits<T>::clip(ss, p);
        }
    }
}

// Resizer::resize with ResizeAlg::Convolution: horizontal into temp, then vertical.
template <typename T>
int resize_lanczos3(const T* data, size_t cols, size_t rows, size_t tcols, size_t trows, T* out) {
    if (tcols == 0 || trows == 0) return 0; // empty destination: nothing to do
    if (cols == 0 || rows == 0) return -1;  // crate rejects empty source with non-empty destination
    const bool need_h = tcols != cols;
    const bool need_v = trows != rows;
    if (!need_h && !need_v) { std::memcpy(out, data, cols * rows * sizeof(T)); return 0; }
    if (need_h && need_v) {
        const Coefficients hc = precompute_coefficients((uint32_t)cols, 0.0, (double)cols, (uint32_t)tcols);
        const Coefficients vc = precompute_coefficients((uint32_t)rows, 0.0, (double)rows, (uint32_t)trows);
        const auto hn = Traits<T>::make(hc);
        const auto vn = Traits<T>::make(vc);
        const size_t y_first = vc.bounds.front().start;
        const size_t y_last = vc.bounds.back().start + vc.bounds.back().size;
        const size_t temp_h = y_last - y_first;
        std::vector<T> temp(tcols * temp_h);
        horiz_pass<T>(data, cols, y_first, temp_h, temp.data(), tcols, hn);
        vert_pass<T>(temp.data(), tcols, out, trows, vn, y_first);
        return 0;
    }
    if (need_h) {
        const Coefficients hc = precompute_coefficients((uint32_t)cols, 0.0, (double)cols, (uint32_t)tcols);
        const auto hn = Traits<T>::make(hc);
        horiz_pass<T>(data, cols, 0, rows, out, tcols, hn);
        return 0;
    }
    const Coefficients vc = precompute_coefficients((uint32_t)rows, 0.0, (double)rows, (uint32_t)trows);
    const auto vn = Traits<T>::make(vc);
    vert_pass<T>(data, cols, out, trows, vn, 0);
    return 0;
}

template <typename T>
void pad_to_square(const T* src, size_t cols, size_t rows, T* dst) {
    // padding.rs:12-14, 24-33 / 37-46
    const size_t max_dim = std::max(cols, rows);
    const size_t pad_cols = (max_dim - cols) / 2;
    const size_t pad_rows = (max_dim - rows) / 2;
    std::fill(dst, dst + max_dim * max_dim, (T)0);
    for (size_t row = 0; row < rows; ++row)
        std::memcpy(dst + (row + pad_rows) * max_dim + pad_cols, src + row * cols, cols * sizeof(T));
}

} // namespace

extern "C" {

void oracle_set_resize_threads(int n) { g_resize_threads = n < 1 ? 1 : n; }

// resize.rs:6-30
void oracle_calculate_resize_dimensions(size_t cols, size_t rows, size_t target, size_t* new_cols,
                                        size_t* new_rows) {
    const size_t short_side = std::min(rows, cols);
    const size_t long_side = std::max(rows, cols);
    if (target > long_side) { *new_cols = cols; *new_rows = rows; return; }
    const double scale_factor = (double)target / (double)long_side;
    const double r = std::round((double)short_side * scale_factor);
    const size_t new_short = r != r || r <= 0.0 ? 0 : (size_t)r;
    if (cols > rows) { *new_cols = target; *new_rows = new_short; }
    else { *new_cols = new_short; *new_rows = target; }
}

int oracle_resize_u8_image(const uint8_t* data, size_t cols, size_t rows, size_t tcols, size_t trows,
                           uint8_t* out) {
    return resize_lanczos3<uint8_t>(data, cols, rows, tcols, trows, out);
}
int oracle_resize_u16_image(const uint16_t* data, size_t cols, size_t rows, size_t tcols, size_t trows,
                            uint16_t* out) {
    return resize_lanczos3<uint16_t>(data, cols, rows, tcols, trows, out);
}

// padding.rs:5-49
int oracle_add_padding_to_square(const uint8_t* u8_data, const uint16_t* u16_data, size_t cols, size_t rows,
                                 int bit_depth, uint8_t* out_u8, uint16_t* out_u16) {
    if (bit_depth == ORACLE_U8) { pad_to_square<uint8_t>(u8_data, cols, rows, out_u8); return 0; }
    if (!u16_data) return -1; // "U16 data required for U16 bit depth" (padding.rs:36)
    pad_to_square<uint16_t>(u16_data, cols, rows, out_u16);
    return 0;
}

void oracle_resize_output_dims(size_t cols, size_t rows, int has_target, size_t target, int pad,
                               size_t* out_cols, size_t* out_rows) {
    size_t c = cols, r = rows;
    if (has_target && std::max(cols, rows) != target) oracle_calculate_resize_dimensions(cols, rows, target, &c, &r);
    if (pad) { const size_t m = std::max(c, r); c = m; r = m; }
    *out_cols = c;
    *out_rows = r;
}

// resize.rs:91-236
int oracle_resize_image_data_with_meta(const uint8_t* u8_data, const uint16_t* u16_data, size_t cols,
                                       size_t rows, int has_target, size_t target, int bit_depth, int pad,
                                       uint8_t* out_u8, uint16_t* out_u16, oracle_resize_meta* meta) {
    oracle_resize_meta m{};
    m.scale_x = 1.0;
    m.scale_y = 1.0;
    size_t cur_cols = cols, cur_rows = rows;
    std::vector<uint8_t> r8;
    std::vector<uint16_t> r16;
    const uint8_t* s8 = u8_data;
    const uint16_t* s16 = u16_data;
    if (has_target && std::max(cols, rows) != target) { // :112-116, 147-170
        size_t nc, nr;
        oracle_calculate_resize_dimensions(cols, rows, target, &nc, &nr);
        if (bit_depth == ORACLE_U8) {
            r8.resize(nc * nr);
            if (resize_lanczos3<uint8_t>(u8_data, cols, rows, nc, nr, r8.data())) return -2;
            s8 = r8.data();
        } else {
            if (!u16_data) return -1; // resize.rs:162
            r16.resize(nc * nr);
            if (resize_lanczos3<uint16_t>(u16_data, cols, rows, nc, nr, r16.data())) return -2;
            s16 = r16.data();
        }
        m.scale_x = (double)nc / (double)cols;
        m.scale_y = (double)nr / (double)rows;
        cur_cols = nc;
        cur_rows = nr;
    }
    if (pad) { // :119-132, 172-185, 199-207
        if (oracle_add_padding_to_square(s8, bit_depth == ORACLE_U16 ? s16 : nullptr, cur_cols, cur_rows, bit_depth,
                                         out_u8, out_u16))
            return -1;
        const size_t final_dim = std::max(cur_cols, cur_rows);
        m.pad_left = (final_dim - cur_cols) / 2;
        m.pad_top = (final_dim - cur_rows) / 2;
        m.cols = final_dim;
        m.rows = final_dim;
    } else {
        if (bit_depth == ORACLE_U8) std::memcpy(out_u8, s8, cur_cols * cur_rows);
        else {
            if (!s16) return -1; // resize.rs:221
            std::memcpy(out_u16, s16, cur_cols * cur_rows * sizeof(uint16_t));
        }
        m.cols = cur_cols;
        m.rows = cur_rows;
    }
    if (meta) *meta = m;
    return 0;
}

} // extern "C"

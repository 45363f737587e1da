// oracle.cpp — CPU ORACLE (test infrastructure only; see oracle.h header for the rules).
//
// Serial restatement of the reference's Rust hot path, one function per reference function,
// each citing the file:line it follows (paths relative to the reference root).
// PARITY UNPINNED: the reference has no tests/golden vectors.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math (Rust never contracts to FMA).
// Rust -> C++ semantics used throughout:
//   `x as uN/usize/isize` from float : truncate toward zero, saturating, NaN -> 0   (rs_cast_*)
//   f64::round / f32::round          : half away from zero                           (round/roundf)
//   f64::max / f64::min              : NaN-ignoring                                  (fmax/fmin)
//   .clamp(a, b)                     : NaN propagates                                (rs_clamp)
//   f64::log10 / powf, f32::powf     : libm log10 / pow / powf
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

inline double rs_clamp(double v, double lo, double hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v; // NaN stays NaN
}
inline float rs_clampf(float v, float lo, float hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
inline uint64_t rs_cast_u64(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}
inline int64_t rs_cast_i64(double x) {
    if (!(x == x)) return 0;
    if (x >= 9223372036854775808.0) return INT64_MAX;
    if (x <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)x;
}
inline uint32_t rs_cast_u32(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 4294967295.0) return UINT32_MAX;
    return (uint32_t)x;
}
inline uint16_t rs_cast_u16(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 65535.0) return 65535;
    return (uint16_t)x;
}
inline uint8_t rs_cast_u8f(float x) {
    if (!(x == x)) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 255.0f) return 255;
    return (uint8_t)x;
}
inline uint8_t rs_cast_u8(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 255.0) return 255;
    return (uint8_t)x;
}

// ---------------------------------------------------------------------------------------------
// autoscale.rs:35-160  compute_histogram_stats
// ---------------------------------------------------------------------------------------------
void compute_histogram_stats(const double* db, const uint8_t* mask, size_t rows, size_t cols,
                             oracle_stats* s, uint64_t* hist_out) {
    const size_t n_px = rows * cols;
    // First pass: min/max + Welford mean/std (autoscale.rs:37-55)
    uint64_t count = 0;
    double min_db = std::numeric_limits<double>::infinity();
    double max_db = -std::numeric_limits<double>::infinity();
    double mean = 0.0, m2 = 0.0;
    for (size_t i = 0; i < n_px; ++i) {
        if (mask[i]) {
            const double v = db[i];
            count += 1;
            if (v < min_db) min_db = v;
            if (v > max_db) max_db = v;
            const double delta = v - mean;
            mean += delta / (double)count;
            const double delta2 = v - mean;
            m2 += delta * delta2;
        }
    }
    std::memset(s, 0, sizeof(*s));
    if (hist_out) std::memset(hist_out, 0, 4096 * sizeof(uint64_t));
    if (count == 0) return; // autoscale.rs:57-76 (all zeros)

    const double std_db = count > 1 ? std::sqrt(m2 / (double)count) : 0.0; // :78
    s->valid_count = count;
    s->min_db = min_db;
    s->max_db = max_db;
    s->mean_db = mean;
    s->std_db = std_db;

    // Degenerate case (autoscale.rs:81-100)
    if (std::fabs(max_db - min_db) < std::numeric_limits<double>::epsilon()) {
        s->median_db = min_db;
        s->p01 = s->p02 = s->p05 = s->p10 = s->p25 = min_db;
        s->p75 = s->p90 = s->p95 = s->p98 = s->p99 = max_db;
        return;
    }

    // Second pass: histogram over [min,max] (autoscale.rs:103-117)
    const size_t NUM_BINS = 4096;
    std::vector<uint64_t> hist(NUM_BINS, 0);
    const double span = max_db - min_db;
    const double inv_span = 1.0 / span;
    for (size_t i = 0; i < n_px; ++i) {
        if (!mask[i]) continue;
        const double t = rs_clamp((db[i] - min_db) * inv_span, 0.0, 1.0);
        uint64_t idx = rs_cast_u64(t * (double)NUM_BINS);
        if (idx >= NUM_BINS) idx = NUM_BINS - 1;
        hist[idx] += 1;
    }
    if (hist_out) std::memcpy(hist_out, hist.data(), NUM_BINS * sizeof(uint64_t));

    // autoscale.rs:120-140 estimate_percentile
    auto estimate_percentile = [&](double p) -> double {
        const uint64_t n = count;
        uint64_t target = rs_cast_u64(std::floor(p * (double)n));
        if (target >= n) target = n - 1;
        uint64_t cumsum = 0;
        for (size_t b = 0; b < NUM_BINS; ++b) {
            const uint64_t h = hist[b];
            const uint64_t next = cumsum + h;
            if (target < next) {
                const uint64_t within = target >= cumsum ? target - cumsum : 0; // saturating_sub
                const double frac = h > 0 ? (double)within / (double)h : 0.0;
                const double bin_width = span / (double)NUM_BINS;
                const double bin_start = min_db + (double)b * bin_width;
                return bin_start + frac * bin_width;
            }
            cumsum = next;
        }
        return max_db;
    };
    // Evaluation order of the struct literal (autoscale.rs:142-159) has no side effects.
    s->median_db = estimate_percentile(0.5);
    s->p01 = estimate_percentile(0.01);
    s->p02 = estimate_percentile(0.02);
    s->p05 = estimate_percentile(0.05);
    s->p10 = estimate_percentile(0.10);
    s->p25 = estimate_percentile(0.25);
    s->p75 = estimate_percentile(0.75);
    s->p90 = estimate_percentile(0.90);
    s->p95 = estimate_percentile(0.95);
    s->p98 = estimate_percentile(0.98);
    s->p99 = estimate_percentile(0.99);
}

inline bool approx_eq(double a, double b) { return std::fabs(a - b) < 1e-9; } // autoscale.rs:26-29

// ---------------------------------------------------------------------------------------------
// autoscale.rs:220-345  clahe_equalize_normalized
// ---------------------------------------------------------------------------------------------
void clahe_equalize_normalized(const double* norm, const uint8_t* mask, size_t rows, size_t cols,
                               size_t tiles_x, size_t tiles_y, double clip_limit, size_t num_bins,
                               double* out, double* cdfs_out) {
    if (rows == 0 || cols == 0 || tiles_x == 0 || tiles_y == 0 || num_bins < 2) { // :231-233
        std::memcpy(out, norm, rows * cols * sizeof(double));
        return;
    }
    const size_t tile_h = (rows + tiles_y - 1) / tiles_y;
    const size_t tile_w = (cols + tiles_x - 1) / tiles_x;
    std::vector<std::vector<double>> cdfs(tiles_x * tiles_y, std::vector<double>(num_bins, 0.0));

    for (size_t ty = 0; ty < tiles_y; ++ty) {
        // NOTE: for rows not reaching this tile row, Rust's `r0..r1` with r0 > r1 would make
        // `r1 - r0` underflow (panic in debug, wrap in release). tiles=8 with rows >= 8 never
        // gets there except for very small images; we mirror with saturating arithmetic and the
        // tests keep rows/cols >= tiles * 1 so that r0 <= rows always holds (see tests).
        const size_t r0 = ty * tile_h;
        const size_t r1 = std::min((ty + 1) * tile_h, rows);
        const size_t tile_rows = r1 >= r0 ? r1 - r0 : 0;
        for (size_t tx = 0; tx < tiles_x; ++tx) {
            const size_t c0 = tx * tile_w;
            const size_t c1 = std::min((tx + 1) * tile_w, cols);
            const size_t tile_cols = c1 >= c0 ? c1 - c0 : 0;

            std::vector<uint32_t> hist(num_bins, 0);
            for (size_t r = r0; r < r1; ++r) {
                for (size_t c = c0; c < c1; ++c) {
                    if (mask[r * cols + c]) {
                        const double v = rs_clamp(norm[r * cols + c], 0.0, 1.0);
                        int64_t bin = rs_cast_i64(std::round(v * ((double)num_bins - 1.0)));
                        if (bin < 0) bin = 0;
                        if ((size_t)bin >= num_bins) bin = (int64_t)(num_bins - 1);
                        hist[(size_t)bin] += 1;
                    }
                }
            }
            // Clip histogram (autoscale.rs:271-280)
            const double avg = (double)(tile_rows * tile_cols) / (double)num_bins;
            const double clip_threshold = std::fmax(clip_limit * avg, 1.0);
            double excess = 0.0;
            for (auto& h : hist) {
                if ((double)h > clip_threshold) {
                    excess += (double)h - clip_threshold;
                    h = rs_cast_u32(clip_threshold);
                }
            }
            // Redistribute (autoscale.rs:281-292)
            const double add_per_bin = std::floor(excess / (double)num_bins);
            uint64_t remainder = rs_cast_u64(std::round(excess - add_per_bin * (double)num_bins));
            for (auto& h : hist) h = rs_cast_u32((double)h + add_per_bin);
            size_t b = 0;
            while (remainder > 0) {
                hist[b] += 1;
                b = (b + 1) % num_bins;
                remainder -= 1;
            }
            // CDF (autoscale.rs:294-302)
            double total = 0.0;
            for (auto h : hist) total += (double)h;
            total = std::fmax(total, 1.0);
            std::vector<double>& cdf = cdfs[ty * tiles_x + tx];
            double acc = 0.0;
            for (size_t i = 0; i < num_bins; ++i) {
                acc += (double)hist[i];
                cdf[i] = rs_clamp(acc / total, 0.0, 1.0);
            }
        }
    }
    if (cdfs_out)
        for (size_t t = 0; t < tiles_x * tiles_y; ++t)
            std::memcpy(cdfs_out + t * num_bins, cdfs[t].data(), num_bins * sizeof(double));

    // autoscale.rs:307-330 sample_cdf
    auto sample_cdf = [&](size_t r, size_t c, double val) -> double {
        const double rf = (double)r / (double)tile_h - 0.5;
        const double cf = (double)c / (double)tile_w - 0.5;
        const int64_t ty = rs_cast_i64(std::fmax(std::floor(rf), 0.0));
        const int64_t tx = rs_cast_i64(std::fmax(std::floor(cf), 0.0));
        const double dy = rf - (double)ty;
        const double dx = cf - (double)tx;
        const int64_t ny = (int64_t)tiles_y - 1, nx = (int64_t)tiles_x - 1;
        const size_t ty0 = (size_t)std::min(std::max(ty, (int64_t)0), ny);
        const size_t tx0 = (size_t)std::min(std::max(tx, (int64_t)0), nx);
        const size_t ty1 = (size_t)std::min(std::max(ty + 1, (int64_t)0), ny);
        const size_t tx1 = (size_t)std::min(std::max(tx + 1, (int64_t)0), nx);
        const size_t bin_pos = (size_t)rs_cast_u64(std::round(rs_clamp(val, 0.0, 1.0) * ((double)num_bins - 1.0)));
        const double cdf00 = cdfs[ty0 * tiles_x + tx0][bin_pos];
        const double cdf01 = cdfs[ty0 * tiles_x + tx1][bin_pos];
        const double cdf10 = cdfs[ty1 * tiles_x + tx0][bin_pos];
        const double cdf11 = cdfs[ty1 * tiles_x + tx1][bin_pos];
        const double top = cdf00 * (1.0 - dx) + cdf01 * dx;
        const double bottom = cdf10 * (1.0 - dx) + cdf11 * dx;
        return top * (1.0 - dy) + bottom * dy;
    };
    for (size_t r = 0; r < rows; ++r)
        for (size_t c = 0; c < cols; ++c)
            out[r * cols + c] = mask[r * cols + c] ? sample_cdf(r, c, norm[r * cols + c]) : 0.0;
}

// autoscale.rs:348-364
void scale_u16_to_u8(const uint16_t* data, size_t n, uint8_t* out) {
    if (n == 0) return;
    uint16_t mn = data[0], mx = data[0];
    for (size_t i = 1; i < n; ++i) {
        if (data[i] < mn) mn = data[i];
        if (data[i] > mx) mx = data[i];
    }
    const float fmin_ = (float)mn, fmax_ = (float)mx;
    const float scale = fmax_ > fmin_ ? 255.0f / (fmax_ - fmin_) : 1.0f;
    for (size_t i = 0; i < n; ++i) {
        const float val = roundf(((float)data[i] - fmin_) * scale);
        out[i] = rs_cast_u8f(rs_clampf(val, 0.0f, 255.0f));
    }
}

inline uint16_t quantize(double v, double low, double high, double range, double gamma, double max_val) {
    // autoscale.rs:440-442 / 649-651
    const double clipped = std::fmin(std::fmax(v, low), high);
    const double normalized = std::pow((clipped - low) / range, gamma);
    return rs_cast_u16(rs_clamp(normalized * max_val, 0.0, max_val));
}

// autoscale.rs:368-448
void autoscale_db_image(const double* db, const uint8_t* mask, size_t rows, size_t cols, int bit_depth,
                        uint16_t* out, oracle_stats* stats_out) {
    oracle_stats st;
    compute_histogram_stats(db, mask, rows, cols, &st, nullptr);
    const size_t n = rows * cols;
    if (st.valid_count == 0) {
        std::fill(out, out + n, (uint16_t)0);
        if (stats_out) *stats_out = st;
        return;
    }
    const double min_db = st.min_db, max_db = st.max_db, median_db = st.median_db;
    const double p02 = st.p02, p25 = st.p25, p75 = st.p75, p98 = st.p98;
    const double max_val = bit_depth == ORACLE_U8 ? 255.0 : 65535.0;
    const double dynamic_range = max_db - min_db;
    const double iqr = p75 - p25;
    double low_clip, high_clip, gamma;
    if (dynamic_range < 15.0) { // :404-408
        const double range = std::fmax(20.0, dynamic_range * 0.8);
        low_clip = median_db - range / 2.0;
        high_clip = median_db + range / 2.0;
        gamma = 1.1;
    } else if (iqr < 5.0) { // :409-413
        const double outlier_factor = 2.5;
        low_clip = p25 - outlier_factor * iqr;
        high_clip = p75 + outlier_factor * iqr;
        gamma = 1.0;
    } else if (dynamic_range > 40.0) { // :414-419
        low_clip = std::fmax(p02, min_db + 0.02 * dynamic_range);
        high_clip = std::fmin(p98, max_db - 0.02 * dynamic_range);
        gamma = 0.9;
    } else { // :420-424
        low_clip = p02;
        high_clip = p98;
        gamma = 1.0;
    }
    low_clip = std::fmax(low_clip, min_db);   // :427
    high_clip = std::fmin(high_clip, max_db); // :428
    const double range = std::fmax(high_clip - low_clip, 1.0);
    st.low_clip = low_clip;
    st.high_clip = high_clip;
    st.gamma = gamma;
    if (stats_out) *stats_out = st;
    for (size_t i = 0; i < n; ++i)
        out[i] = mask[i] ? quantize(db[i], low_clip, high_clip, range, gamma, max_val) : (uint16_t)0;
}

// autoscale.rs:452-659 (use_local_enhancement is always false: :498,537,542,547,552,556,560)
void autoscale_db_image_advanced(const double* db, const uint8_t* mask, size_t rows, size_t cols,
                                 int bit_depth, int strategy, uint16_t* out, oracle_stats* stats_out) {
    const double max_val = bit_depth == ORACLE_U8 ? 255.0 : 65535.0;
    oracle_stats st;
    compute_histogram_stats(db, mask, rows, cols, &st, nullptr);
    const size_t n = rows * cols;
    if (st.valid_count == 0) {
        std::fill(out, out + n, (uint16_t)0);
        if (stats_out) *stats_out = st;
        return;
    }
    const double min_db = st.min_db, max_db = st.max_db, mean_db = st.mean_db, median_db = st.median_db;
    const double std_db = st.std_db, p01 = st.p01, p05 = st.p05, p25 = st.p25, p75 = st.p75, p95 = st.p95,
                 p99 = st.p99;
    const double iqr = p75 - p25;
    double low_clip, high_clip, gamma;
    switch (strategy) {
    case ORACLE_STRATEGY_ROBUST: { // :492-499
        const double outlier_threshold = 2.5 * iqr;
        low_clip = std::fmax(std::fmax(p25 - outlier_threshold, p01), min_db);
        high_clip = std::fmin(std::fmin(p75 + outlier_threshold, p99), max_db);
        gamma = 1.0;
        break;
    }
    case ORACLE_STRATEGY_ADAPTIVE: { // :500-538
        const double skew_factor = (mean_db - median_db) / std::fmax(std::fabs(std_db), 1.0);
        const double tail_heaviness = (p99 - p95) / std::fmax(p95 - p75, 1.0);
        double low_pct, high_pct, gamma_adj;
        if (std::fabs(skew_factor) > 0.5) {
            if (skew_factor > 0.0) { low_pct = 0.02; high_pct = 0.98; gamma_adj = 0.9; }
            else { low_pct = 0.05; high_pct = 0.95; gamma_adj = 1.1; }
        } else if (tail_heaviness > 2.0) { low_pct = 0.10; high_pct = 0.90; gamma_adj = 0.8; }
        else { low_pct = 0.05; high_pct = 0.95; gamma_adj = 1.0; }
        double low, high;
        if (approx_eq(low_pct, 0.10)) low = st.p10;
        else if (approx_eq(low_pct, 0.02)) low = st.p02;
        else if (approx_eq(low_pct, 0.05)) low = st.p05;
        else if (approx_eq(low_pct, 0.25)) low = st.p25;
        else if (approx_eq(low_pct, 0.75)) low = st.p75;
        else if (approx_eq(low_pct, 0.95)) low = st.p95;
        else if (approx_eq(low_pct, 0.99)) low = st.p99;
        else low = st.p05;
        if (approx_eq(high_pct, 0.90)) high = st.p90;
        else if (approx_eq(high_pct, 0.98)) high = st.p98;
        else if (approx_eq(high_pct, 0.95)) high = st.p95;
        else if (approx_eq(high_pct, 0.75)) high = st.p75;
        else if (approx_eq(high_pct, 0.99)) high = st.p99;
        else high = st.p95;
        low_clip = low; high_clip = high; gamma = gamma_adj;
        break;
    }
    case ORACLE_STRATEGY_EQUALIZED: // :539-543
    case ORACLE_STRATEGY_CLAHE:     // :544-548
        low_clip = p01; high_clip = p99; gamma = 1.0; break;
    case ORACLE_STRATEGY_TAMED:     // :549-553
        low_clip = p25; high_clip = p99; gamma = 1.0; break;
    case ORACLE_STRATEGY_STANDARD:  // :554-557 (unreachable through pipeline.rs:50-52)
    case ORACLE_STRATEGY_DEFAULT:   // :558-561
    default:
        low_clip = p05; high_clip = p95; gamma = 1.0; break;
    }
    const double range = std::fmax(high_clip - low_clip, 1.0); // :564
    st.low_clip = low_clip;
    st.high_clip = high_clip;
    st.gamma = gamma;
    if (stats_out) *stats_out = st;

    if (strategy == ORACLE_STRATEGY_CLAHE) { // :572-608
        std::vector<double> norm(n), eq(n);
        for (size_t i = 0; i < n; ++i) {
            if (mask[i]) {
                const double clipped = std::fmin(std::fmax(db[i], low_clip), high_clip);
                norm[i] = (clipped - low_clip) / range;
            } else {
                norm[i] = 0.0;
            }
        }
        clahe_equalize_normalized(norm.data(), mask, rows, cols, 8, 8, 2.0, 256, eq.data(), nullptr);
        for (size_t i = 0; i < n; ++i)
            out[i] = mask[i] ? rs_cast_u16(rs_clamp(eq[i], 0.0, 1.0) * max_val) : (uint16_t)0;
        return;
    }
    for (size_t i = 0; i < n; ++i) // :645-655
        out[i] = mask[i] ? quantize(db[i], low_clip, high_clip, range, gamma, max_val) : (uint16_t)0;
}

// autoscale.rs:710-742
void autoscale_tamed_synrgb_u8(const double* db, const uint8_t* mask, size_t rows, size_t cols, int is_copol,
                               uint8_t* out) {
    oracle_stats st;
    compute_histogram_stats(db, mask, rows, cols, &st, nullptr);
    const size_t n = rows * cols;
    if (st.valid_count == 0) { std::fill(out, out + n, (uint8_t)0); return; }
    double low_clip, high_clip;
    if (is_copol) { low_clip = std::fmin(st.p02, st.p05); high_clip = st.p99; }
    else { low_clip = st.p05; high_clip = st.p99; }
    const double range = std::fmax(high_clip - low_clip, 1.0);
    for (size_t i = 0; i < n; ++i) {
        if (mask[i]) {
            const double clipped = std::fmin(std::fmax(db[i], low_clip), high_clip);
            const double normalized = (clipped - low_clip) / range;
            out[i] = rs_cast_u8(rs_clamp(normalized * 255.0, 0.0, 255.0));
        } else out[i] = 0;
    }
}

// pipeline.rs:8-40
void process_scalar_data_inplace(const float* v, size_t n, double* db, uint8_t* mask) {
    for (size_t i = 0; i < n; ++i) {
        const double magnitude = std::fmax((double)v[i], 1e-10);
        const double db_val = 10.0 * std::log10(magnitude);
        db[i] = db_val;
        mask[i] = db_val > -50.0 ? 1 : 0;
    }
}

// pipeline.rs:42-66 + autoscale.rs:662-704
void process_scalar_data_pipeline(const float* v, size_t rows, size_t cols, int bit_depth, int strategy,
                                  double* db, uint8_t* mask, uint8_t* out_u8, uint16_t* out_u16,
                                  oracle_stats* stats_out) {
    const size_t n = rows * cols;
    process_scalar_data_inplace(v, n, db, mask);
    std::vector<uint16_t> tmp;
    uint16_t* q = out_u16;
    if (bit_depth == ORACLE_U8) { tmp.resize(n); q = tmp.data(); }
    if (strategy == ORACLE_STRATEGY_STANDARD) autoscale_db_image(db, mask, rows, cols, bit_depth, q, stats_out);
    else autoscale_db_image_advanced(db, mask, rows, cols, bit_depth, strategy, q, stats_out);
    if (bit_depth == ORACLE_U8) scale_u16_to_u8(q, n, out_u8);
}

} // namespace

// ---- resize / padding / synrgb live in oracle_resize.cpp / below ------------------------------
extern "C" {

void oracle_process_scalar_data_inplace(const float* v, size_t n, double* db, uint8_t* mask) {
    process_scalar_data_inplace(v, n, db, mask);
}
void oracle_compute_histogram_stats(const double* db, const uint8_t* mask, size_t rows, size_t cols,
                                    oracle_stats* out, uint64_t* hist) {
    compute_histogram_stats(db, mask, rows, cols, out, hist);
}
void oracle_clahe_equalize_normalized(const double* norm, const uint8_t* mask, size_t rows, size_t cols,
                                      size_t tiles_x, size_t tiles_y, double clip_limit, size_t num_bins,
                                      double* out, double* cdfs) {
    clahe_equalize_normalized(norm, mask, rows, cols, tiles_x, tiles_y, clip_limit, num_bins, out, cdfs);
}
void oracle_scale_u16_to_u8(const uint16_t* data, size_t n, uint8_t* out) { scale_u16_to_u8(data, n, out); }
void oracle_autoscale_db_image(const double* db, const uint8_t* mask, size_t rows, size_t cols, int bit_depth,
                               uint16_t* out, oracle_stats* st) {
    autoscale_db_image(db, mask, rows, cols, bit_depth, out, st);
}
void oracle_autoscale_db_image_advanced(const double* db, const uint8_t* mask, size_t rows, size_t cols,
                                        int bit_depth, int strategy, uint16_t* out, oracle_stats* st) {
    autoscale_db_image_advanced(db, mask, rows, cols, bit_depth, strategy, out, st);
}
void oracle_autoscale_db_image_tamed_synrgb_u8(const double* db, const uint8_t* mask, size_t rows, size_t cols,
                                               int is_copol, uint8_t* out) {
    autoscale_tamed_synrgb_u8(db, mask, rows, cols, is_copol, out);
}
void oracle_process_scalar_data_pipeline(const float* v, size_t rows, size_t cols, int bit_depth, int strategy,
                                         double* db_or_null, uint8_t* mask_or_null, uint8_t* out_u8,
                                         uint16_t* out_u16, oracle_stats* st) {
    const size_t n = rows * cols;
    std::vector<double> dbv;
    std::vector<uint8_t> mv;
    double* db = db_or_null;
    uint8_t* mask = mask_or_null;
    if (!db) { dbv.resize(n); db = dbv.data(); }
    if (!mask) { mv.resize(n); mask = mv.data(); }
    process_scalar_data_pipeline(v, rows, cols, bit_depth, strategy, db, mask, out_u8, out_u16, st);
}

// ops.rs:4-44 (f32 arithmetic; log_ratio == ratio)
void oracle_pol_op(int op, const float* a, const float* b, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) {
        const float av = a[i], bv = b[i];
        switch (op) {
        case ORACLE_OP_SUM: out[i] = av + bv; break;
        case ORACLE_OP_DIFF: out[i] = av - bv; break;
        case ORACLE_OP_RATIO:
        case ORACLE_OP_LOGRATIO: out[i] = std::fabs(bv) > 1e-10f ? av / bv : 0.0f; break;
        case ORACLE_OP_NDIFF: {
            const float denom = av + bv;
            out[i] = std::fabs(denom) > 1e-10f ? (av - bv) / denom : 0.0f;
            break;
        }
        default: out[i] = 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// synthetic_rgb.rs
// ---------------------------------------------------------------------------------------------
// synthetic_rgb.rs:10-67
void oracle_create_synthetic_rgb(const uint8_t* b1, const uint8_t* b2, size_t n, uint8_t* rgb) {
    const float GAMMA_R = 0.7f, GAMMA_G = 0.9f, GAMMA_B = 0.1f, SCALE_255 = 255.0f, BLUE_SCALE = 0.24f;
    uint8_t lut_r[256], lut_g[256];
    for (int v = 0; v <= 255; ++v) {
        const float vf = (float)v / SCALE_255;
        lut_r[v] = rs_cast_u8f(rs_clampf(roundf(powf(vf, GAMMA_R) * SCALE_255), 0.0f, 255.0f));
        lut_g[v] = rs_cast_u8f(rs_clampf(roundf(powf(vf, GAMMA_G) * SCALE_255), 0.0f, 255.0f));
    }
    std::vector<uint8_t> lut_b(65536);
    for (int v1 = 0; v1 <= 255; ++v1)
        for (int v2 = 0; v2 <= 255; ++v2) {
            uint8_t blue;
            if (v2 == 0) blue = 0;
            else {
                const float r = (float)lut_r[v1], g = (float)lut_g[v2];
                const float ratio = r / g; // g==0 -> inf (or NaN for 0/0 -> cast gives 0)
                blue = rs_cast_u8f(roundf(rs_clampf(powf(ratio, GAMMA_B) * SCALE_255 * BLUE_SCALE, 0.0f, 255.0f)));
            }
            lut_b[(v1 << 8) | v2] = blue;
        }
    for (size_t i = 0; i < n; ++i) {
        const int v1 = b1[i], v2 = b2[i];
        rgb[3 * i + 0] = lut_r[v1];
        rgb[3 * i + 1] = lut_g[v2];
        rgb[3 * i + 2] = lut_b[(v1 << 8) | v2];
    }
}

// synthetic_rgb.rs:88-178
void oracle_create_synthetic_rgb_suppressed(const uint8_t* b1, const uint8_t* b2, size_t n, uint8_t* rgb) {
    uint32_t histogram[256] = {0};
    auto sat_inc = [](uint32_t& x) { if (x != UINT32_MAX) x += 1; };
    for (size_t i = 0; i < n; ++i) sat_inc(histogram[b1[i]]);
    for (size_t i = 0; i < n; ++i) sat_inc(histogram[b2[i]]);
    const uint32_t total_count = (uint32_t)(n + n);               // `as u32` wraps (:99)
    const uint32_t target_count = rs_cast_u32(std::round((double)total_count * 0.05));
    uint32_t cumulative = 0;
    size_t floor_value = 0;
    for (int i = 0; i <= 255; ++i) {
        const uint64_t s = (uint64_t)cumulative + histogram[i];
        cumulative = s > UINT32_MAX ? UINT32_MAX : (uint32_t)s;
        if (cumulative >= target_count) { floor_value = (size_t)i; break; }
    }
    const uint8_t floor_with_cushion = (uint8_t)std::min<size_t>(floor_value + 3, 40);
    const float SCALE_255 = 255.0f, GAMMA_R_SUPP = 1.15f, GAMMA_G_SUPP = 1.10f;
    const float floor_f = (float)floor_with_cushion;
    const float denom = std::fmax(255.0f - floor_f, 1.0f);
    uint8_t lut_r[256], lut_g[256];
    for (int v = 0; v <= 255; ++v) {
        if ((uint8_t)v <= floor_with_cushion) { lut_r[v] = 0; lut_g[v] = 0; }
        else {
            const float shifted = ((float)v - floor_f) / denom;
            lut_r[v] = rs_cast_u8f(rs_clampf(roundf(powf(shifted, GAMMA_R_SUPP) * SCALE_255), 0.0f, 255.0f));
            lut_g[v] = rs_cast_u8f(rs_clampf(roundf(powf(shifted, GAMMA_G_SUPP) * SCALE_255), 0.0f, 255.0f));
        }
    }
    const float GAMMA_B = 0.1f, BLUE_SCALE_SUPP = 0.18f, EPS = 8.0f;
    std::vector<uint8_t> lut_b(65536);
    for (int v1 = 0; v1 <= 255; ++v1)
        for (int v2 = 0; v2 <= 255; ++v2) {
            const float r = (float)lut_r[v1], g = (float)lut_g[v2];
            const float ratio = (r + EPS) / (g + EPS);
            lut_b[(v1 << 8) | v2] =
                rs_cast_u8f(roundf(rs_clampf(powf(ratio, GAMMA_B) * SCALE_255 * BLUE_SCALE_SUPP, 0.0f, 255.0f)));
        }
    for (size_t i = 0; i < n; ++i) {
        const uint8_t v1 = b1[i], v2 = b2[i];
        if (v1 <= floor_with_cushion && v2 <= floor_with_cushion) {
            rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = 0;
            continue;
        }
        rgb[3 * i + 0] = lut_r[v1];
        rgb[3 * i + 1] = lut_g[v2];
        rgb[3 * i + 2] = lut_b[((int)v1 << 8) | v2];
    }
}

// synthetic_rgb.rs:72-79, 182-197 (all four modes alias Default)
void oracle_create_synthetic_rgb_by_mode_and_strategy(int mode, int strategy, const uint8_t* b1,
                                                      const uint8_t* b2, size_t n, uint8_t* rgb) {
    (void)mode;
    if (strategy == ORACLE_STRATEGY_TAMED || strategy == ORACLE_STRATEGY_CLAHE)
        oracle_create_synthetic_rgb_suppressed(b1, b2, n, rgb);
    else
        oracle_create_synthetic_rgb(b1, b2, n, rgb);
}

// ---------------------------------------------------------------------------------------------
// Orchestration (caller side)
// ---------------------------------------------------------------------------------------------
// save.rs:49-65 (TIFF) / 119-134 (JPEG forces U8); api/mod.rs:84-130, 250-281
int oracle_pipeline_single(const float* v, size_t rows, size_t cols, int format, int bit_depth, int strategy,
                           int has_target, size_t target, int pad, uint8_t* out_u8, uint16_t* out_u16,
                           oracle_resize_meta* meta) {
    if (format == ORACLE_FORMAT_JPEG) bit_depth = ORACLE_U8;
    const size_t n = rows * cols;
    std::vector<uint8_t> s8(bit_depth == ORACLE_U8 ? n : 0);
    std::vector<uint16_t> s16(bit_depth == ORACLE_U16 ? n : 0);
    oracle_process_scalar_data_pipeline(v, rows, cols, bit_depth, strategy, nullptr, nullptr, s8.data(),
                                        s16.data(), nullptr);
    return oracle_resize_image_data_with_meta(s8.data(), bit_depth == ORACLE_U16 ? s16.data() : nullptr, cols,
                                              rows, has_target, target, bit_depth, pad, out_u8, out_u16, meta);
}

// save.rs:199-316; api/mod.rs:133-200. Two bands, independent statistics, band 2 resized with band 1's dims.
int oracle_pipeline_multiband_tiff(const float* v1, const float* v2, size_t rows, size_t cols, int bit_depth,
                                   int strategy, int has_target, size_t target, int pad, uint8_t* out1_u8,
                                   uint16_t* out1_u16, uint8_t* out2_u8, uint16_t* out2_u16,
                                   oracle_resize_meta* meta) {
    int rc = oracle_pipeline_single(v1, rows, cols, ORACLE_FORMAT_TIFF, bit_depth, strategy, has_target, target,
                                    pad, out1_u8, out1_u16, meta);
    if (rc) return rc;
    oracle_resize_meta m2;
    return oracle_pipeline_single(v2, rows, cols, ORACLE_FORMAT_TIFF, bit_depth, strategy, has_target, target,
                                  pad, out2_u8, out2_u16, &m2);
}

// save.rs:317-368 (tamed_band_step != 0) or api/mod.rs:203-247 (tamed_band_step == 0)
int oracle_pipeline_synrgb_jpeg(const float* v1, const float* v2, size_t rows, size_t cols, int strategy,
                                int mode, int has_target, size_t target, int pad, int tamed_band_step,
                                uint8_t* out_rgb, oracle_resize_meta* meta) {
    const size_t n = rows * cols;
    size_t oc, orr;
    oracle_resize_output_dims(cols, rows, has_target, target, pad, &oc, &orr);
    std::vector<uint8_t> f1(oc * orr), f2(oc * orr);
    const float* bands[2] = {v1, v2};
    uint8_t* finals[2] = {f1.data(), f2.data()};
    for (int b = 0; b < 2; ++b) {
        std::vector<double> db(n);
        std::vector<uint8_t> mask(n), s8(n);
        oracle_process_scalar_data_pipeline(bands[b], rows, cols, ORACLE_U8, strategy, db.data(), mask.data(),
                                            s8.data(), nullptr, nullptr);
        if (tamed_band_step && strategy == ORACLE_STRATEGY_TAMED) // save.rs:324-328, 347-351
            autoscale_tamed_synrgb_u8(db.data(), mask.data(), rows, cols, b == 0 ? 1 : 0, s8.data());
        oracle_resize_meta m;
        int rc = oracle_resize_image_data_with_meta(s8.data(), nullptr, cols, rows, has_target, target, ORACLE_U8,
                                                    pad, finals[b], nullptr, b == 0 ? meta : &m);
        if (rc) return rc;
    }
    oracle_create_synthetic_rgb_by_mode_and_strategy(mode, strategy, f1.data(), f2.data(), oc * orr, out_rgb);
    return 0;
}

} // extern "C"

// plan.cpp — host-side planner (see plan.h). Compiled with -ffp-contract=off: the reference is Rust,
// which never fuses a*b+c, and its f64/f32 results are reproduced operation by operation.
#include "plan.h"

#include <chrono>
#include <memory>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>

namespace sarpro {
namespace {

// Rust `as` casts from float: truncate toward zero, saturate, NaN -> 0.
inline uint64_t cast_u64(double x) {
    if (!(x == x) || x <= 0.0) return 0;
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}
inline uint32_t cast_u32(double x) {
    if (!(x == x) || x <= 0.0) return 0;
    if (x >= 4294967295.0) return UINT32_MAX;
    return (uint32_t)x;
}
inline uint16_t cast_u16(double x) {
    if (!(x == x) || x <= 0.0) return 0;
    if (x >= 65535.0) return 65535;
    return (uint16_t)x;
}
inline uint8_t cast_u8(double x) {
    if (!(x == x) || x <= 0.0) return 0;
    if (x >= 255.0) return 255;
    return (uint8_t)x;
}
inline uint8_t cast_u8f(float x) {
    if (!(x == x) || x <= 0.0f) return 0;
    if (x >= 255.0f) return 255;
    return (uint8_t)x;
}
// Rust clamp: NaN propagates.
// v.max(lo).min(hi) (autoscale.rs:440, 583, 649, 734) for operands that are never NaN, without the libm calls
inline double clip_nn(double v, double lo, double hi) {
    const double a = v < lo ? lo : v;
    return a > hi ? hi : a;
}
// f64::round (half away from zero) for 0 <= v < 2^31: v - trunc(v) is exact
inline double round_nn(double v) {
    const double t = (double)(int32_t)v;
    return (v - t >= 0.5) ? t + 1.0 : t;
}
inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

std::vector<double> g_db_table;
std::once_flag g_db_once;

// autoscale.rs:120-140 over a 4096-bin histogram
double estimate_percentile(const uint64_t* hist, uint64_t n, double min_db, double max_db, double span, double p) {
    uint64_t target = cast_u64(std::floor(p * (double)n));
    if (target >= n) target = n - 1;
    uint64_t cumsum = 0;
    for (int b = 0; b < kStatBins; ++b) {
        const uint64_t h = hist[b];
        const uint64_t next = cumsum + h;
        if (target < next) {
            const uint64_t within = target >= cumsum ? target - cumsum : 0;
            const double frac = h > 0 ? (double)within / (double)h : 0.0;
            const double bin_width = span / (double)kStatBins;
            const double bin_start = min_db + (double)b * bin_width;
            return bin_start + frac * bin_width;
        }
        cumsum = next;
    }
    return max_db;
}

inline bool approx_eq(double a, double b) { return std::fabs(a - b) < 1e-9; }

} // namespace

const double* dn_db_table() {
    std::call_once(g_db_once, [] {
        g_db_table.resize(kDnBins);
        for (int dn = 0; dn < kDnBins; ++dn) {
            const float v = (float)dn; // gdal.rs:123 reads the u16 raster as f32
            const double magnitude = std::fmax((double)v, 1e-10);
            g_db_table[dn] = 10.0 * std::log10(magnitude);
        }
    });
    return g_db_table.data();
}

void stats_from_stat_histogram(const uint64_t* hist, uint64_t count, double min_db, double max_db, double mean_db,
                               double std_db, sarpro_stats* st) {
    std::memset(st, 0, sizeof(*st));
    if (count == 0) return; // autoscale.rs:57-76
    st->valid_count = count;
    st->min_db = min_db;
    st->max_db = max_db;
    st->mean_db = mean_db;
    st->std_db = count > 1 ? std_db : 0.0;
    if (std::fabs(max_db - min_db) < std::numeric_limits<double>::epsilon()) { // autoscale.rs:81-100
        st->median_db = min_db;
        st->p01 = st->p02 = st->p05 = st->p10 = st->p25 = min_db;
        st->p75 = st->p90 = st->p95 = st->p98 = st->p99 = max_db;
        return;
    }
    const double span = max_db - min_db;
    auto pct = [&](double p) { return estimate_percentile(hist, count, min_db, max_db, span, p); };
    st->median_db = pct(0.5);
    st->p01 = pct(0.01);
    st->p02 = pct(0.02);
    st->p05 = pct(0.05);
    st->p10 = pct(0.10);
    st->p25 = pct(0.25);
    st->p75 = pct(0.75);
    st->p90 = pct(0.90);
    st->p95 = pct(0.95);
    st->p98 = pct(0.98);
    st->p99 = pct(0.99);
}

void choose_window(int strategy, PlanKind kind, sarpro_stats* st) {
    const double min_db = st->min_db, max_db = st->max_db;
    double low, high, gamma = 1.0;
    if (kind == PlanKind::TamedSynRgbCopol) { // autoscale.rs:721-723
        low = std::fmin(st->p02, st->p05);
        high = st->p99;
    } else if (kind == PlanKind::TamedSynRgbCross) { // autoscale.rs:724-727
        low = st->p05;
        high = st->p99;
    } else if (strategy == SARPRO_STRATEGY_STANDARD) { // pipeline.rs:50-52 -> autoscale.rs:404-428
        const double dynamic_range = max_db - min_db;
        const double iqr = st->p75 - st->p25;
        if (dynamic_range < 15.0) {
            const double range = std::fmax(20.0, dynamic_range * 0.8);
            low = st->median_db - range / 2.0;
            high = st->median_db + range / 2.0;
            gamma = 1.1;
        } else if (iqr < 5.0) {
            const double outlier_factor = 2.5;
            low = st->p25 - outlier_factor * iqr;
            high = st->p75 + outlier_factor * iqr;
            gamma = 1.0;
        } else if (dynamic_range > 40.0) {
            low = std::fmax(st->p02, min_db + 0.02 * dynamic_range);
            high = std::fmin(st->p98, max_db - 0.02 * dynamic_range);
            gamma = 0.9;
        } else {
            low = st->p02;
            high = st->p98;
            gamma = 1.0;
        }
        low = std::fmax(low, min_db);
        high = std::fmin(high, max_db);
    } else {
        const double iqr = st->p75 - st->p25;
        switch (strategy) {
        case SARPRO_STRATEGY_ROBUST: { // autoscale.rs:492-499
            const double outlier_threshold = 2.5 * iqr;
            low = std::fmax(std::fmax(st->p25 - outlier_threshold, st->p01), min_db);
            high = std::fmin(std::fmin(st->p75 + outlier_threshold, st->p99), max_db);
            break;
        }
        case SARPRO_STRATEGY_ADAPTIVE: { // autoscale.rs:500-538
            const double skew_factor = (st->mean_db - st->median_db) / std::fmax(std::fabs(st->std_db), 1.0);
            const double tail_heaviness = (st->p99 - st->p95) / std::fmax(st->p95 - st->p75, 1.0);
            double low_pct, high_pct;
            if (std::fabs(skew_factor) > 0.5) {
                if (skew_factor > 0.0) { low_pct = 0.02; high_pct = 0.98; gamma = 0.9; }
                else { low_pct = 0.05; high_pct = 0.95; gamma = 1.1; }
            } else if (tail_heaviness > 2.0) { low_pct = 0.10; high_pct = 0.90; gamma = 0.8; }
            else { low_pct = 0.05; high_pct = 0.95; gamma = 1.0; }
            if (approx_eq(low_pct, 0.10)) low = st->p10;
            else if (approx_eq(low_pct, 0.02)) low = st->p02;
            else low = st->p05;
            if (approx_eq(high_pct, 0.90)) high = st->p90;
            else if (approx_eq(high_pct, 0.98)) high = st->p98;
            else high = st->p95;
            break;
        }
        case SARPRO_STRATEGY_EQUALIZED: // autoscale.rs:539-543
        case SARPRO_STRATEGY_CLAHE:     // autoscale.rs:544-548
            low = st->p01; high = st->p99; break;
        case SARPRO_STRATEGY_TAMED:     // autoscale.rs:549-553
            low = st->p25; high = st->p99; break;
        default:                        // Default (and the unreachable Standard arm), autoscale.rs:554-561
            low = st->p05; high = st->p95; break;
        }
    }
    st->low_clip = low;
    st->high_clip = high;
    st->gamma = gamma;
}

void make_u16_to_u8_remap(uint16_t mn, uint16_t mx, int n_entries, uint8_t* remap) {
    // autoscale.rs:352-363
    const float fmin_ = (float)mn, fmax_ = (float)mx;
    const float scale = fmax_ > fmin_ ? 255.0f / (fmax_ - fmin_) : 1.0f;
    for (int x = 0; x < n_entries; ++x) {
        const float val = roundf(((float)x - fmin_) * scale);
        remap[x] = cast_u8f(clampf(val, 0.0f, 255.0f));
    }
}

namespace {

// All 11 percentiles of autoscale.rs:142-159 in one walk over the 4096-bin histogram (same arithmetic as
// estimate_percentile, autoscale.rs:120-140).
void stats_percentiles_one_walk(const uint64_t* hist, uint64_t n, double min_db, double max_db, sarpro_stats* st) {
    static const double ps[11] = {0.01, 0.02, 0.05, 0.10, 0.25, 0.5, 0.75, 0.90, 0.95, 0.98, 0.99};
    double* dst[11] = {&st->p01, &st->p02, &st->p05, &st->p10, &st->p25, &st->median_db, &st->p75, &st->p90, &st->p95, &st->p98, &st->p99};
    uint64_t target[11];
    for (int i = 0; i < 11; ++i) {
        uint64_t t = cast_u64(std::floor(ps[i] * (double)n));
        if (t >= n) t = n - 1;
        target[i] = t; // non-decreasing in i
        *dst[i] = max_db;
    }
    const double span = max_db - min_db;
    const double bin_width = span / (double)kStatBins;
    uint64_t cumsum = 0;
    int i = 0;
    for (int b = 0; b < kStatBins && i < 11; ++b) {
        const uint64_t h = hist[b];
        const uint64_t next = cumsum + h;
        while (i < 11 && target[i] < next) {
            const uint64_t within = target[i] >= cumsum ? target[i] - cumsum : 0;
            const double frac = h > 0 ? (double)within / (double)h : 0.0;
            const double bin_start = min_db + (double)b * bin_width;
            *dst[i] = bin_start + frac * bin_width;
            ++i;
        }
        cumsum = next;
    }
}

} // namespace

// SARPRO_TRACE: host time stamps inside the planner (us since the call): scan done, stats, percentiles, table
double g_plan_trace_us[6];
bool g_plan_trace_on = false;
static inline void plan_stamp(int i, const std::chrono::steady_clock::time_point& t0) {
    if (g_plan_trace_on) g_plan_trace_us[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
}

struct Present { uint32_t dn; uint64_t h; };
// the valid present DNs of the plan in progress, ascending (one uninitialised buffer per thread, sized for every DN)
struct PresentList {
    Present* b = nullptr;
    Present* e = nullptr;
    const Present* begin() const { return b; }
    const Present* end() const { return e; }
    size_t size() const { return (size_t)(e - b); }
    const Present& operator[](size_t i) const { return b[i]; }
};
static thread_local std::unique_ptr<Present[]> t_present_buf;

// Start of a plan: cleared outputs. Returns the (thread-local) list the caller fills with the valid present DNs in
// ascending order through a PlanCollector.
static PresentList plan_begin(BandPlan* out) {
    if (out->lut.size() != (size_t)kDnBins) out->lut.resize(kDnBins);
    std::memset(out->lut.data(), 0, (size_t)kDnBins * sizeof(uint16_t)); // (vector::assign is a 2-byte loop at -O2: 20 us)
    out->clahe = false;
    out->any_valid = false;
    out->have_invalid = false;
    out->pre_min = out->pre_max = 0;
    out->max_present_dn = 0;
    out->sat_from_dn = 0;
    out->px_total = out->px_ge1024 = out->px_ge2048 = 0;
    std::memset(&out->stats, 0, sizeof(out->stats));
    if (!t_present_buf) t_present_buf.reset(new Present[kDnBins]);
    PresentList l;
    l.b = l.e = t_present_buf.get();
    return l;
}
// Collector of the present DNs (ascending): counters stay in registers, the list is written through a raw pointer
// (updating the plan's fields and push_back per entry made this the slowest part of the planner).
struct PlanCollector {
    const double* db;
    Present* w;
    uint32_t max_dn = 0;
    uint64_t total = 0, ge1k = 0, ge2k = 0;
    bool invalid = false;
    inline void take(uint32_t dn, uint64_t h) {
        if (!h) return;
        max_dn = dn;
        total += h;
        ge1k += dn >= 1024 ? h : 0;
        ge2k += dn >= 2048 ? h : 0;
        if (db[dn] > -50.0) *w++ = Present{dn, h}; // pipeline.rs:22
        else invalid = true;
    }
    void finish(BandPlan* out, PresentList& present) {
        present.e = w;
        out->max_present_dn = max_dn;
        out->px_total = total;
        out->px_ge1024 = ge1k;
        out->px_ge2048 = ge2k;
        out->have_invalid = invalid;
    }
};
static void plan_finish(const PresentList& present, int bit_depth, int strategy, PlanKind kind, BandPlan* out,
                        const std::chrono::steady_clock::time_point& t0);

template <typename CountT>
static void plan_from_dn_histogram_t(const CountT* hist, int bit_depth, int strategy, PlanKind kind, BandPlan* out, int top_hint = -1) {
    const auto t0 = std::chrono::steady_clock::now();
    const double* db = dn_db_table();
    // Distinct sample values present in the raster (typically ~1e3 of the 65,536 DNs): everything below is
    // evaluated once per distinct value instead of once per pixel.
    PresentList present = plan_begin(out);
    plan_stamp(0, t0);
    // the brightest present DN first (wide loads from the top): a GRD band uses a few thousand of the 65,536 bins, and
    // this scan sits on the critical path between pass A and pass B
    int top = kDnBins;
    if (top_hint >= 0 && top_hint <= kDnBins) {
        top = top_hint; // (the pinned histogram was just written by the device: every line read here comes from DRAM)
    } else if (sizeof(CountT) == 4) {
        const uint64_t* h8 = reinterpret_cast<const uint64_t*>(hist);
        int i = kDnBins / 2;
        while (i >= 4 && !(h8[i - 1] | h8[i - 2] | h8[i - 3] | h8[i - 4])) i -= 4;
        top = i * 2;
    }
    while (top > 0 && !hist[top - 1]) --top;
    PlanCollector col{db, present.b};
    int dn0 = 0;
    if (sizeof(CountT) == 4) { // above the speckle range only point targets are present: skip empty bins eight at a time
        for (; dn0 + 8 <= top; dn0 += 8) {
            uint64_t q[4];
            std::memcpy(q, hist + dn0, 32);
            if (!(q[0] | q[1] | q[2] | q[3])) continue;
            for (int k = 0; k < 8; ++k) col.take((uint32_t)(dn0 + k), hist[dn0 + k]);
        }
    }
    for (; dn0 < top; ++dn0) col.take((uint32_t)dn0, hist[dn0]);
    col.finish(out, present);
    plan_finish(present, bit_depth, strategy, kind, out, t0);
}

bool plan_from_present_list(const uint32_t* blk, const uint32_t* pairs, uint32_t cap, int bit_depth, int strategy, PlanKind kind,
                            BandPlan* out) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int b = 0; b < 256; ++b)
        if ((uint64_t)blk[2 * b] + blk[2 * b + 1] > cap) return false; // more present DNs than the list holds
    const double* db = dn_db_table();
    PresentList present = plan_begin(out);
    plan_stamp(0, t0);
    PlanCollector col{db, present.b};
    for (int b = 0; b < 256; ++b) {
        const uint32_t* p = pairs + 2 * (size_t)blk[2 * b];
        for (uint32_t i = 0; i < blk[2 * b + 1]; ++i) col.take(p[2 * i] & 0xffffu, p[2 * i + 1]);
    }
    col.finish(out, present);
    plan_finish(present, bit_depth, strategy, kind, out, t0);
    return true;
}

static void plan_finish(const PresentList& present, int bit_depth, int strategy, PlanKind kind, BandPlan* out,
                        const std::chrono::steady_clock::time_point& t0) {
    const double* db = dn_db_table();
    const bool have_invalid = out->have_invalid;
    plan_stamp(1, t0);
    // Pass 1 of compute_histogram_stats (autoscale.rs:37-55) over distinct values.
    uint64_t count = 0;
    double min_db = std::numeric_limits<double>::infinity();
    double max_db = -std::numeric_limits<double>::infinity();
    long double sum = 0.0L;
    for (const Present& p : present) {
        count += p.h;
        if (db[p.dn] < min_db) min_db = db[p.dn];
        if (db[p.dn] > max_db) max_db = db[p.dn];
        sum += (long double)p.h * (long double)db[p.dn];
    }
    if (count == 0) return; // all-zero output (autoscale.rs:376-378, 466-468, 716-718); lut already 0
    out->any_valid = true;
    // Mean / population std. The reference runs Welford serially over 4e8 pixels (autoscale.rs:49-53,78);
    // its rounding depends on pixel order and cannot be reproduced from a histogram. The values below are
    // the correctly-rounded-to-~1e-15 statistics; they feed log lines and the Adaptive branch test only.
    const long double mean_l = sum / (long double)count;
    long double m2 = 0.0L;
    for (const Present& p : present) {
        const long double d = (long double)db[p.dn] - mean_l;
        m2 += (long double)p.h * d * d;
    }
    const double mean_db = (double)mean_l;
    const double std_db = count > 1 ? (double)sqrtl(m2 / (long double)count) : 0.0;

    plan_stamp(2, t0);
    // Pass 2 (autoscale.rs:103-117): 4096-bin histogram over [min,max].
    uint64_t h4096[kStatBins];
    const bool degenerate = std::fabs(max_db - min_db) < std::numeric_limits<double>::epsilon();
    sarpro_stats& st = out->stats;
    st.valid_count = count;
    st.min_db = min_db;
    st.max_db = max_db;
    st.mean_db = mean_db;
    st.std_db = std_db;
    if (degenerate) { // autoscale.rs:81-100
        st.median_db = min_db;
        st.p01 = st.p02 = st.p05 = st.p10 = st.p25 = min_db;
        st.p75 = st.p90 = st.p95 = st.p98 = st.p99 = max_db;
    } else {
        std::memset(h4096, 0, sizeof(h4096));
        const double span = max_db - min_db;
        const double inv_span = 1.0 / span;
        for (const Present& p : present) {
            const double t = clampd((db[p.dn] - min_db) * inv_span, 0.0, 1.0);
            uint64_t idx = cast_u64(t * (double)kStatBins);
            if (idx >= (uint64_t)kStatBins) idx = kStatBins - 1;
            h4096[idx] += p.h;
        }
        stats_percentiles_one_walk(h4096, count, min_db, max_db, &st);
    }
    plan_stamp(3, t0);
    choose_window(strategy, kind, &st);
    // The lowest present DN from which every present DN up to the brightest one carries that one's table word (the tables
    // are monotone and saturate above the window): bounds the shared-memory table of pass B (hpipe_hot_from_plan).
    auto set_sat_from = [&]() {
        const uint32_t top_word = out->lut[out->max_present_dn] & 255u;
        uint32_t h = out->max_present_dn;
        bool all = true;
        for (size_t i = present.size(); i-- > 0;) {
            if (present[i].dn > out->max_present_dn) continue;
            if ((out->lut[present[i].dn] & 255u) == top_word) h = present[i].dn; else { all = false; break; }
        }
        if (all && have_invalid && top_word == 0) h = 0; // invalid DNs are present with word 0
        out->sat_from_dn = h;
    };
    const double low = st.low_clip, high = st.high_clip, gamma = st.gamma;
    const double range = std::fmax(high - low, 1.0); // autoscale.rs:429, 564, 729

    const bool tamed_rgb = kind != PlanKind::Autoscale;
    if (!tamed_rgb && strategy == SARPRO_STRATEGY_CLAHE) {
        // autoscale.rs:582-591 normalisation + :263 / :320 bin index; the blend runs on the device.
        out->clahe = true;
        for (const Present& p : present) {
            const double clipped = clip_nn(db[p.dn], low, high);
            const double n = (clipped - low) / range;
            const double v = clampd(n, 0.0, 1.0);
            const double b = round_nn(v * ((double)kClaheBins - 1.0)); // v in [0, 1]
            long long bin = (b == b) ? (long long)b : 0;
            if (bin < 0) bin = 0;
            if (bin >= kClaheBins) bin = kClaheBins - 1;
            out->lut[p.dn] = (uint16_t)bin;
        }
        set_sat_from();
        plan_stamp(4, t0);
        return;
    }

    const double max_val = (tamed_rgb || bit_depth == SARPRO_U8) ? 255.0 : 65535.0;
    uint16_t mn = 65535, mx = 0;
    if (have_invalid) { mn = 0; mx = 0; } // invalid pixels are written as 0 (autoscale.rs:444, 653, 738)
    for (const Present& p : present) {
        const double clipped = clip_nn(db[p.dn], low, high);
        uint16_t q;
        if (tamed_rgb) { // autoscale.rs:734-736
            const double normalized = (clipped - low) / range;
            q = cast_u8(clampd(normalized * 255.0, 0.0, 255.0));
        } else {         // autoscale.rs:440-442 / 649-651
            const double normalized = std::pow((clipped - low) / range, gamma);
            q = cast_u16(clampd(normalized * max_val, 0.0, max_val));
        }
        out->lut[p.dn] = q;
        if (q < mn) mn = q;
        if (q > mx) mx = q;
    }
    out->pre_min = mn;
    out->pre_max = mx;
    if (!tamed_rgb && bit_depth == SARPRO_U8) {
        // scale_u16_to_u8 over ALL pixels incl. invalid zeros (autoscale.rs:669-670, 691-693)
        uint8_t remap[256];
        make_u16_to_u8_remap(mn, mx, 256, remap);
        for (const Present& p : present) out->lut[p.dn] = remap[out->lut[p.dn] > 255 ? 255 : out->lut[p.dn]];
    }
    set_sat_from();
}

void plan_from_dn_histogram(const uint64_t* hist, int bit_depth, int strategy, PlanKind kind, BandPlan* out) {
    plan_from_dn_histogram_t<uint64_t>(hist, bit_depth, strategy, kind, out);
}
void plan_from_dn_histogram32(const uint32_t* hist, int bit_depth, int strategy, PlanKind kind, BandPlan* out, int top_hint) {
    plan_from_dn_histogram_t<uint32_t>(hist, bit_depth, strategy, kind, out, top_hint);
}

ClaheGeom clahe_geometry(uint64_t rows, uint64_t cols) {
    ClaheGeom g;
    g.rows = rows;
    g.cols = cols;
    g.tile_h = (rows + kClaheTiles - 1) / kClaheTiles; // autoscale.rs:235
    g.tile_w = (cols + kClaheTiles - 1) / kClaheTiles; // autoscale.rs:236
    return g;
}

// ---------------------------------------------------------------------------------------------
// Lanczos3 tables: fast_image_resize 5.x `precompute_coefficients` + Normalizer16/32 (see the
// restatement notes in DESIGN.md; the crate is not vendored in the reference tree).
// ---------------------------------------------------------------------------------------------
namespace {
inline double sinc(double x) {
    if (x == 0.0) return 1.0;
    x *= M_PI;
    return std::sin(x) / x;
}
inline double lanczos3(double x) { return (x >= -3.0 && x < 3.0) ? sinc(x) * sinc(x / 3.0) : 0.0; }
} // namespace

void build_lanczos3_axis(uint32_t in_size, uint32_t out_size, bool wide, ResampleAxis* ax) {
    ax->in_size = in_size;
    ax->out_size = out_size;
    ax->start.assign(out_size, 0);
    ax->size.assign(out_size, 0);
    ax->coef.clear();
    ax->window = 0;
    ax->precision = 0;
    if (in_size == 0 || out_size == 0) return;
    const double scale = (double)in_size / (double)out_size;
    const double filter_scale = std::fmax(scale, 1.0);
    const double radius = 3.0 * filter_scale;
    const uint32_t window = (uint32_t)std::ceil(radius) * 2 + 1;
    const double recip = 1.0 / filter_scale;
    ax->window = window;
    std::vector<double> w((size_t)window * out_size, 0.0);
    double max_w = 0.0;
    bool first = true;
    for (uint32_t ox = 0; ox < out_size; ++ox) {
        const double in_center = ((double)ox + 0.5) * scale;
        const uint32_t x_min = (uint32_t)std::fmax(std::floor(in_center - radius), 0.0);
        const uint32_t x_max = (uint32_t)std::fmin(std::ceil(in_center + radius), (double)in_size);
        const double center = in_center - 0.5;
        double* wo = &w[(size_t)ox * window];
        uint32_t n = 0, bstart = x_min, bend = x_max;
        double ww = 0.0;
        for (uint32_t x = x_min; x < x_max; ++x) {
            const double v = lanczos3(((double)x - center) * recip);
            if (x == bstart && v == 0.0) { bstart += 1; continue; } // leading zero taps are dropped
            wo[n++] = v;
            ww += v;
        }
        for (uint32_t k = n; k > 0; --k) { // trailing zero taps shrink the bound
            if (bend <= bstart || wo[k - 1] != 0.0) break;
            bend -= 1;
        }
        if (ww != 0.0)
            for (uint32_t k = 0; k < n; ++k) wo[k] /= ww;
        ax->start[ox] = bstart;
        ax->size[ox] = bend - bstart;
        for (uint32_t k = 0; k < window; ++k) { // max over the whole padded table, like the crate
            if (first || wo[k] > max_w) { max_w = wo[k]; first = false; }
        }
    }
    int precision = 0;
    if (!wide) {
        for (int cur = 0; cur < 22; ++cur) {
            precision = cur;
            if (std::round(max_w * (double)(1 << (precision + 1))) >= (double)(1 << 15)) break;
        }
    } else {
        for (int cur = 0; cur < 46; ++cur) {
            precision = cur;
            if (std::round(max_w * (double)((int64_t)1 << (precision + 1))) >= (double)((int64_t)1 << 31)) break;
        }
    }
    ax->precision = precision;
    const double fscale = (double)((int64_t)1 << precision);
    ax->coef.assign((size_t)window * out_size, 0);
    for (uint32_t ox = 0; ox < out_size; ++ox)
        for (uint32_t k = 0; k < ax->size[ox]; ++k) {
            const double r = std::round(w[(size_t)ox * window + k] * fscale);
            ax->coef[(size_t)ox * window + k] = wide ? (int32_t)r : (int32_t)(int16_t)r;
        }
}

void calculate_resize_dimensions(size_t cols, size_t rows, size_t target, size_t* new_cols, size_t* new_rows) {
    const size_t short_side = std::min(rows, cols), long_side = std::max(rows, cols);
    if (target > long_side) { *new_cols = cols; *new_rows = rows; return; } // resize.rs:14-20 (no upscaling)
    const double scale_factor = (double)target / (double)long_side;
    const size_t new_short = (size_t)cast_u64(std::round((double)short_side * scale_factor));
    if (cols > rows) { *new_cols = target; *new_rows = new_short; }
    else { *new_cols = new_short; *new_rows = target; }
}

void resize_output_dims(size_t cols, size_t rows, bool has_target, size_t target, bool pad, size_t* rc, size_t* rr,
                        size_t* oc, size_t* orr) {
    size_t c = cols, r = rows;
    if (has_target && std::max(cols, rows) != target) calculate_resize_dimensions(cols, rows, target, &c, &r);
    *rc = c;
    *rr = r;
    if (pad) { const size_t m = std::max(c, r); c = m; r = m; }
    *oc = c;
    *orr = r;
}

// ---------------------------------------------------------------------------------------------
// Synthetic RGB LUTs
// ---------------------------------------------------------------------------------------------
void build_synrgb_default_lut(SynRgbLut* lut) {
    const float GAMMA_R = 0.7f, GAMMA_G = 0.9f, GAMMA_B = 0.1f, S = 255.0f, BLUE_SCALE = 0.24f;
    for (int v = 0; v < 256; ++v) {
        const float vf = (float)v / S;
        lut->r[v] = cast_u8f(clampf(roundf(powf(vf, GAMMA_R) * S), 0.0f, 255.0f));
        lut->g[v] = cast_u8f(clampf(roundf(powf(vf, GAMMA_G) * S), 0.0f, 255.0f));
    }
    lut->b.assign(65536, 0);
    for (int v1 = 0; v1 < 256; ++v1)
        for (int v2 = 1; v2 < 256; ++v2) { // v2 == 0 -> blue = 0 (synthetic_rgb.rs:38-39)
            const float ratio = (float)lut->r[v1] / (float)lut->g[v2];
            lut->b[(v1 << 8) | v2] = cast_u8f(roundf(clampf(powf(ratio, GAMMA_B) * S * BLUE_SCALE, 0.0f, 255.0f)));
        }
    lut->floor_with_cushion = -1;
}

void build_synrgb_suppressed_lut(int fwc, SynRgbLut* lut) {
    const float S = 255.0f, GAMMA_R = 1.15f, GAMMA_G = 1.10f, GAMMA_B = 0.1f, BLUE_SCALE = 0.18f, EPS = 8.0f;
    const float floor_f = (float)fwc;
    const float denom = std::fmax(255.0f - floor_f, 1.0f);
    for (int v = 0; v < 256; ++v) {
        if (v <= fwc) { lut->r[v] = 0; lut->g[v] = 0; continue; }
        const float shifted = ((float)v - floor_f) / denom;
        lut->r[v] = cast_u8f(clampf(roundf(powf(shifted, GAMMA_R) * S), 0.0f, 255.0f));
        lut->g[v] = cast_u8f(clampf(roundf(powf(shifted, GAMMA_G) * S), 0.0f, 255.0f));
    }
    lut->b.assign(65536, 0);
    for (int v1 = 0; v1 < 256; ++v1)
        for (int v2 = 0; v2 < 256; ++v2) {
            const float ratio = ((float)lut->r[v1] + EPS) / ((float)lut->g[v2] + EPS);
            lut->b[(v1 << 8) | v2] = cast_u8f(roundf(clampf(powf(ratio, GAMMA_B) * S * BLUE_SCALE, 0.0f, 255.0f)));
        }
    lut->floor_with_cushion = fwc;
}

int synrgb_floor_from_histogram(const uint32_t* hist256, uint64_t n_per_band) {
    // synthetic_rgb.rs:99-113. hist256 holds exact counts; the reference saturates u32 adds and
    // wraps the total to u32, which only matters beyond 2^31 pixels per band.
    const uint32_t total = (uint32_t)(n_per_band + n_per_band);
    const uint32_t target = cast_u32(std::round((double)total * 0.05));
    uint64_t cumulative = 0;
    size_t floor_value = 0;
    for (int i = 0; i < 256; ++i) {
        cumulative = std::min<uint64_t>(cumulative + hist256[i], UINT32_MAX);
        if (cumulative >= target) { floor_value = (size_t)i; break; }
    }
    return (int)std::min<size_t>(floor_value + 3, 40);
}

} // namespace sarpro

// plan_f32.cpp — host thresholds for rasters that are not u16-valued (see kernels_f32.cu).
// Every index the reference computes from a sample is monotone in the sample value; the boundaries are
// located here by evaluating the reference's own f64 expression (same libm) on f32 bit patterns.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>

#include "host_pool.h"
#include "plan.h"

namespace sarpro {
namespace {

inline uint32_t bits_of(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float float_of(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
constexpr uint32_t kInfBits = 0x7f800000u;

inline uint64_t cast_u64(double x) {
    if (!(x == x) || x <= 0.0) return 0;
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}
inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Smallest bit pattern b in [lo, hi] with level(b) >= k, given a monotone `level`; kInfBits if none.
// `guess` is an analytic estimate; a gallop around it brackets the boundary in a handful of evaluations.
template <typename F>
uint32_t first_reaching(F&& level, uint32_t k, uint32_t lo, uint32_t hi, uint32_t guess, uint32_t level_lo, uint32_t level_hi) {
    if (level_hi < k) return kInfBits;
    if (level_lo >= k) return lo;
    // invariant: level(lo) < k <= level(hi)
    uint32_t g = std::min(std::max(guess, lo + 1), hi);
    uint32_t step = 1;
    if (level(g) >= k) {
        hi = g;
        while (hi - lo > 1) {
            const uint32_t t = hi - lo > step ? hi - step : lo + 1;
            if (t <= lo) break;
            if (level(t) >= k) { hi = t; step *= 4; }
            else { lo = t; break; }
        }
    } else {
        lo = g;
        while (hi - lo > 1) {
            const uint32_t t = hi - lo > step ? lo + step : hi - 1;
            if (t >= hi) break;
            if (level(t) < k) { lo = t; step *= 4; }
            else { hi = t; break; }
        }
    }
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (level(mid) >= k) hi = mid; else lo = mid;
    }
    return hi;
}

// ---- analytic thresholds ------------------------------------------------------------------------
// For an index that is LINEAR in dB, idx(v) = trunc(q(v)), q(v) = (10 log10 v - low) / range * n, the real-valued position
// of boundary k is v* = 10^((low + range k/n) / 10), and the threshold is the smallest f32 >= v* -- provided the reference's
// f64 evaluation of q cannot land on the other side of k for the two f32 neighbours of v*. Its error against the real q is
//   |q_f64(v) - q(v)| <= n/range * 2.4e-13 + n * 3.3e-16     (glibc log10 < 2 ulp, |dB| <= 400, one rounding each for the
//                                                             product by 10, the subtraction, the quotient, the product by n)
// while moving v by a relative m moves q by n/range * 10/ln 10 * m = n/range * 4.34 m. With m = 1e-9 (and range <= 1e4 dB)
// the second exceeds the first by four orders of magnitude, and v* itself (exp(dB ln10/10): the rounded argument, |arg| < 90,
// contributes 2e-14 relative, exp < 1 ulp) is known to ~3e-14: if both f32 neighbours of v* are at least m v* away from it, the f64 expression reaches k at
// the upper one and not at the lower one, which is the definition of the threshold. f32 spacing is 6e-8..1.2e-7 relative, so
// ~2 % of the boundaries fall within m of an f32 and take the search below instead. One exp replaces pow + two log10 + the
// clamp arithmetic; tests/test_host_cpu.py compares whole tables of both methods.
constexpr double kAnalyticMargin = 1e-9;
constexpr double kLn10Over10 = 0.23025850929940457;
constexpr double kAnalyticDbMargin = 1e-6;   // the boundary's f32 neighbours (<= 5.2e-7 dB away) must stay clear of the clip ends
std::atomic<bool> g_analytic{true};          // test hook: tables by search only
std::atomic<uint64_t> g_analytic_hits{0};
inline bool analytic_threshold(double vstar, uint32_t lo, uint32_t hi, uint32_t* out) {
    if (!(vstar > 0.0) || !(vstar < 3.0e38)) return false;
    if (!(vstar > 1e-30)) return false;           // normal f32 only: neighbours are bit pattern +- 1
    uint32_t b = bits_of((float)vstar);           // nearest f32
    if ((double)float_of(b) < vstar) ++b;         // smallest f32 >= v* (positive floats are ordered like their bit patterns)
    const double m = kAnalyticMargin * vstar;
    if (!((double)float_of(b) - vstar >= m) || !(vstar - (double)float_of(b - 1) >= m)) return false;
    if (b <= lo || b > hi) return false;          // outside the data range: let the search apply its own end rules
    *out = b;
    return true;
}

} // namespace

double db_of_sample(float v) { return 10.0 * std::log10(std::fmax((double)v, 1e-10)); }

float valid_threshold() {
    uint32_t lo = 0, hi = bits_of(1.0f);
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (db_of_sample(float_of(mid)) > -50.0) hi = mid; else lo = mid;
    }
    return float_of(hi);
}

void build_stat_edges(float min_v, float max_v, std::vector<float>* edges) {
    edges->assign(kStatBins, 0.0f);
    const double min_db = db_of_sample(min_v), max_db = db_of_sample(max_v);
    const double span = max_db - min_db;
    const double inv_span = 1.0 / span;
    auto level = [&](uint32_t b) -> uint32_t {
        const double t = clampd((db_of_sample(float_of(b)) - min_db) * inv_span, 0.0, 1.0);
        uint64_t idx = cast_u64(t * (double)kStatBins);
        if (idx >= (uint64_t)kStatBins) idx = kStatBins - 1;
        return (uint32_t)idx;
    };
    const uint32_t lo = bits_of(min_v), hi = bits_of(max_v);
    const uint32_t l_lo = level(lo), l_hi = level(hi);
    // analytic boundaries (see analytic_threshold): the index is linear in dB over [min_db, max_db]
    const bool analytic = g_analytic.load(std::memory_order_relaxed) && std::isfinite(span) && span >= kStatBins * kAnalyticDbMargin &&
                          span <= 1e4 && std::fabs(min_db) <= 400.0 && std::fabs(max_db) <= 400.0;
    parallel_for(kStatBins - 1, [&](uint32_t a, uint32_t b) {
        uint64_t hits = 0;
        for (uint32_t i = a; i < b; ++i) {
            const uint32_t k = i + 1;
            const double db = min_db + span * ((double)k / (double)kStatBins);
            const double vstar = std::exp(db * kLn10Over10);
            uint32_t bits;
            if (analytic && l_lo < k && k <= l_hi && analytic_threshold(vstar, lo, hi, &bits)) { (*edges)[k] = float_of(bits); ++hits; continue; }
            (*edges)[k] = float_of(first_reaching(level, k, lo, hi, bits_of((float)vstar), l_lo, l_hi));
        }
        if (hits) g_analytic_hits.fetch_add(hits, std::memory_order_relaxed);
    });
}

void build_level_edges(LevelKind kind, double low, double high, double gamma, uint32_t n_levels, float min_v, float max_v,
                       std::vector<float>* edges, uint32_t* level_of_min, uint32_t* level_of_max) {
    edges->assign((size_t)n_levels + 1, 0.0f);
    const double range = std::fmax(high - low, 1.0);
    const double max_val = (double)n_levels;
    auto level = [&](uint32_t b) -> uint32_t {
        const double db = db_of_sample(float_of(b));
        const double clipped = std::fmin(std::fmax(db, low), high);
        const double n = (clipped - low) / range;
        double q;
        if (kind == LevelKind::Quantize) q = clampd((gamma == 1.0 ? n : std::pow(n, gamma)) * max_val, 0.0, max_val); // pow(n, 1.0) == n exactly
        else if (kind == LevelKind::TamedLinearU8) q = clampd(n * 255.0, 0.0, 255.0);
        else {
            q = std::round(clampd(n, 0.0, 1.0) * 255.0);
            if (!(q == q)) q = 0.0;
            q = clampd(q, 0.0, 255.0);
        }
        if (!(q == q) || q <= 0.0) return 0;
        return q >= max_val ? n_levels : (uint32_t)q;
    };
    const uint32_t lo = bits_of(min_v), hi = bits_of(max_v);
    const uint32_t l_lo = level(lo), l_hi = level(hi);
    if (level_of_min) *level_of_min = l_lo;
    if (level_of_max) *level_of_max = l_hi;
    // analytic boundaries (see analytic_threshold) where the level is trunc of a quantity linear in dB: gamma == 1, not the
    // CLAHE bins (a round), boundaries strictly inside the clip window (the top level of a window narrower than 1 dB, whose
    // range was raised, and the level at high_clip itself go through the search)
    const bool analytic = g_analytic.load(std::memory_order_relaxed) && kind != LevelKind::ClaheBin && gamma == 1.0 && std::isfinite(low) &&
                          std::isfinite(high) && range <= 1e4 && std::fabs(low) <= 400.0 && std::fabs(high) <= 400.0 &&
                          (kind != LevelKind::TamedLinearU8 || n_levels == 255);
    parallel_for(n_levels, [&](uint32_t a, uint32_t b) {
        uint64_t hits = 0;
        for (uint32_t i = a; i < b; ++i) {
            const uint32_t k = i + 1;
            double frac = (kind == LevelKind::ClaheBin) ? ((double)k - 0.5) / 255.0 : (double)k / max_val;
            if (kind == LevelKind::Quantize && gamma != 1.0) frac = std::pow(frac, 1.0 / gamma);
            const double db = low + range * frac;
            const double vstar = std::exp(db * kLn10Over10);
            uint32_t bits;
            if (analytic && l_lo < k && k <= l_hi && db - kAnalyticDbMargin >= low && db + kAnalyticDbMargin <= high &&
                analytic_threshold(vstar, lo, hi, &bits)) {
                (*edges)[k] = float_of(bits);
                ++hits;
                continue;
            }
            (*edges)[k] = float_of(first_reaching(level, k, lo, hi, bits_of((float)vstar), l_lo, l_hi));
        }
        if (hits) g_analytic_hits.fetch_add(hits, std::memory_order_relaxed);
    });
}

void f32_edges_set_analytic(bool on) { g_analytic.store(on); }
uint64_t f32_edges_analytic_hits() { return g_analytic_hits.load(); }

void f32_guard(bool on, double low_db, double range_db, uint32_t n, float min_v, float max_v, int* e0, float* f0, float* scale,
               float* guard) {
    *e0 = 0; *f0 = 0.f; *scale = 0.f; *guard = 1.0f;
    if (!on || !(range_db > 0.0) || !std::isfinite(range_db) || !std::isfinite(low_db) || !(min_v > 0.f) || !std::isfinite(max_v)) return;
    const double k = 10.0 * std::log10(2.0);         // dB per octave
    const double y_lo = low_db / k;
    const double fl = std::floor(y_lo);
    if (std::fabs(fl) > 1e6) return;
    const double sc = (double)n * k / range_db;
    const double dmax = std::fmax(std::fabs(std::log2((double)min_v) - y_lo), std::fabs(std::log2((double)max_v) - y_lo)) + 1.0;
    const double e_d = std::ldexp(1.0, -22) + 2.0 * std::ldexp(1.0, -25) + std::ldexp(1.0, -24) * dmax;
    const double e_t = e_d * sc + ((double)n + 2.0) * std::ldexp(1.0, -23) + 1e-6;
    const double g = 1.5 * e_t;
    if (!(g < 0.45) || !(sc < 1e30)) return;         // too coarse to help: compare thresholds
    *e0 = (int)fl;
    *f0 = (float)(y_lo - fl);
    *scale = (float)sc;
    *guard = (float)g;
}

} // namespace sarpro

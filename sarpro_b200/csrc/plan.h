// plan.h — host-side planner of the B200 raster path.
//
// The device passes reduce a raster to integer histograms; everything that is O(bins) and needs
// the host libm (log10 / pow / powf, so results equal the reference's, which calls the same libm
// through Rust's f64::log10 / powf) happens here, once per band, and is shipped back as small
// look-up tables.  Follows, from the reference:
//   pipeline.rs:19-22            dB + validity per distinct sample value
//   autoscale.rs:35-160          compute_histogram_stats, evaluated over the value histogram
//   autoscale.rs:368-448         autoscale_db_image      (Standard)
//   autoscale.rs:452-659         autoscale_db_image_advanced
//   autoscale.rs:348-364         scale_u16_to_u8 folded into the LUT
//   autoscale.rs:710-742         autoscale_db_image_tamed_synrgb_u8
//   synthetic_rgb.rs:10-67,88-178  channel LUTs
//   fast_image_resize 5.x (Cargo.toml:33) Lanczos3 coefficient tables (resize.rs:39-40)
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../include/sarpro_gpu.h"

namespace sarpro {

constexpr int kDnBins = 65536;
constexpr int kStatBins = 4096;   // autoscale.rs:103
constexpr int kClaheBins = 256;   // autoscale.rs:593
constexpr int kClaheTiles = 8;    // autoscale.rs:593
constexpr double kClaheClip = 2.0;

enum class PlanKind {
    Autoscale,        // process_scalar_data_pipeline result (U8 incl. scale_u16_to_u8, or U16)
    TamedSynRgbCopol, // autoscale_db_image_tamed_synrgb_u8(is_copol = true)
    TamedSynRgbCross, // autoscale_db_image_tamed_synrgb_u8(is_copol = false)
};

// A "value histogram": distinct sample values (as dB, with validity) and their counts.
// For u16 DN rasters the key is the DN itself (value table = dB of every u16, computed once).
struct BandPlan {
    sarpro_stats stats{};
    bool any_valid = false;
    bool clahe = false;            // lut holds 256-bin indices, blend happens on the device
    uint16_t pre_min = 0, pre_max = 0; // min/max of the quantised samples before scale_u16_to_u8
    std::vector<uint16_t> lut;     // kDnBins entries: DN -> final sample (or CLAHE bin)
    uint32_t max_present_dn = 0;   // highest DN with a non-zero count
    uint32_t sat_from_dn = 0;      // lowest present DN from which all present DNs share the brightest one's table word (low byte)
    bool have_invalid = false;     // some present DN is invalid (dB <= -50: DN 0)
    uint64_t px_total = 0, px_ge1024 = 0, px_ge2048 = 0; // pixels in all / in the bright bins (pass-A table shape of the next call)
};

// dB value of every u16 DN after the f32 cast (pipeline.rs:19-20); valid iff > -50 (pipeline.rs:22).
const double* dn_db_table();
extern double g_plan_trace_us[6]; // SARPRO_TRACE: stamps inside the last plan_from_dn_histogram* call
extern bool g_plan_trace_on;

// Plan one band from its 65,536-bin DN histogram.
void plan_from_dn_histogram(const uint64_t* hist, int bit_depth, int strategy, PlanKind kind, BandPlan* out);
// top_hint: number of leading bins that can be non-zero (brightest present DN + 1) when the caller knows it, else -1
void plan_from_dn_histogram32(const uint32_t* hist, int bit_depth, int strategy, PlanKind kind, BandPlan* out, int top_hint = -1);
// The same from the device-compacted list of non-empty bins (k_hist_total): blk = 256 {offset, count} entries, one per
// block of 256 DNs, into pairs = {dn, count} entries. false (nothing planned) when the list overflowed `cap`.
constexpr uint32_t kPresentCap = 8192; // a 400 MP GRD scene with point targets has ~5,500 distinct DNs
constexpr size_t kPresentWords = 2 * (256 + (size_t)kPresentCap); // u32 words of the block table + the pairs
bool plan_from_present_list(const uint32_t* blk, const uint32_t* pairs, uint32_t cap, int bit_depth, int strategy, PlanKind kind,
                            BandPlan* out);

// scale_u16_to_u8 (autoscale.rs:348-364) as a 65536-entry (or 256-entry) remap for given min/max.
void make_u16_to_u8_remap(uint16_t mn, uint16_t mx, int n_entries, uint8_t* remap);

// ---- general f32 rasters (polarization ops, calibrated inputs) -----------------------------
// Stats from the device-built 4096-bin histogram over [min_db, max_db] (autoscale.rs:103-159).
void stats_from_stat_histogram(const uint64_t* hist4096, uint64_t count, double min_db, double max_db, double mean_db,
                               double std_db, sarpro_stats* st);
// Window selection shared by both paths: fills low_clip / high_clip / gamma in st.
void choose_window(int strategy, PlanKind kind, sarpro_stats* st);

// ---- CLAHE tile statistics (autoscale.rs:235-302) ----------------------------------------
struct ClaheGeom {
    uint64_t rows = 0, cols = 0, tile_h = 0, tile_w = 0;
};
ClaheGeom clahe_geometry(uint64_t rows, uint64_t cols);

// ---- Lanczos3 coefficient tables ---------------------------------------------------------
struct ResampleAxis {
    uint32_t in_size = 0, out_size = 0;
    uint32_t window = 0;            // taps stored per output sample (padded with zeros)
    int precision = 0;
    std::vector<uint32_t> start;    // first source index per output sample
    std::vector<uint32_t> size;     // taps actually used per output sample
    std::vector<int32_t> coef;      // out_size * window fixed-point coefficients (fit i16 for u8 pixels)
};
// wide == false: u8 pixels (i16 coefficients, i32 accumulate); wide == true: u16 pixels (i32 / i64).
void build_lanczos3_axis(uint32_t in_size, uint32_t out_size, bool wide, ResampleAxis* ax);

// resize.rs:6-30
void calculate_resize_dimensions(size_t cols, size_t rows, size_t target, size_t* new_cols, size_t* new_rows);
// dims after the control flow of resize.rs:112-236 (resize unless long side == target, then pad)
void resize_output_dims(size_t cols, size_t rows, bool has_target, size_t target, bool pad, size_t* resized_cols,
                        size_t* resized_rows, size_t* out_cols, size_t* out_rows);

// ---- synthetic RGB LUTs ------------------------------------------------------------------
struct SynRgbLut {
    uint8_t r[256];
    uint8_t g[256];
    std::vector<uint8_t> b; // 65536, index (v1 << 8) | v2
    int floor_with_cushion = -1; // suppressed variant only
};
void build_synrgb_default_lut(SynRgbLut* lut);                       // synthetic_rgb.rs:10-50
void build_synrgb_suppressed_lut(int floor_with_cushion, SynRgbLut* lut); // synthetic_rgb.rs:115-154
// synthetic_rgb.rs:92-113: floor_with_cushion from the combined 256-bin histogram of both bands
int synrgb_floor_from_histogram(const uint32_t* hist256, uint64_t n_per_band);

} // namespace sarpro

// ---- general f32 rasters: thresholds (plan_f32.cpp) -------------------------------------------
namespace sarpro {
// dB of an f32 sample exactly as pipeline.rs:19-20 computes it
double db_of_sample(float v);
// smallest f32 with dB > -50 (pipeline.rs:22)
float valid_threshold();
// edges[k] (k = 1..4095) = smallest f32 v in [min_v, max_v] whose stat-histogram index (autoscale.rs:113-116)
// is >= k; +inf when no sample value reaches k. edges has 4096 entries, edges[0] = 0.
void build_stat_edges(float min_v, float max_v, std::vector<float>* edges);
// ---- downsample-on-read (plan_read.cpp) ----
void read_dims_for_target(uint64_t cols, uint64_t rows, uint64_t target, uint64_t* out_cols, uint64_t* out_rows, int* alg);
struct ReadAverageAxisHost {
    std::vector<int> start, end;
    std::vector<double> w_first, w_last;
};
struct ReadLanczosAxisHost {
    std::vector<int> start, count;
    std::vector<double> w;
    int window = 0;
};
void build_read_average_axis(uint64_t in, uint64_t out, ReadAverageAxisHost* a);
void build_read_lanczos_axis(uint64_t in, uint64_t out, ReadLanczosAxisHost* a);

enum class LevelKind { Quantize, TamedLinearU8, ClaheBin };
// Parameters and error bound of the device's direct fp32 evaluation of a linear-in-dB index
//   index(v) = trunc((dB(v) - low_db) / range_db * n)        (stat bins: n = 4096; quantised levels with gamma == 1)
// for samples in [min_v, max_v]: the device computes t = ((e - e0) + (lg2(m) - f0)) * scale (kernels_f32.cu) and trusts
// floor(t) only when frac(t) is at least `guard` away from both integers. guard covers: MUFU.LG2 on [1,2) (2^-22 absolute),
// the fp32 roundings of f0, of the two additions and of the product, the fp32 rounding of scale, and the reference's own f64
// roundings; times 1.5. on == false (gamma != 1, CLAHE bins, degenerate ranges): guard = 1, every sample compares thresholds.
void f32_guard(bool on, double low_db, double range_db, uint32_t n, float min_v, float max_v, int* e0, float* f0, float* scale,
               float* guard);
// edges[k] (k = 1..n_levels) = smallest valid f32 v in [min_v, max_v] whose level is >= k, where level is
//   Quantize      : cast_u16(clamp(pow((clip(db)-low)/range, gamma) * max_val, 0, max_val))   autoscale.rs:440-442
//   TamedLinearU8 : cast_u8(clamp(((clip(db)-low)/range) * 255, 0, 255))                      autoscale.rs:734-736
//   ClaheBin      : round(clamp((clip(db)-low)/range, 0, 1) * 255)                            autoscale.rs:585-587, 263
void build_level_edges(LevelKind kind, double low, double high, double gamma, uint32_t n_levels, float min_v, float max_v,
                       std::vector<float>* edges, uint32_t* level_of_min, uint32_t* level_of_max);
// Both builders place most boundaries analytically (one pow, plan_f32.cpp: analytic_threshold) and search only the rest.
// Test hook: on == false builds every entry by search; hits = entries placed analytically since process start.
void f32_edges_set_analytic(bool on);
uint64_t f32_edges_analytic_hits();
// narrow.cpp: f32 samples of a u16-valued raster -> their DN with the rule of k_f32_to_dn (not >= valid_thresh -> 0), on the
// library's host threads. false: some valid sample is not a whole number <= 65535 (dst is then unspecified).
bool narrow_f32_to_dn(const float* src, uint16_t* dst, size_t n, float valid_thresh);
} // namespace sarpro

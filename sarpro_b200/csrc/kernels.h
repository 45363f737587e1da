// kernels.h — host-callable launchers of the sm_100a kernels (definitions in kernels_*.cu).
// All launchers are asynchronous on `stream` and return the cudaError_t of the launch.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/sarpro_gpu.h"

namespace sarpro {

// ---- the plan of one band, in device memory ---------------------------------------------------
// Written by the device planner (kernels_plan.cu) or uploaded by the host planner (plan.cpp); read by the kernels downstream,
// so that no host round trip is needed between pass A and pass B. Mirrored to pinned host memory at the end of a call.
constexpr uint32_t kHmmaMaxHot = 2000; // largest DN range the shared-memory table of kernels_hmma.cu takes
struct PlanDev {
    uint32_t any_valid;      // some pixel is valid (dB > -50, pipeline.rs:22)
    uint32_t have_invalid;   // some present DN is invalid
    uint32_t max_present_dn; // brightest DN with a non-zero count
    uint32_t sat_from_dn;    // lowest present DN from which all present DNs share the brightest one's table word
    uint32_t hot;            // table range of kernels_hmma.cu (0: more than kHmmaMaxHot entries needed)
    uint32_t hot_top;        // table word of the DNs >= hot - 1
    uint32_t use_generic;    // 1: the tensor-core pass B cannot take this band; the generic exact kernel runs instead
    uint32_t clahe;
    uint32_t pre_min, pre_max; // min / max of the quantised samples before scale_u16_to_u8
    uint32_t pad0, pad1;
    unsigned long long px_total, px_ge1024, px_ge2048, pad2;
    sarpro_stats stats;
};
struct PlanParams {
    int strategy;  // sarpro_strategy
    int kind;      // 0 = autoscale, 1 = Tamed-synRGB co-pol, 2 = Tamed-synRGB cross-pol (PlanKind)
    int bit_depth; // sarpro_bit_depth
    int clahe;
};
// strategies the device planner covers (gamma == 1: no pow); the others are planned on the host
bool plan_on_device_supported(int strategy, int kind);
// total: [65536] u32 DN counts; db_table: [65536] f64 dB of every DN (plan.cpp dn_db_table); lut: [65536] u16 out
// One planner job per band; scratch: plan_scratch_bytes() of device memory (the list of present DNs when it does not fit
// shared memory). Up to two bands per launch (one CTA each).
struct PlanJob {
    const uint32_t* total;
    uint16_t* lut;
    PlanDev* out;
    void* scratch;
    PlanParams params;
};
struct PlanJobs { PlanJob j[2]; };
size_t plan_scratch_bytes();
cudaError_t launch_plan_bands(const PlanJobs& jobs, int n_bands, const double* db_table, cudaStream_t stream);

// ---- work decomposition -----------------------------------------------------------------
// A histogram work unit: rows [r0,r1) x cols [c0,c1) of the local raster, all inside one tile.
struct HistUnit {
    uint32_t r0, r1, c0, c1;
    uint32_t tile; // index into the per-tile histogram array
    uint32_t pad;
};

// ---- pass A: DN histograms ----------------------------------------------------------------
// tile_hist: [n_tiles][65536] u32, zeroed by the caller. Counts every pixel of every unit.
// counter: one zeroed device word (work-unit counter of the third-generation kernel, variants >= 20); may be null.
cudaError_t launch_dn_hist(const uint16_t* dn, uint64_t cols, const HistUnit* units_dev, uint32_t n_units,
                           uint32_t* tile_hist, int sm_count, int variant, cudaStream_t stream, uint32_t* counter = nullptr);
// total[dn] = sum_t tile_hist[t][dn]; max_dn[0] = highest DN with a non-zero total (atomicMax).
// present: 256 {offset, count} block entries followed by `cap` {dn, count} pairs (nullptr: totals only); *n_present
// must be zero before the launch
cudaError_t launch_hist_total(const uint32_t* tile_hist, uint32_t n_tiles, uint32_t* total, uint32_t* max_dn,
                              uint32_t* n_present, uint2* present, uint32_t cap,
                              cudaStream_t stream);

// ---- CLAHE tile statistics ------------------------------------------------------------------
// tile256[t][bin] = sum over dn>=1 of tile_hist[t][dn] where lut[dn] == bin (autoscale.rs:259-269)
// plan->max_present_dn bounds the walk (read on the device). The *2 variants take both bands of a pair in one launch.
cudaError_t launch_clahe_tile256(const uint32_t* tile_hist, const uint16_t* lut, uint32_t n_tiles, const PlanDev* plan,
                                 uint32_t* tile256, cudaStream_t stream);
struct ClaheStatJob {
    const uint32_t* tile_hist;
    const uint16_t* lut;
    const PlanDev* plan;
    uint32_t* tile256;
    double* cdf;
    float* cdf32;
};
struct ClaheStatJobs { ClaheStatJob j[2]; };
cudaError_t launch_clahe_tile256_2(const ClaheStatJobs& jobs, int n_bands, uint32_t n_tiles, cudaStream_t stream);
cudaError_t launch_clahe_cdf_2(const ClaheStatJobs& jobs, int n_bands, const uint64_t* tile_px, uint32_t n_tiles, cudaStream_t stream);
// clip / redistribute / CDF per tile (autoscale.rs:271-302). tile_px[t] = tile_rows*tile_cols.
cudaError_t launch_clahe_cdf(const uint32_t* tile256, const uint64_t* tile_px, uint32_t n_tiles, double* cdf,
                             float* cdf32, cudaStream_t stream);
// per-row / per-column bilinear geometry (autoscale.rs:308-318): t01 = t0 | t1 << 8
// t01 bit 7 is set when fl(omd + d) == 1.0, bit 6 when it is 1 - 2^-53; m = 2*g - tile*(2*t+1) (so d == m / (2*tile) exactly);
// sat = trunc(clamp(fl(omd + d), 0, 1) * 255): the sample of a pixel whose four CDF values are exactly 1.0 when the
// other axis has fl(omd' + d') == 1.0; sat1: the same when the other axis has 1 - 2^-53.
cudaError_t launch_clahe_axis(uint32_t n, uint32_t global_offset, uint32_t tile_size, uint32_t n_tiles, double* d,
                              double* omd, uint16_t* t01, int32_t* m, uint16_t* sat, cudaStream_t stream,
                              uint16_t* sat1 = nullptr);

struct ClaheDev {
    const double* cdf;      // [64][256]
    const float* cdf32;     // [64][256]
    const double* col_dx;   // [cols]
    const double* col_omdx; // [cols]
    const uint16_t* col_t;  // [cols]
    const double* row_dy;   // [local rows]
    const double* row_omdy;
    const uint16_t* row_t;
    const int32_t* col_m;    // [cols]  2c - tile_w*(2tx+1)
    const uint16_t* row_sat; // [local rows]
    const uint16_t* row_sat1; // [local rows]
    float inv2tw;            // 1 / (2*tile_w)
    uint32_t tile_w;         // CLAHE tile width (autoscale.rs:237)
    int tiles_x;
};

// ---- pass B: full-resolution apply ----------------------------------------------------------
// out[i] = lut[dn[i]] (u8 or u16 samples). lut has 65536 u16 entries.
cudaError_t launch_apply_lut(const uint16_t* dn, uint64_t n, const uint16_t* lut, uint8_t* out_u8, uint16_t* out_u16,
                             int sm_count, cudaStream_t stream);
// CLAHE apply: bin = lut[dn]; v = blend; out = trunc(clamp(v) * max_val); invalid (dn == 0) -> 0.
// minmax[0] = min, minmax[1] = max over all written samples (atomicMin/Max; init {65535, 0}).
cudaError_t launch_apply_clahe(const uint16_t* dn, uint32_t rows, uint32_t cols, const uint16_t* lut, ClaheDev cl,
                               int max_val, uint8_t* out_u8, uint16_t* out_u16, uint32_t* minmax, int sm_count,
                               cudaStream_t stream);
// packs (unpack = 0) / unpacks (1) {max, ~min} of two bands' {min, max} words into / from a 4-word vector
// columns [width, pitch) of every row <- the row's last real sample (re-pitched rasters, see HResizeArgs::src_width)
cudaError_t launch_pad_cols(uint16_t* dn, uint32_t rows, uint32_t width, uint32_t pitch, cudaStream_t stream);
// the whole re-pitch of a device raster in one pass (edge replication included)
cudaError_t launch_repitch(const uint16_t* src, uint16_t* dst, uint32_t rows, uint32_t width, uint32_t pitch, int sm_count,
                           cudaStream_t stream);
cudaError_t launch_minmax_pack(uint32_t* scalars0, uint32_t* scalars1, uint32_t* packed4, int unpack, cudaStream_t stream);
// Sharded scene, last exchange (comm.cu): every rank's slot of the all-gathered buffer holds its own output rows of both bands
// (max_rows x out_pitch bytes per band, row oy of rank r at local row oy - oy0[r]) and a 16-byte tail {min0, max0, min1, max1}
// with its CLAHE sample extrema. The kernel copies the rows into the two canvases (canvas row pad_top + oy), merges the extrema
// over the ranks into the bands' scalars and writes flag[0] = 1 when scale_u16_to_u8 is not the identity for some band.
struct GatherGeom {
    uint32_t world, out_pitch, max_rows, slot_bytes, pad_top, out_rows, clahe, reserved;
    uint32_t oy0[16], oy1[16];
};
cudaError_t launch_gather_unpack(const unsigned char* gathered, GatherGeom gg, unsigned char* canvas0, unsigned char* canvas1,
                                 uint32_t* scalars0, uint32_t* scalars1, uint32_t* flag, cudaStream_t stream);
cudaError_t launch_gather_tail(const uint32_t* scalars0, const uint32_t* scalars1, uint32_t* tail4, cudaStream_t stream);
// in-place u8 remap through a 256-entry table
// skip: optional device flag, the kernel returns at once when *skip != 0
cudaError_t launch_remap_u8(uint8_t* data, uint64_t n, const uint8_t* remap256, int sm_count, cudaStream_t stream,
                            const uint32_t* skip = nullptr);
// min/max of a u16 array + u16 -> u8 remap (scale_u16_to_u8, autoscale.rs:348-364)
cudaError_t launch_minmax_u16(const uint16_t* data, uint64_t n, uint32_t* minmax, int sm_count, cudaStream_t stream);
cudaError_t launch_scale_u16_to_u8(const uint16_t* data, uint64_t n, const uint32_t* minmax, uint8_t* out, int sm_count,
                                   cudaStream_t stream);

// ---- resize -------------------------------------------------------------------------------
// Device-side description of one Lanczos axis (built by plan.cpp, uploaded by the context).
struct AxisDev {
    const uint32_t* start; // [out]
    const uint32_t* size;  // [out]
    const int32_t* coef;   // [out][window]  (plain, one i32 per tap)
    const uint32_t* packed; // u8 horizontal only: [out][pairs] two i16 taps per word, tap 0 at source (start & ~3)
    uint32_t window, pairs, out_size, in_size;
    int precision;
};

enum HSrcKind { HSRC_IMAGE = 0, HSRC_DN_LUT = 1, HSRC_DN_CLAHE = 2 };

struct HResizeArgs {
    // source
    const void* src;       // u8/u16 image (HSRC_IMAGE) or u16 DN raster
    uint32_t src_rows;     // rows available in src (local)
    uint32_t src_cols;
    const uint16_t* lut;   // HSRC_DN_*
    const uint8_t* remap;  // HSRC_DN_CLAHE u8: 256-entry post-blend remap (scale_u16_to_u8) or nullptr
    ClaheDev clahe;        // HSRC_DN_CLAHE
    uint32_t* minmax;      // HSRC_DN_CLAHE: min/max of the blended samples (before remap)
    const uint32_t* skip;  // optional device flag: the kernel returns at once when *skip != 0
    const uint32_t* run_if; // optional device flag (generic kernel): the kernel returns at once when *run_if == 0
    const PlanDev* plan;   // HSRC_DN_* through kernels_hmma.cu: the band's plan (table range, kernel choice) in device memory
    // rows to produce: temp row i <- source row (row0 + i), i < n_rows
    uint32_t row0, n_rows;
    void* temp;            // [n_rows][out_cols] same pixel type
    AxisDev ax;
    // Re-pitched raster (source width not a multiple of 8: the rows were copied to a pitch of src_cols = the next multiple of 8
    // and the last 1..7 columns replicate the edge pixel, so that every row starts 16-byte aligned): the true width. 0 = src_cols.
    uint32_t src_width;
};
// One CTA of the horizontal pass owns a strip of output columns; the strip's source span is staged per row.
struct HStrip {
    uint32_t sc0;  // first staged source column (multiple of 8)
    uint32_t nvec; // staged 8-sample vectors per row
};
} // namespace sarpro
#include <vector>
namespace sarpro {
// Strip table + launch geometry for an axis (host arrays). pix16: u16 pixels. Returns the block width
// (output columns per CTA), the shared row-buffer pitch in bytes and the dynamic shared memory size.
cudaError_t hresize_build_strips(const uint32_t* start_h, const uint32_t* size_h, uint32_t out_size, uint32_t in_size,
                                 uint32_t window, uint32_t pairs, int pix16, int src_kind, uint32_t* oxb_out,
                                 std::vector<HStrip>* strips, uint32_t* rbw_out, uint32_t* smem_out);
// pix16 == 0: u8 pixels (i16 taps / i32 accumulate); 1: u16 pixels (i32 taps / i64 accumulate)
cudaError_t launch_hresize_planned(const HResizeArgs& a, int src_kind, int pix16, const HStrip* strips_dev,
                                   uint32_t n_strips, uint32_t oxb, uint32_t rbw, uint32_t smem, int sm_count,
                                   cudaStream_t stream);
// Equal-weight contiguous runs of (strip, rows) pieces, one run per persistent CTA of the tensor-core pass B. pieces_flat:
// 4 words per piece (strip, r0, r1, 0); cta_first: n_ctas + 1 entries. cuts: row positions no piece may straddle (0 ... rows);
// unit: rows a CTA takes per round (16 per warp).
void hmma_build_pieces(const std::vector<HStrip>& strips, const std::vector<uint64_t>& cuts, uint32_t n_ctas, uint32_t unit,
                       std::vector<uint32_t>* pieces_flat, std::vector<uint32_t>* cta_first, uint32_t* max_rows);
// A piece of the persistent pass-B kernels: rows [r0, r1) (inside one vertical CLAHE cell) of strip `strip`.
struct HPiece {
    uint32_t strip, r0, r1, pad;
};
// Production pass B for u8 samples (kernels_hmma.cu): horizontal Lanczos taps on the integer tensor-core path (IMMA.16832),
// samples packed straight into the A fragments. Plan: n-tiles of 8 output columns over 64-column source blocks.
struct HMmaPlanHost {
    std::vector<uint4> btab;     // permuted tap bytes, one uint4 per (n-tile block, k-step, lane)
    std::vector<int4> ntile;     // {first block, last block, offset into btab in blocks, 0}
    std::vector<uint4> strips;   // {first n-tile, end n-tile, first block, end block}
    std::vector<HStrip> weights; // per strip, for hmma_build_pieces (nvec = 8 * blocks)
    uint32_t b_bytes = 0;        // tap bytes of the largest strip (staged in shared memory)
};
// max_span: longest source span of a strip in columns (CLAHE: the tile width), 0 = unbounded. false when the axis
// does not fit the kernel (in_size not a multiple of 8, more than three n-tiles in flight, ...).
bool hmma_build_plan(const uint32_t* start_h, const uint32_t* size_h, const int32_t* coef_h, uint32_t window, uint32_t out_size,
                     uint32_t in_size, uint32_t max_span, HMmaPlanHost* plan, uint32_t max_strip_ntiles = 32);
bool hmma_replay_row(const HMmaPlanHost& plan, const uint8_t* samples, uint32_t in_size, uint32_t out_size, int precision, uint8_t* out);
size_t hmma_smem_bytes(int src_kind, uint32_t b_bytes);
uint32_t hmma_warps(bool clahe); // warps per CTA of the instantiation
cudaError_t launch_hmma(const HResizeArgs& a, int src_kind, const uint4* btab_dev, const int4* ntile_dev, const uint4* strips_dev,
                        const uint32_t* pieces_dev, const uint32_t* cta_first_dev, uint32_t n_ctas, uint32_t b_bytes,
                        cudaStream_t stream);
// vertical pass: out row oy (oy in [oy0, oy1)) from temp rows (start[oy] - temp_row0 + k)
cudaError_t launch_vresize(const void* temp, uint32_t temp_row0, uint32_t width, AxisDev ax, uint32_t oy0, uint32_t oy1,
                           void* out, uint32_t out_pitch, uint32_t out_x0, int pix16, cudaStream_t stream,
                           const uint32_t* skip = nullptr, const uint32_t* run_if = nullptr);
// scale_u16_to_u8 decision on the device (autoscale.rs:348-364 over the CLAHE samples): from minmax = {min, max}
// builds the 256-entry remap and sets skip[0] = 1 when it is the identity (min == 0 && max == 255, or no sample);
// skip[1] = 1 when, in addition, the band did not need the generic horizontal kernel (plan->use_generic == 0; plan may be null).
cudaError_t launch_clahe_remap_decide(const uint32_t* minmax, uint8_t* remap256, uint32_t* skip, cudaStream_t stream,
                                      const PlanDev* plan = nullptr);

// ---- small-image stages ---------------------------------------------------------------------
// dst (dcols x drows) zero-filled, src (scols x srows) copied at (pad_left, pad_top). elem = 1 or 2 bytes.
cudaError_t launch_pad(const void* src, uint32_t scols, uint32_t srows, void* dst, uint32_t dcols, uint32_t drows,
                       uint32_t pad_left, uint32_t pad_top, int elem, cudaStream_t stream);
cudaError_t launch_hist256_pair(const uint8_t* b1, const uint8_t* b2, uint64_t n, uint32_t* hist256,
                                cudaStream_t stream);
// suppressed-floor selection on the device: floor_idx[0] = floor_with_cushion (synthetic_rgb.rs:99-113)
cudaError_t launch_synrgb_floor(const uint32_t* hist256, uint64_t n_per_band, uint32_t* floor_idx, cudaStream_t stream);
// lut_sets: [41] sets of (r[256], g[256], b[65536]); set index = floor (suppressed) or a fixed index (default)
constexpr size_t kSynRgbSetBytes = 256 + 256 + 65536;
cudaError_t launch_synrgb(const uint8_t* b1, const uint8_t* b2, uint64_t n, const uint8_t* lut_sets,
                          const uint32_t* set_idx_dev, uint32_t fixed_set, int suppressed, uint8_t* rgb,
                          cudaStream_t stream);

// ---- f32 rasters (API boundary Array2<f32>; polarization ops) --------------------------------
// f32 -> u16 when every sample is an integer in [0,65535] (a u16 raster read as f32, gdal.rs:123);
// flag[0] is set non-zero when any sample is not. op != -1 fuses the polarization op (ops.rs:4-44).
cudaError_t launch_f32_to_dn(const float* a, const float* b, int op, uint64_t n, float valid_thresh, uint16_t* dn,
                             uint32_t* flag, int sm_count, cudaStream_t stream);
cudaError_t launch_pol_op(const float* a, const float* b, int op, uint64_t n, float* out, int sm_count,
                          cudaStream_t stream);
// general f32 path (kernels_f32.cu; launchers declared in api_f32.cu): min/max/count over valid samples.
// valid <=> v >= valid_thresh (the smallest f32 with dB > -50); valid samples are positive, so the bit
// pattern orders like the value.
// ---- downsample-on-read (kernels_read.cu; tables from plan_read.cpp) ----
// Average: output index d covers source [start[d], end[d]) with the first / last sample weighted by its covered fraction
struct ReadAvgAxis {
    const int* start;
    const int* end;
    const double* w_first;
    const double* w_last;
};
// Lanczos: output index d = sum_k src[start[d] + k] * w[d * window + k], k < count[d]
struct ReadConvAxis {
    const int* start;
    const int* count;
    const double* w;
    int window;
};
cudaError_t launch_read_average(const void* src, int src_u16, uint32_t rows, uint32_t cols, const ReadAvgAxis& ax, const ReadAvgAxis& ay,
                                float* out, uint32_t out_rows, uint32_t out_cols, cudaStream_t stream);
cudaError_t launch_read_lanczos(const void* src, int src_u16, uint32_t rows, uint32_t cols, const ReadConvAxis& ax, const ReadConvAxis& ay,
                                double* tmp, float* out, uint32_t out_rows, uint32_t out_cols, cudaStream_t stream);
// Parameters of the guarded direct index of the general f32 path (kernels_f32.cu f32_guarded_index; built by f32_guard in
// plan_f32.cpp). guard >= 0.5: the shortcut is off and every sample takes the threshold comparison.
struct F32GuardHost {
    int e0;
    float f0, scale, guard;
};
struct F32Scan {
    uint32_t min_key, max_key; // bit patterns of the min / max valid sample
    unsigned long long valid_count;
};
// dB plane + mask (pipeline.rs:8-40) for callers that want them (device log10; see DESIGN.md tolerance)
cudaError_t launch_db_mask(const float* v, uint64_t n, double* db, uint8_t* mask, int sm_count, cudaStream_t stream);

} // namespace sarpro

/*
 * sarpro_gpu.h — C ABI of libsarpro_gpu.so: the B200 (sm_100a) implementation of SARPRO's
 * per-pixel raster hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * The reference (bogwi/sarpro, Rust) has no FFI/plugin interface for this path; the seam is the
 * set of pure functions in src/core/processing (pipeline, autoscale, ops, resize, padding, synthetic_rgb).  Each entry point below names the reference
 * function (file:line, relative to the reference root) whose body it replaces.  The Rust-side
 * binding (`sarpro-gpu-sys`) a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions (mirroring the reference):
 *   - inputs are borrowed, outputs are caller-allocated; the library never frees caller memory;
 *   - every call is synchronous and returns 0 on success or a negative sarpro_status;
 *     sarpro_last_error(ctx) then holds the message the Rust side wraps in Error::External(String)
 *     (src/error.rs:43-47).  No exception or abort crosses the boundary;
 *   - a sarpro_ctx is single-threaded (the reference calls this path from one thread at a time);
 *     distinct contexts may be used concurrently;
 *   - enum discriminants equal the Rust declaration order;
 *   - pointers are HOST pointers unless the parameter is named *_dev or the descriptor says
 *     SARPRO_LOC_DEVICE.  Pinned host memory (sarpro_host_alloc) makes the copies asynchronous.
 *     Device buffers are read and written on the context's stream (sarpro_ctx_set_stream, or the
 *     library's own non-blocking stream): work that produces them on another stream must have
 *     completed (or that stream must be the context's) before the call; every call returns with its
 *     outputs complete.
 *   - there is NO CPU fallback: without a CUDA device sarpro_ctx_create fails with
 *     SARPRO_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef SARPRO_GPU_H
#define SARPRO_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SARPRO_GPU_ABI_VERSION 1

/* ---- enums (discriminants = Rust declaration order) ------------------------------------ */
/* AutoscaleStrategy, src/types.rs:115-123 */
typedef enum sarpro_strategy {
    SARPRO_STRATEGY_STANDARD = 0, SARPRO_STRATEGY_ROBUST = 1, SARPRO_STRATEGY_ADAPTIVE = 2,
    SARPRO_STRATEGY_EQUALIZED = 3, SARPRO_STRATEGY_CLAHE = 4, SARPRO_STRATEGY_TAMED = 5,
    SARPRO_STRATEGY_DEFAULT = 6
} sarpro_strategy;
/* BitDepth, src/types.rs:170-173 */
typedef enum sarpro_bit_depth { SARPRO_U8 = 0, SARPRO_U16 = 1 } sarpro_bit_depth;
/* PolarizationOperation, src/types.rs:8-14 */
typedef enum sarpro_pol_operation {
    SARPRO_OP_SUM = 0, SARPRO_OP_DIFF = 1, SARPRO_OP_RATIO = 2, SARPRO_OP_NDIFF = 3, SARPRO_OP_LOGRATIO = 4,
    SARPRO_OP_NONE = -1
} sarpro_pol_operation;
/* SyntheticRgbMode, src/types.rs:177-182 (all four alias Default, synthetic_rgb.rs:72-79) */
typedef enum sarpro_synrgb_mode {
    SARPRO_SYNRGB_DEFAULT = 0, SARPRO_SYNRGB_RGB_RATIO = 1, SARPRO_SYNRGB_SAR_URBAN = 2, SARPRO_SYNRGB_ENHANCED = 3
} sarpro_synrgb_mode;
/* OutputFormat, src/types.rs:162-165 (JPEG forces U8, save.rs:121,321) */
typedef enum sarpro_output_format { SARPRO_FORMAT_TIFF = 0, SARPRO_FORMAT_JPEG = 1 } sarpro_output_format;

typedef enum sarpro_dtype { SARPRO_DT_F32 = 0, SARPRO_DT_U16 = 1 } sarpro_dtype;
typedef enum sarpro_location {
    SARPRO_LOC_HOST = 0,
    SARPRO_LOC_DEVICE = 1,
    SARPRO_LOC_NONE = 2 /* outputs only: nothing is copied out, the result stays in the context (see sarpro_encode_last_jpeg) */
} sarpro_location;

typedef enum sarpro_status {
    SARPRO_OK = 0,
    SARPRO_ERR_INVALID_ARGUMENT = -1,
    SARPRO_ERR_NO_DEVICE = -2,      /* no CUDA device / driver: the library has no CPU path */
    SARPRO_ERR_CUDA = -3,
    SARPRO_ERR_OUT_OF_MEMORY = -4,
    SARPRO_ERR_U16_REQUIRED = -5,   /* "U16 data required for U16 bit depth" resize.rs:162,221 padding.rs:36 */
    SARPRO_ERR_TOO_LARGE = -6,
    SARPRO_ERR_COMM = -7,
    SARPRO_ERR_INTERNAL = -8
} sarpro_status;

/* ---- plain structs --------------------------------------------------------------------- */
/* HistogramStats (autoscale.rs:7-24) + the chosen window, so the host can emit the reference's
 * info!/debug! lines (autoscale.rs:398-401, 431-434, 485-488, 566-569). */
typedef struct sarpro_stats {
    uint64_t valid_count;
    double min_db, max_db, mean_db, std_db, median_db;
    double p01, p02, p05, p10, p25, p75, p90, p95, p98, p99;
    double low_clip, high_clip, gamma;
} sarpro_stats;

/* tuple returned by resize_image_data_with_meta, resize.rs:99-110 */
typedef struct sarpro_resize_meta {
    uint64_t cols, rows;
    double scale_x, scale_y;
    uint64_t pad_left, pad_top;
} sarpro_resize_meta;

/* One input band: a row-major rows x cols raster (Array2<f32> at the reference boundary,
 * gdal.rs:123-131; or the raw u16 DN the f32 was read from). */
typedef struct sarpro_band {
    const void* data;
    int32_t dtype;     /* sarpro_dtype */
    int32_t location;  /* sarpro_location */
    uint64_t rows, cols;
} sarpro_band;

/* One output raster, caller-allocated. channels: 1 (gray) or 3 (interleaved RGB). */
typedef struct sarpro_image {
    void* data;              /* u8 or u16 samples */
    int32_t location;        /* sarpro_location */
    int32_t bit_depth;       /* sarpro_bit_depth of `data` */
    uint64_t capacity_bytes; /* size of the caller's buffer; checked */
    uint64_t cols, rows;     /* filled by the library */
    int32_t channels;        /* filled by the library */
    int32_t reserved;
    sarpro_resize_meta meta; /* filled by the library */
} sarpro_image;

/* Time spent on the device for the last pipeline call (CUDA events on the ctx stream). */
typedef struct sarpro_timing {
    float total_ms;      /* first kernel/copy -> last kernel/copy on the ctx stream */
    float h2d_ms, d2h_ms;
    float kernel_ms;     /* total_ms minus copies when they are serialised */
    uint32_t kernel_launches;
    uint32_t host_syncs; /* planner round trips */
    uint64_t h2d_bytes, d2h_bytes;
    /* per-stage device time of the last call (CUDA events around the launches, summed over bands). Pass A, pass B and the
       collectives are timed by default; the environment variable SARPRO_STAGE_TIMING=all times every launch (each event pair
       costs a few microseconds of host time per call). */
    float stage_ms[8];        /* indexed by sarpro_stage */
    uint32_t stage_launches[8];
} sarpro_timing;

typedef enum sarpro_stage {
    SARPRO_STAGE_HIST = 0,     /* pass A: DN / value histograms */
    SARPRO_STAGE_PLAN = 1,     /* histogram totals, CLAHE tile statistics, table uploads */
    SARPRO_STAGE_APPLY = 2,    /* pass B: LUT / CLAHE apply (+ fused horizontal Lanczos) */
    SARPRO_STAGE_VRESIZE = 3,  /* vertical Lanczos */
    SARPRO_STAGE_RGB = 4,      /* synthetic RGB composition */
    SARPRO_STAGE_CONVERT = 5,  /* f32 -> DN bridge, polarization ops */
    SARPRO_STAGE_COMM = 6,     /* collectives */
    SARPRO_STAGE_OTHER = 7
} sarpro_stage;

typedef struct sarpro_ctx sarpro_ctx;

/* ---- lifecycle ------------------------------------------------------------------------- */
int sarpro_abi_version(void);
/* Creates a context on CUDA device `device_id`. Fails (SARPRO_ERR_NO_DEVICE) without a GPU. */
int sarpro_ctx_create(sarpro_ctx** out, int device_id);
void sarpro_ctx_destroy(sarpro_ctx* ctx);
/* Message for the last failing call on ctx (ctx==NULL: last sarpro_ctx_create failure on this thread). */
const char* sarpro_last_error(const sarpro_ctx* ctx);
/* Run the context on a caller-owned CUDA stream (cudaStream_t) instead of its own. */
int sarpro_ctx_set_stream(sarpro_ctx* ctx, void* cuda_stream);
int sarpro_ctx_synchronize(sarpro_ctx* ctx);
int sarpro_last_timing(const sarpro_ctx* ctx, sarpro_timing* out);
/* Pinned host memory for inputs/outputs (optional; pageable pointers also work). */
int sarpro_host_alloc(void** out, size_t bytes);
void sarpro_host_free(void* p);
int sarpro_host_register(void* p, size_t bytes);
void sarpro_host_unregister(void* p);

/* ---- stage level: 1:1 with the reference functions ------------------------------------- */
/* ops.rs:4-44  sum_arrays / difference_arrays / ratio_arrays / normalized_diff_arrays / log_ratio_arrays */
int sarpro_pol_op(sarpro_ctx* ctx, int op, const float* a, const float* b, size_t rows, size_t cols, float* out);

/* pipeline.rs:42-66  process_scalar_data_pipeline (dB + mask + autoscale dispatch + bit depth,
 * autoscale.rs:662-704). The f64 dB plane and mask are not materialised (callers only use .dim(),
 * save.rs:53,122,201,329). out_u8 is filled for U8, out_u16 for U16; the other may be NULL. */
int sarpro_process_scalar_data_pipeline(sarpro_ctx* ctx, const float* v, size_t rows, size_t cols, int bit_depth,
                                        int strategy, uint8_t* out_u8, uint16_t* out_u16, sarpro_stats* stats);
/* Same result as casting the u16 DN raster to f32 first (gdal.rs:123) and calling the above. */
int sarpro_process_dn_pipeline(sarpro_ctx* ctx, const uint16_t* dn, size_t rows, size_t cols, int bit_depth,
                               int strategy, uint8_t* out_u8, uint16_t* out_u16, sarpro_stats* stats);
/* pipeline.rs:8-40  process_scalar_data_inplace (only for callers that really want the planes) */
int sarpro_process_scalar_data_inplace(sarpro_ctx* ctx, const float* v, size_t rows, size_t cols, double* db,
                                       uint8_t* valid_mask);
/* autoscale.rs:710-742  autoscale_db_image_tamed_synrgb_u8 (takes the linear band, not the dB plane) */
int sarpro_autoscale_tamed_synrgb_u8(sarpro_ctx* ctx, const float* v, size_t rows, size_t cols, int is_copol,
                                     uint8_t* out);
/* autoscale.rs:348-364  scale_u16_to_u8 */
int sarpro_scale_u16_to_u8(sarpro_ctx* ctx, const uint16_t* data, size_t n, uint8_t* out);

/* resize.rs:6-30 calculate_resize_dimensions + the control flow of resize.rs:112-236 + padding.rs:12
 * (pure host arithmetic; no context needed): dims of the buffer resize_image_data_with_meta returns. */
int sarpro_resize_output_dims(size_t cols, size_t rows, int has_target, size_t target, int pad, size_t* out_cols,
                              size_t* out_rows);
/* resize.rs:91-236 resize_image_data_with_meta (Lanczos3 of resize.rs:32-89 + padding.rs:5-49).
 * u8_data is used for U8, u16_data for U16 (NULL -> SARPRO_ERR_U16_REQUIRED). */
int sarpro_resize_image_data_with_meta(sarpro_ctx* ctx, const uint8_t* u8_data, const uint16_t* u16_data,
                                       size_t cols, size_t rows, int has_target, size_t target, int bit_depth,
                                       int pad, uint8_t* out_u8, uint16_t* out_u16, sarpro_resize_meta* meta);
/* padding.rs:5-49 add_padding_to_square */
int sarpro_add_padding_to_square(sarpro_ctx* ctx, const uint8_t* u8_data, const uint16_t* u16_data, size_t cols,
                                 size_t rows, int bit_depth, uint8_t* out_u8, uint16_t* out_u16);
/* synthetic_rgb.rs:182-197 create_synthetic_rgb_by_mode_and_strategy (-> :10-67 or :88-178) */
int sarpro_create_synthetic_rgb_by_mode_and_strategy(sarpro_ctx* ctx, int mode, int strategy, const uint8_t* band1,
                                                     const uint8_t* band2, size_t n, uint8_t* rgb);

/* ---- fused pipelines: what save.rs / api/mod.rs run per product; data stays in HBM ------ */
/* Single band (optionally a polarization op of two bands, sentinel1.rs:1497-1579 -> ops.rs):
 * save.rs:49-65 / 119-134 and api/mod.rs:84-130, 250-281, 284-369.
 * b is ignored when op == SARPRO_OP_NONE. JPEG forces U8. */
int sarpro_pipeline_single(sarpro_ctx* ctx, const sarpro_band* a, const sarpro_band* b, int op, int format,
                           int bit_depth, int strategy, int has_target, size_t target, int pad, sarpro_image* out,
                           sarpro_stats* stats);
/* Two-band TIFF: save.rs:199-316, api/mod.rs:133-200 (independent statistics per band). */
int sarpro_pipeline_multiband_tiff(sarpro_ctx* ctx, const sarpro_band* b1, const sarpro_band* b2, int bit_depth,
                                   int strategy, int has_target, size_t target, int pad, sarpro_image* out1,
                                   sarpro_image* out2, sarpro_stats* stats2 /* [2] or NULL */);
/* Synthetic-RGB JPEG: save.rs:317-368 (tamed_band_step != 0: band-specific Tamed autoscale of
 * save.rs:324-328,347-351) or api/mod.rs:203-247 (tamed_band_step == 0). out is interleaved RGB u8. */
int sarpro_pipeline_synrgb(sarpro_ctx* ctx, const sarpro_band* b1, const sarpro_band* b2, int strategy, int mode,
                           int has_target, size_t target, int pad, int tamed_band_step, sarpro_image* out,
                           sarpro_stats* stats2 /* [2] or NULL */);

/* ---- multi-GPU: one process per GPU, a scene row-band-sharded across ranks --------------- */
/* The library dlopen()s libnccl.so.2 and runs its own small collectives (DN-histogram and CLAHE
 * tile-histogram all-reduce, min/max all-reduce, gather of the resized rows) on the ctx stream.
 * The host only ships the 128-byte ncclUniqueId between ranks (any transport). */
int sarpro_comm_unique_id(void* out128);
int sarpro_comm_init(sarpro_ctx* ctx, const void* unique_id128, int rank, int world);
int sarpro_comm_destroy(sarpro_ctx* ctx);
/* Row band [*r0,*r1) owned by `rank`; with clahe != 0 the edges are aligned to the CLAHE tile
 * height ceil(rows/8) (autoscale.rs:235). Pure host arithmetic. */
int sarpro_shard_rows(size_t rows, int world, int rank, int clahe, size_t* r0, size_t* r1);
/* Source rows [*h0,*h1) a rank must hold to produce its share of a resize to `target` (its own
 * band plus the vertical Lanczos halo). Pure host arithmetic. */
int sarpro_shard_halo_rows(size_t rows, size_t cols, int has_target, size_t target, int world, int rank, int clahe,
                           size_t* h0, size_t* h1);
/* Sharded synthetic-RGB pipeline. b1/b2 describe THIS rank's rows [h0,h1) of the scene
 * (rows field = h1-h0); scene_rows is the full height. The full output lands on rank 0. */
int sarpro_pipeline_synrgb_sharded(sarpro_ctx* ctx, const sarpro_band* b1, const sarpro_band* b2, size_t scene_rows,
                                   int strategy, int mode, int has_target, size_t target, int pad,
                                   int tamed_band_step, sarpro_image* out);

/* One or two polarization operations over the same pair (a, b), each autoscaled on its own into a full-resolution band (BASELINE
 * config 4: log-ratio and normalised difference of VV / VH -> Equalized -> two u16 bands; the calls the reference makes one after
 * the other, sentinel1.rs:1497-1579 -> ops.rs:4-44 -> pipeline.rs:42-66 -> save.rs:199-316 without a target size). The operands
 * (f32, or the raw u16 DN) are read once per pass for both operations: 3 x (a + b) reads + one write per band. `outs` / `stats` hold
 * n_ops entries; device-resident outputs are written in place.
 * scene_rows == 0 or == a->rows: the bands are the whole scene. Otherwise the call is ONE RANK's contiguous row band of a scene of
 * scene_rows rows (sarpro_comm_init first; any split of the rows): the library merges the scan (min / max / valid count) and the
 * 4096-bin stat histograms over the ranks (integers and bit patterns: identical to the unsharded values), every rank derives the
 * same windows, and the rank's rows of each result are written to outs[k]. All strategies except CLAHE (whose tile statistics need
 * the scene geometry: sarpro_pipeline_single per operation, unsharded). */
int sarpro_pipeline_polops(sarpro_ctx* ctx, const sarpro_band* a, const sarpro_band* b, size_t scene_rows, int n_ops, const int* ops,
                           int bit_depth, int strategy, sarpro_image* outs, sarpro_stats* stats);
/* n_ops == 1 form of the above. */
int sarpro_pipeline_single_sharded(sarpro_ctx* ctx, const sarpro_band* a, const sarpro_band* b, size_t scene_rows, int op, int bit_depth,
                                   int strategy, sarpro_image* out, sarpro_stats* stats);

/* ---- downsample-on-read (the reader's --size flow, sentinel1.rs:1074-1109 -> gdal.rs:145-177) ---------------------- */
typedef enum sarpro_resample { SARPRO_RESAMPLE_AVERAGE = 0, SARPRO_RESAMPLE_LANCZOS = 1 } sarpro_resample;
/* Output shape for a long-side target (aspect preserved, never enlarged) and the resampler the reader picks when the user
 * gives none: Average for a reduction of 4 or more, Lanczos below (sentinel1.rs:1083-1102). Host only. */
int sarpro_read_dims_for_target(size_t cols, size_t rows, size_t target, size_t* out_cols, size_t* out_rows, int* alg);
/* GdalSarReader::read_band_resampled (gdal.rs:145-177): the raster (u16 DN as stored in the GRD TIFF, or f32) resampled to
 * out_rows x out_cols f32 samples, host or device memory (out_location). The arithmetic is GDAL's RasterIO resampling, which
 * lives in the system libgdal (version unpinned) and not in the reference tree: restated from the published algorithm of
 * GDAL >= 3.3 - parity unpinned. */
int sarpro_read_band_resampled(sarpro_ctx* ctx, const sarpro_band* in, size_t out_cols, size_t out_rows, int alg, float* out,
                               int out_location);

/* ---- encoder hand-off (writers) ---------------------------------------------------------------------------------------- */
/* write_gray_jpeg / write_rgb_jpeg (io/writers/jpeg.rs:6-30; jpeg_encoder at quality 100: baseline, 4:4:4) on the GPU (nvJPEG,
 * loaded at run time): a u8 gray (channels 1) or interleaved RGB (channels 3) image in host or device memory -> a complete JPEG
 * stream in `out` (host memory). out == NULL: only *out_bytes (the stream length) is returned. The stream is a valid baseline
 * JPEG of the same pixels, not the byte sequence jpeg_encoder would write. */
int sarpro_encode_jpeg(sarpro_ctx* ctx, const sarpro_image* img, int quality, void* out, size_t capacity, size_t* out_bytes);
/* The same for the u8 result the LAST pipeline call on this context left on the device: which = 0 the interleaved RGB of
 * sarpro_pipeline_synrgb, 1 / 2 the gray band(s) of sarpro_pipeline_single / _multiband_tiff / _synrgb. With an output image of
 * location SARPRO_LOC_NONE in the pipeline call, the raw pixels never cross PCIe: only the JPEG stream does. For the GeoTIFF
 * writers (io/writers/tiff.rs:6-78) the hand-off is a pinned output buffer (sarpro_host_alloc) the raster is delivered into. */
int sarpro_encode_last_jpeg(sarpro_ctx* ctx, int which, int quality, void* out, size_t capacity, size_t* out_bytes);

/* ---- batch (BASELINE config 5) ------------------------------------------------------------ */
/* One scene of a batch: the band pair of one product. A scene whose b1.data is NULL (or that has no pixels) is counted as
 * skipped, like a product SafeReader::open_with_warnings_with_options returns None for (api/mod.rs:498-529). */
typedef struct sarpro_scene {
    sarpro_band b1, b2;
} sarpro_scene;
typedef enum sarpro_batch_kind {
    SARPRO_BATCH_MULTIBAND = 0, /* two gray bands per scene (save.rs:199-316): outs[2k], outs[2k+1]; stats[2k], stats[2k+1] */
    SARPRO_BATCH_SYNRGB = 1     /* one interleaved RGB image per scene (save.rs:317-368): outs[k]; stats[2k], stats[2k+1] */
} sarpro_batch_kind;
/* BatchReport (api/mod.rs:451-457) */
typedef struct sarpro_batch_report {
    uint64_t processed, skipped, errors;
} sarpro_batch_report;
/* The scene loop of process_directory_to_path (api/mod.rs:474-536; cli/runner.rs:294-345) over rasters that are already
 * decoded: every scene goes through sarpro_pipeline_multiband_tiff or sarpro_pipeline_synrgb with the same parameters. Host
 * rasters are staged through two device slots on a copy stream of their own, so scene k+1 uploads while scene k computes and
 * its result goes back (pinned host memory - sarpro_host_alloc / sarpro_host_register - is what makes the copies asynchronous).
 * continue_on_error != 0: a failing scene is counted in report->errors, its status goes to statuses[k] (may be NULL) and the
 * loop goes on; otherwise the first error is returned. bit_depth is ignored for SARPRO_BATCH_SYNRGB (JPEG forces U8). */
int sarpro_pipeline_batch(sarpro_ctx* ctx, const sarpro_scene* scenes, size_t n, int kind, int bit_depth, int strategy, int mode,
                          int has_target, size_t target, int pad, int tamed_band_step, int continue_on_error, sarpro_image* outs,
                          sarpro_stats* stats, int* statuses, sarpro_batch_report* report);

/* ---- host-only planner entry points (pure CPU; used by tests and by multi-rank hosts) ---- */
/* Statistics + window + DN->sample LUT from a 65,536-bin DN histogram, i.e. what
 * compute_histogram_stats (autoscale.rs:35-160) + autoscale_db_image[_advanced] (:368-659) +
 * scale_u16_to_u8 (:348-364) yield for a u16 raster with that histogram. lut16 has 65536 entries.
 * For CLAHE the LUT holds the 256-bin index of autoscale.rs:263 (the blend needs pixel positions). */
int sarpro_plan_from_dn_histogram(const uint64_t* hist65536, int bit_depth, int strategy, sarpro_stats* stats,
                                  uint16_t* lut16);
/* The same plan from the list of non-empty DN bins in the layout the device writes for the planner (k_hist_total): blocks =
 * 256 {offset, count} pairs of u32, one per block of 256 consecutive DNs, into pairs = {dn, count} entries of u32; the DNs of a
 * block are ascending, the blocks may sit in pairs[] in any order. Returns 1 and fills stats / lut16 like the call above, or 0
 * (nothing written) when some block's offset + count exceeds cap, i.e. the list overflowed and the caller must use the dense
 * histogram. */
int sarpro_plan_from_present_list(const uint32_t* blocks, const uint32_t* pairs, uint32_t cap, int bit_depth, int strategy,
                                  sarpro_stats* stats, uint16_t* lut16);

/* Test hook: the DEVICE planner (kernels_plan.cu: the single-CTA restatement of the host planner for the strategies without
 * pow(): Robust, Equalized, Clahe, Tamed, Default and the Tamed-synRGB windows) on a 65,536-bin DN histogram given as u32
 * counts in host memory. plan_kind: 0 = autoscale, 1 / 2 = autoscale_db_image_tamed_synrgb_u8 co- / cross-pol. Fills stats and
 * lut16 like sarpro_plan_from_dn_histogram, and hot2 = {table range, table word of the saturated DNs} of the tensor-core pass B.
 * Returns SARPRO_ERR_INVALID_ARGUMENT for a strategy the device planner does not cover. */
int sarpro_plan_on_device(sarpro_ctx* ctx, const uint32_t* hist65536, int bit_depth, int strategy, int plan_kind,
                          sarpro_stats* stats, uint16_t* lut16, uint32_t* hot2);
/* The host planner on the same inputs (what the device planner must reproduce bit for bit, mean / std within 1e-12):
 * sarpro_plan_from_dn_histogram with the plan kind and the table range as extra outputs. No GPU needed. */
int sarpro_plan_kind_from_dn_histogram(const uint64_t* hist65536, int bit_depth, int strategy, int plan_kind,
                                       sarpro_stats* stats, uint16_t* lut16, uint32_t* hot2);

/* Horizontal Lanczos3 pass of one row of u8 samples (resize.rs:39-50, first pass of the crate's separable resize) computed
 * twice on the host: directly from the fixed-point taps (out_direct) and by replaying the tensor-core kernel's plan — strips,
 * n-tile slots, k-step windows and the permuted hi/lo tap bytes of its B fragments — in the device's order (out_replay).
 * Returns 1 when the plan exists for this axis (0: the axis falls back to the other kernels, out_replay untouched).
 * max_span: longest strip in source columns (the CLAHE tile width), 0 = unbounded. strip_ntiles: n-tiles (8 output columns) per
 * strip, 0 = the default of 32 (a rank's band of a sharded scene uses shorter strips, see choose_strip_nt in api.cu). */
int sarpro_lanczos_row_plan_check(const uint8_t* samples, size_t in_size, size_t out_size, size_t max_span, size_t strip_ntiles,
                                  uint8_t* out_direct, uint8_t* out_replay);

/* Test hook (host only): one row of u16 samples through the Lanczos3 table of the u16 kernels (i32 taps, i64 accumulate: the
 * crate's U16 convolution, resize.rs:62-81) in the kernels' arithmetic. The CPU tests compare it with the oracle. */
int sarpro_lanczos_row_check_u16(const uint16_t* samples, size_t in_size, size_t out_size, uint16_t* out);

/* Test hook (host only): one row of u16 samples through the downsample-on-read tables the kernels use (plan_read.cpp: spans and
 * weights of GDAL's Average / Lanczos resampling, see sarpro_read_band_resampled), accumulated in f64 in the kernels' order.
 * out has out_size f32 samples. No GPU needed: the CPU tests compare it with the oracle's restatement. */
int sarpro_read_row_plan_check(const uint16_t* samples, size_t in_size, size_t out_size, int alg, float* out);

/* Test hook (host only): parameters and error bound of the guarded direct index the general f32 kernels use for the indices
 * that are linear in dB, trunc((10 log10 v - low_db) / range_db * n) (stat bins autoscale.rs:113-116; quantised levels with
 * gamma == 1, :440-442 / :649-651 / :732-734). The device evaluates t = ((e - e0) + (lg2(m) - f0)) * scale in fp32 (v = m 2^e,
 * MUFU.LG2 on m) and takes floor(t) only when frac(t) lies in [guard, 1 - guard]; every other sample compares thresholds.
 * guard == 1: the shortcut is off. */
int sarpro_f32_guard_params(double low_db, double range_db, uint32_t n, float min_v, float max_v, int* e0, float* f0, float* scale,
                            float* guard);

/* Test hook (host only, not thread-safe against concurrent calls of the library): the threshold tables of the general f32 path
 * (smallest f32 reaching each stat bin, kind = -1, autoscale.rs:113-116; or each quantised level: kind 0 = autoscale.rs:440-442,
 * 1 = the Tamed u8 levels :732-734, 2 = the CLAHE bins :585-587), built twice: the way the pipelines build them (boundaries
 * placed analytically where the f64 expression provably cannot disagree, plan_f32.cpp) and by bracketing every boundary with
 * the reference's own f64 expression. Returns the table length (4096 or n_levels + 1); n_analytic = entries placed
 * analytically, n_mismatch = entries whose bit patterns differ (must be 0); edges_out (optional) receives the table. */
int sarpro_f32_edges_check(int kind, double low_db, double high_db, double gamma, uint32_t n_levels, float min_v, float max_v,
                           uint32_t* n_analytic, uint32_t* n_mismatch, float* edges_out);

/* Test hook (host only): one of the 42 channel-LUT sets the synthetic-RGB kernels read (set 0..40: the suppressed variant for that
 * floor_with_cushion, synthetic_rgb.rs:115-154; 41: the default variant, :10-50; -1: the suppressed set chosen from the combined
 * 256-bin histogram of both u8 bands, :92-113, the rule k_synrgb_floor applies on the device). lut_r / lut_g: 256 entries,
 * lut_b: 65536 entries indexed (band1 << 8) | band2. The CPU tests compose every (band1, band2) pair through them and
 * compare with the oracle. */
int sarpro_synrgb_lut_check(int set, const uint32_t* hist256, uint64_t n_per_band, int* floor_with_cushion, uint8_t* lut_r, uint8_t* lut_g,
                            uint8_t* lut_b);

/* Test hook (host only): the host step of the general f32 path between its histogram pass and its quantisation pass (api_f32.cu):
 * percentiles from the 4096-bin stat histogram over [dB(min_v), dB(max_v)] (autoscale.rs:103-160) and the strategy's window
 * (autoscale.rs:404-428, 492-562; tamed_synrgb_kind 1 / 2 = the co- / cross-pol windows of :719-728, 0 = the strategy's).
 * Together with sarpro_f32_edges_check the CPU tests replay the whole path (numpy stands in for the comparing kernels) against
 * the oracle. */
int sarpro_plan_from_stat_histogram(const uint64_t* hist4096, uint64_t valid_count, float min_v, float max_v, double mean_db, double std_db,
                                    int strategy, int tamed_synrgb_kind, sarpro_stats* stats);

/* Test hook (host only): the host-side narrowing the pipelines apply to large f32 HOST rasters on their way to the device
 * (the reference's boundary is the f32 raster GDAL made of a u16 band, gdal.rs:123; narrowing halves the PCIe bytes). dst[i]
 * = the DN of src[i]: 0 for samples that are not valid (negative, NaN, <= -50 dB: pipeline.rs:19-22), the sample itself
 * otherwise. u16_valued = 0 when some valid sample is not a whole number <= 65535: such rasters are uploaded as f32 and take
 * the general path (dst is then unspecified). Same rule as the device kernel that narrows f32 DEVICE rasters. */
int sarpro_narrow_f32_check(const float* src, size_t n, uint16_t* dst, int* u16_valued);

#ifdef __cplusplus
}
#endif
#endif /* SARPRO_GPU_H */

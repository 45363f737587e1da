/*
 * oracle.h — C interface of the CPU ORACLE for the SARPRO per-pixel raster path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (sarpro_b200/, include/sarpro_gpu.h) never
 * links, imports or calls it; the product fails loudly when its CUDA library is missing.
 *
 * What it is: a serial, line-by-line C++ restatement of the reference's Rust functions in
 *   src/core/processing/{pipeline,autoscale,ops,resize,padding,synthetic_rgb}.rs
 * and of the call orders in src/core/processing/save.rs and src/api/mod.rs.
 * Every function below cites the reference file:line it follows.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
 * path, and it cannot be compiled in this image (no Rust toolchain, no
 * GDAL), so this restatement could not be checked against reference outputs.  It is
 * cross-checked against an independent numpy re-derivation (tests/ref_numpy.py) and, for
 * the Lanczos stage, against Pillow (sanity only).  The Lanczos3 arithmetic lives in the
 * un-vendored third-party crate fast_image_resize (Cargo.toml:33, "5.2.1", Cargo.lock not
 * committed); its published algorithm (a Rust port of Pillow-SIMD's fixed-point separable
 * convolution) is restated in oracle_resize.cpp.
 *
 * Enum discriminants equal the Rust declaration order:
 *   AutoscaleStrategy  src/types.rs:115-123   BitDepth src/types.rs:170-173
 *   PolarizationOperation src/types.rs:8-14   SyntheticRgbMode src/types.rs:177-182
 */
#ifndef SARPRO_ORACLE_H
#define SARPRO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_STRATEGY_STANDARD = 0, ORACLE_STRATEGY_ROBUST, ORACLE_STRATEGY_ADAPTIVE,
       ORACLE_STRATEGY_EQUALIZED, ORACLE_STRATEGY_CLAHE, ORACLE_STRATEGY_TAMED,
       ORACLE_STRATEGY_DEFAULT };
enum { ORACLE_U8 = 0, ORACLE_U16 = 1 };
enum { ORACLE_OP_SUM = 0, ORACLE_OP_DIFF, ORACLE_OP_RATIO, ORACLE_OP_NDIFF, ORACLE_OP_LOGRATIO };
enum { ORACLE_FORMAT_TIFF = 0, ORACLE_FORMAT_JPEG = 1 };

/* HistogramStats, autoscale.rs:7-24 (field order kept) plus the chosen window for logging parity. */
typedef struct oracle_stats {
    uint64_t valid_count;
    double min_db, max_db, mean_db, std_db, median_db;
    double p01, p02, p05, p10, p25, p75, p90, p95, p98, p99;
    double low_clip, high_clip, gamma; /* filled by the autoscale entry points */
} oracle_stats;

/* resize.rs:91-111 return tuple */
typedef struct oracle_resize_meta {
    uint64_t cols, rows;
    double scale_x, scale_y;
    uint64_t pad_left, pad_top;
} oracle_resize_meta;

/* pipeline.rs:8-40 */
void oracle_process_scalar_data_inplace(const float* v, size_t n, double* db, uint8_t* mask);
/* autoscale.rs:35-160 */
void oracle_compute_histogram_stats(const double* db, const uint8_t* mask, size_t rows, size_t cols,
                                    oracle_stats* out, uint64_t* hist4096_or_null);
/* autoscale.rs:220-345 */
void oracle_clahe_equalize_normalized(const double* norm, const uint8_t* mask, size_t rows, size_t cols,
                                      size_t tiles_x, size_t tiles_y, double clip_limit, size_t num_bins,
                                      double* out, double* cdfs_or_null);
/* autoscale.rs:348-364 */
void oracle_scale_u16_to_u8(const uint16_t* data, size_t n, uint8_t* out);
/* autoscale.rs:368-448 */
void oracle_autoscale_db_image(const double* db, const uint8_t* mask, size_t rows, size_t cols,
                               int bit_depth, uint16_t* out, oracle_stats* stats_or_null);
/* autoscale.rs:452-659 */
void oracle_autoscale_db_image_advanced(const double* db, const uint8_t* mask, size_t rows, size_t cols,
                                        int bit_depth, int strategy, uint16_t* out,
                                        oracle_stats* stats_or_null);
/* autoscale.rs:710-742 */
void oracle_autoscale_db_image_tamed_synrgb_u8(const double* db, const uint8_t* mask, size_t rows,
                                               size_t cols, int is_copol, uint8_t* out);
/* pipeline.rs:42-66; db/mask outputs optional (NULL); out_u8 filled for U8, out_u16 for U16 */
void oracle_process_scalar_data_pipeline(const float* v, size_t rows, size_t cols, int bit_depth,
                                         int strategy, double* db_or_null, uint8_t* mask_or_null,
                                         uint8_t* out_u8, uint16_t* out_u16, oracle_stats* stats_or_null);
/* ops.rs:4-44 */
void oracle_pol_op(int op, const float* a, const float* b, size_t n, float* out);

/* resize.rs:6-30 */
void oracle_calculate_resize_dimensions(size_t cols, size_t rows, size_t target, size_t* new_cols,
                                        size_t* new_rows);
/* resize.rs:32-53 / 55-89 (fast_image_resize Lanczos3, restated); returns 0 on success */
int oracle_resize_u8_image(const uint8_t* data, size_t cols, size_t rows, size_t tcols, size_t trows,
                           uint8_t* out);
int oracle_resize_u16_image(const uint16_t* data, size_t cols, size_t rows, size_t tcols, size_t trows,
                            uint16_t* out);
/* padding.rs:5-49; out has max(cols,rows)^2 elements; returns 0 / -1 ("U16 data required...") */
int oracle_add_padding_to_square(const uint8_t* u8_data, const uint16_t* u16_data, size_t cols, size_t rows,
                                 int bit_depth, uint8_t* out_u8, uint16_t* out_u16);
/* resize.rs:91-236. Output dims via oracle_resize_output_dims; out buffers sized accordingly. */
void oracle_resize_output_dims(size_t cols, size_t rows, int has_target, size_t target, int pad,
                               size_t* out_cols, size_t* out_rows);
int oracle_resize_image_data_with_meta(const uint8_t* u8_data, const uint16_t* u16_data, size_t cols,
                                       size_t rows, int has_target, size_t target, int bit_depth, int pad,
                                       uint8_t* out_u8, uint16_t* out_u16, oracle_resize_meta* meta);

/* synthetic_rgb.rs:10-67 / 88-178 / 182-197 */
void oracle_create_synthetic_rgb(const uint8_t* b1, const uint8_t* b2, size_t n, uint8_t* rgb);
void oracle_create_synthetic_rgb_suppressed(const uint8_t* b1, const uint8_t* b2, size_t n, uint8_t* rgb);
void oracle_create_synthetic_rgb_by_mode_and_strategy(int mode, int strategy, const uint8_t* b1,
                                                      const uint8_t* b2, size_t n, uint8_t* rgb);

/* Orchestration orders (caller side).
 * single band:   save.rs:49-65 (TIFF) / 119-134 (JPEG, forces U8); api/mod.rs:84-130, 250-281
 * multiband:     save.rs:199-316 (TIFF, two independent bands); api/mod.rs:133-200
 * synRGB JPEG:   save.rs:317-368 (with the Tamed band-specific step) when tamed_band_step!=0,
 *                api/mod.rs:203-247 (without it) when tamed_band_step==0.
 * Output sizes from oracle_resize_output_dims. Return 0 / -1. */
int oracle_pipeline_single(const float* v, size_t rows, size_t cols, int format, int bit_depth, int strategy,
                           int has_target, size_t target, int pad, uint8_t* out_u8, uint16_t* out_u16,
                           oracle_resize_meta* meta);
int oracle_pipeline_multiband_tiff(const float* v1, const float* v2, size_t rows, size_t cols, int bit_depth,
                                   int strategy, int has_target, size_t target, int pad, uint8_t* out1_u8,
                                   uint16_t* out1_u16, uint8_t* out2_u8, uint16_t* out2_u16,
                                   oracle_resize_meta* meta);
int oracle_pipeline_synrgb_jpeg(const float* v1, const float* v2, size_t rows, size_t cols, int strategy,
                                int mode, int has_target, size_t target, int pad, int tamed_band_step,
                                uint8_t* out_rgb, oracle_resize_meta* meta);

/* threads used by the Lanczos stage only (the crate's `rayon` feature, Cargo.toml:33); the rest of
 * the reference path is serial. 0/1 = serial. */
void oracle_set_resize_threads(int n);

/* ---- downsample-on-read (oracle_read.cpp; sentinel1.rs:1074-1109 -> gdal.rs:145-177). PARITY UNPINNED: restates the
 * published resampling of the system libgdal (version unpinned, not on this box). alg: 0 = Average, 1 = Lanczos. */
void oracle_read_dims_for_target(size_t cols, size_t rows, size_t target, size_t* out_cols, size_t* out_rows, int* alg);
void oracle_read_band_resampled_u16(const uint16_t* src, size_t rows, size_t cols, size_t out_cols, size_t out_rows, int alg, float* out);
void oracle_read_band_resampled_f32(const float* src, size_t rows, size_t cols, size_t out_cols, size_t out_rows, int alg, float* out);

#ifdef __cplusplus
}
#endif
#endif

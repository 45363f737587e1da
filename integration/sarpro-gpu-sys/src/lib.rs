//! Raw bindings to `include/sarpro_gpu.h` (ABI version 1): the B200 implementation of SARPRO's per-pixel raster path
//! (`src/core/processing/{pipeline,autoscale,ops,resize,padding,synthetic_rgb}.rs`). Hand-written `extern "C"`, no bindgen.
//! Generated from the header by `integration/gen_sys.py` in the sarpro-b200 tree; NOT compiled there (no Rust toolchain in
//! that image) — the same ABI is exercised through the Python/ctypes mirror by its test-suite.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct sarpro_ctx {
    _private: [u8; 0],
}

/// `HistogramStats` (autoscale.rs:7-24) plus the window the strategy chose (low / high clip, gamma).
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct sarpro_stats {
    pub valid_count: u64,
    pub min_db: f64,
    pub max_db: f64,
    pub mean_db: f64,
    pub std_db: f64,
    pub median_db: f64,
    pub p01: f64,
    pub p02: f64,
    pub p05: f64,
    pub p10: f64,
    pub p25: f64,
    pub p75: f64,
    pub p90: f64,
    pub p95: f64,
    pub p98: f64,
    pub p99: f64,
    pub low_clip: f64,
    pub high_clip: f64,
    pub gamma: f64,
}

/// The tuple tail of `resize_image_data_with_meta` (resize.rs:91-108).
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct sarpro_resize_meta {
    pub cols: u64,
    pub rows: u64,
    pub scale_x: f64,
    pub scale_y: f64,
    pub pad_left: u64,
    pub pad_top: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct sarpro_band {
    pub data: *const c_void,
    pub dtype: i32,
    pub location: i32,
    pub rows: u64,
    pub cols: u64,
}

#[repr(C)]
pub struct sarpro_image {
    pub data: *mut c_void,
    pub location: i32,
    pub bit_depth: i32,
    pub capacity_bytes: u64,
    pub cols: u64,
    pub rows: u64,
    pub channels: i32,
    pub reserved: i32,
    pub meta: sarpro_resize_meta,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct sarpro_timing {
    pub total_ms: f32,
    pub h2d_ms: f32,
    pub d2h_ms: f32,
    pub kernel_ms: f32,
    pub kernel_launches: u32,
    pub host_syncs: u32,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub stage_ms: [f32; 8],
    pub stage_launches: [u32; 8],
}

/// One scene of a batch (sarpro_scene): the band pair of one product.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct sarpro_scene {
    pub b1: sarpro_band,
    pub b2: sarpro_band,
}

/// BatchReport (api/mod.rs:451-457)
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct sarpro_batch_report {
    pub processed: u64,
    pub skipped: u64,
    pub errors: u64,
}

pub const SARPRO_BATCH_MULTIBAND: i32 = 0;
pub const SARPRO_BATCH_SYNRGB: i32 = 1;

pub const SARPRO_DT_F32: i32 = 0;
pub const SARPRO_DT_U16: i32 = 1;
pub const SARPRO_LOC_HOST: i32 = 0;
pub const SARPRO_LOC_DEVICE: i32 = 1;
pub const SARPRO_LOC_NONE: i32 = 2;
pub const SARPRO_U8: c_int = 0;
pub const SARPRO_U16: c_int = 1;
pub const SARPRO_FORMAT_TIFF: c_int = 0;
pub const SARPRO_FORMAT_JPEG: c_int = 1;
pub const SARPRO_RESAMPLE_AVERAGE: c_int = 0;
pub const SARPRO_RESAMPLE_LANCZOS: c_int = 1;
pub const SARPRO_OP_NONE: c_int = -1;
pub const SARPRO_OK: c_int = 0;
pub const SARPRO_ERR_NO_DEVICE: c_int = -2;

extern "C" {
    pub fn sarpro_abi_version() -> c_int;
    pub fn sarpro_ctx_create(out: *mut *mut sarpro_ctx, device_id: c_int) -> c_int;
    pub fn sarpro_ctx_destroy(ctx: *mut sarpro_ctx);
    pub fn sarpro_last_error(ctx: *const sarpro_ctx) -> *const c_char;
    pub fn sarpro_ctx_set_stream(ctx: *mut sarpro_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn sarpro_ctx_synchronize(ctx: *mut sarpro_ctx) -> c_int;
    pub fn sarpro_last_timing(ctx: *const sarpro_ctx, out: *mut sarpro_timing) -> c_int;
    pub fn sarpro_host_alloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn sarpro_host_free(p: *mut c_void);
    pub fn sarpro_host_register(p: *mut c_void, bytes: usize) -> c_int;
    pub fn sarpro_host_unregister(p: *mut c_void);
    pub fn sarpro_pol_op(ctx: *mut sarpro_ctx, op: c_int, a: *const f32, b: *const f32, rows: usize, cols: usize, out: *mut f32) -> c_int;
    pub fn sarpro_process_scalar_data_pipeline(ctx: *mut sarpro_ctx, v: *const f32, rows: usize, cols: usize, bit_depth: c_int, strategy: c_int, out_u8: *mut u8, out_u16: *mut u16, stats: *mut sarpro_stats) -> c_int;
    pub fn sarpro_process_dn_pipeline(ctx: *mut sarpro_ctx, dn: *const u16, rows: usize, cols: usize, bit_depth: c_int, strategy: c_int, out_u8: *mut u8, out_u16: *mut u16, stats: *mut sarpro_stats) -> c_int;
    pub fn sarpro_process_scalar_data_inplace(ctx: *mut sarpro_ctx, v: *const f32, rows: usize, cols: usize, db: *mut f64, valid_mask: *mut u8) -> c_int;
    pub fn sarpro_autoscale_tamed_synrgb_u8(ctx: *mut sarpro_ctx, v: *const f32, rows: usize, cols: usize, is_copol: c_int, out: *mut u8) -> c_int;
    pub fn sarpro_scale_u16_to_u8(ctx: *mut sarpro_ctx, data: *const u16, n: usize, out: *mut u8) -> c_int;
    pub fn sarpro_resize_output_dims(cols: usize, rows: usize, has_target: c_int, target: usize, pad: c_int, out_cols: *mut usize, out_rows: *mut usize) -> c_int;
    pub fn sarpro_resize_image_data_with_meta(ctx: *mut sarpro_ctx, u8_data: *const u8, u16_data: *const u16, cols: usize, rows: usize, has_target: c_int, target: usize, bit_depth: c_int, pad: c_int, out_u8: *mut u8, out_u16: *mut u16, meta: *mut sarpro_resize_meta) -> c_int;
    pub fn sarpro_add_padding_to_square(ctx: *mut sarpro_ctx, u8_data: *const u8, u16_data: *const u16, cols: usize, rows: usize, bit_depth: c_int, out_u8: *mut u8, out_u16: *mut u16) -> c_int;
    pub fn sarpro_create_synthetic_rgb_by_mode_and_strategy(ctx: *mut sarpro_ctx, mode: c_int, strategy: c_int, band1: *const u8, band2: *const u8, n: usize, rgb: *mut u8) -> c_int;
    pub fn sarpro_pipeline_single(ctx: *mut sarpro_ctx, a: *const sarpro_band, b: *const sarpro_band, op: c_int, format: c_int, bit_depth: c_int, strategy: c_int, has_target: c_int, target: usize, pad: c_int, out: *mut sarpro_image, stats: *mut sarpro_stats) -> c_int;
    pub fn sarpro_pipeline_multiband_tiff(ctx: *mut sarpro_ctx, b1: *const sarpro_band, b2: *const sarpro_band, bit_depth: c_int, strategy: c_int, has_target: c_int, target: usize, pad: c_int, out1: *mut sarpro_image, out2: *mut sarpro_image, stats2: *mut sarpro_stats) -> c_int;
    pub fn sarpro_pipeline_synrgb(ctx: *mut sarpro_ctx, b1: *const sarpro_band, b2: *const sarpro_band, strategy: c_int, mode: c_int, has_target: c_int, target: usize, pad: c_int, tamed_band_step: c_int, out: *mut sarpro_image, stats2: *mut sarpro_stats) -> c_int;
    pub fn sarpro_comm_unique_id(out128: *mut c_void) -> c_int;
    pub fn sarpro_comm_init(ctx: *mut sarpro_ctx, unique_id128: *const c_void, rank: c_int, world: c_int) -> c_int;
    pub fn sarpro_comm_destroy(ctx: *mut sarpro_ctx) -> c_int;
    pub fn sarpro_shard_rows(rows: usize, world: c_int, rank: c_int, clahe: c_int, r0: *mut usize, r1: *mut usize) -> c_int;
    pub fn sarpro_shard_halo_rows(rows: usize, cols: usize, has_target: c_int, target: usize, world: c_int, rank: c_int, clahe: c_int, h0: *mut usize, h1: *mut usize) -> c_int;
    pub fn sarpro_pipeline_synrgb_sharded(ctx: *mut sarpro_ctx, b1: *const sarpro_band, b2: *const sarpro_band, scene_rows: usize, strategy: c_int, mode: c_int, has_target: c_int, target: usize, pad: c_int, tamed_band_step: c_int, out: *mut sarpro_image) -> c_int;
    pub fn sarpro_pipeline_polops(ctx: *mut sarpro_ctx, a: *const sarpro_band, b: *const sarpro_band, scene_rows: usize, n_ops: c_int, ops: *const c_int, bit_depth: c_int, strategy: c_int, outs: *mut sarpro_image, stats: *mut sarpro_stats) -> c_int;
    pub fn sarpro_pipeline_single_sharded(ctx: *mut sarpro_ctx, a: *const sarpro_band, b: *const sarpro_band, scene_rows: usize, op: c_int, bit_depth: c_int, strategy: c_int, out: *mut sarpro_image, stats: *mut sarpro_stats) -> c_int;
    pub fn sarpro_read_dims_for_target(cols: usize, rows: usize, target: usize, out_cols: *mut usize, out_rows: *mut usize, alg: *mut c_int) -> c_int;
    pub fn sarpro_read_band_resampled(ctx: *mut sarpro_ctx, in_: *const sarpro_band, out_cols: usize, out_rows: usize, alg: c_int, out: *mut f32, out_location: c_int) -> c_int;
    pub fn sarpro_encode_jpeg(ctx: *mut sarpro_ctx, img: *const sarpro_image, quality: c_int, out: *mut c_void, capacity: usize, out_bytes: *mut usize) -> c_int;
    pub fn sarpro_encode_last_jpeg(ctx: *mut sarpro_ctx, which: c_int, quality: c_int, out: *mut c_void, capacity: usize, out_bytes: *mut usize) -> c_int;
    pub fn sarpro_pipeline_batch(ctx: *mut sarpro_ctx, scenes: *const sarpro_scene, n: usize, kind: c_int, bit_depth: c_int, strategy: c_int, mode: c_int, has_target: c_int, target: usize, pad: c_int, tamed_band_step: c_int, continue_on_error: c_int, outs: *mut sarpro_image, stats: *mut sarpro_stats, statuses: *mut c_int, report: *mut sarpro_batch_report) -> c_int;
    pub fn sarpro_plan_from_dn_histogram(hist65536: *const u64, bit_depth: c_int, strategy: c_int, stats: *mut sarpro_stats, lut16: *mut u16) -> c_int;
    pub fn sarpro_plan_from_present_list(blocks: *const u32, pairs: *const u32, cap: u32, bit_depth: c_int, strategy: c_int, stats: *mut sarpro_stats, lut16: *mut u16) -> c_int;
    pub fn sarpro_plan_on_device(ctx: *mut sarpro_ctx, hist65536: *const u32, bit_depth: c_int, strategy: c_int, plan_kind: c_int, stats: *mut sarpro_stats, lut16: *mut u16, hot2: *mut u32) -> c_int;
    pub fn sarpro_plan_kind_from_dn_histogram(hist65536: *const u64, bit_depth: c_int, strategy: c_int, plan_kind: c_int, stats: *mut sarpro_stats, lut16: *mut u16, hot2: *mut u32) -> c_int;
    pub fn sarpro_lanczos_row_plan_check(samples: *const u8, in_size: usize, out_size: usize, max_span: usize, strip_ntiles: usize, out_direct: *mut u8, out_replay: *mut u8) -> c_int;
    pub fn sarpro_lanczos_row_check_u16(samples: *const u16, in_size: usize, out_size: usize, out: *mut u16) -> c_int;
    pub fn sarpro_read_row_plan_check(samples: *const u16, in_size: usize, out_size: usize, alg: c_int, out: *mut f32) -> c_int;
    pub fn sarpro_f32_guard_params(low_db: f64, range_db: f64, n: u32, min_v: f32, max_v: f32, e0: *mut c_int, f0: *mut f32, scale: *mut f32, guard: *mut f32) -> c_int;
    pub fn sarpro_f32_edges_check(kind: c_int, low_db: f64, high_db: f64, gamma: f64, n_levels: u32, min_v: f32, max_v: f32, n_analytic: *mut u32, n_mismatch: *mut u32, edges_out: *mut f32) -> c_int;
    pub fn sarpro_synrgb_lut_check(set: c_int, hist256: *const u32, n_per_band: u64, floor_with_cushion: *mut c_int, lut_r: *mut u8, lut_g: *mut u8, lut_b: *mut u8) -> c_int;
    pub fn sarpro_plan_from_stat_histogram(hist4096: *const u64, valid_count: u64, min_v: f32, max_v: f32, mean_db: f64, std_db: f64, strategy: c_int, tamed_synrgb_kind: c_int, stats: *mut sarpro_stats) -> c_int;
    pub fn sarpro_narrow_f32_check(src: *const f32, n: usize, dst: *mut u16, u16_valued: *mut c_int) -> c_int;
}

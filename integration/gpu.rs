//! `src/core/processing/gpu.rs` of SARPRO, `#[cfg(feature = "gpu")]`: safe wrapper over `sarpro-gpu-sys` for the call sites
//! patched by `integration/patches/`. Source only — not compiled in the sarpro-b200 tree (no Rust
//! toolchain there); `tests/test_host_cpu.py::test_rust_wrapper_covers_the_c_abi` keeps the entry points it names in sync
//! with `include/sarpro_gpu.h`.
//!
//! Two granularities, like the header:
//!  * stage level — same signatures and return tuples as `pipeline.rs`, `resize.rs`, `padding.rs`, `synthetic_rgb.rs`,
//!    `ops.rs`, `autoscale.rs:710`, so a function body can be replaced one for one;
//!  * fused pipelines — the call orders of `save.rs:49-65,119-134,199-316,317-368` and `api/mod.rs:84-369` in one call each,
//!    data staying in HBM between the stages.
use ndarray::Array2;
use sarpro_gpu_sys as sys;
use std::ffi::CStr;

use crate::types::{AutoscaleStrategy, BitDepth, OutputFormat, PolarizationOperation, SyntheticRgbMode};

type DynResult<T> = Result<T, Box<dyn std::error::Error>>;

/// One per thread: a `sarpro_ctx` is single-threaded, like the reference path (CLI main thread, GUI worker).
pub struct Gpu(*mut sys::sarpro_ctx);

/// `(cols, rows, u8, u16, scale_x, scale_y, pad_left, pad_top)` — the tuple of `resize_image_data_with_meta` (resize.rs:99-110).
pub type ResizedWithMeta = (usize, usize, Vec<u8>, Option<Vec<u16>>, f64, f64, usize, usize);

/// One processed gray band or band pair as the save / api arms consume it.
pub struct Processed {
    pub cols: usize,
    pub rows: usize,
    pub u8_data: Vec<u8>,
    pub u16_data: Option<Vec<u16>>,
    pub band2_u8: Vec<u8>,
    pub band2_u16: Option<Vec<u16>>,
    pub meta: sys::sarpro_resize_meta,
    pub stats: [sys::sarpro_stats; 2],
}

fn last_error(ctx: *const sys::sarpro_ctx) -> String {
    unsafe { CStr::from_ptr(sys::sarpro_last_error(ctx)).to_string_lossy().into_owned() }
}
fn depth(b: BitDepth) -> i32 {
    match b { BitDepth::U8 => sys::SARPRO_U8, BitDepth::U16 => sys::SARPRO_U16 }
}
fn format(f: OutputFormat) -> i32 {
    match f { OutputFormat::TIFF => sys::SARPRO_FORMAT_TIFF, OutputFormat::JPEG => sys::SARPRO_FORMAT_JPEG }
}
fn op_code(op: PolarizationOperation) -> i32 {
    op as i32 // declaration order of types.rs:8-14 == sarpro_pol_op
}

impl Gpu {
    pub fn new(device: i32) -> DynResult<Self> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { sys::sarpro_ctx_create(&mut p, device) };
        if rc != sys::SARPRO_OK {
            return Err(last_error(std::ptr::null()).into()); // no CUDA device: the GPU build has no CPU fallback
        }
        Ok(Gpu(p))
    }

    fn check(&self, rc: i32) -> DynResult<()> {
        if rc == sys::SARPRO_OK { Ok(()) } else { Err(last_error(self.0).into()) } // same Box<dyn Error> as resize.rs / padding.rs
    }

    fn band(a: &Array2<f32>) -> sys::sarpro_band {
        let (rows, cols) = a.dim();
        assert!(a.is_standard_layout()); // from_shape_vec arrays are row-major contiguous (gdal.rs:131, pipeline.rs:25)
        sys::sarpro_band { data: a.as_ptr() as *const _, dtype: sys::SARPRO_DT_F32, location: sys::SARPRO_LOC_HOST, rows: rows as u64, cols: cols as u64 }
    }
    /// Raw GRD DN as stored in the measurement TIFF (the reader can hand these over instead of widening to f32: half the upload).
    pub fn band_dn(dn: &[u16], rows: usize, cols: usize) -> sys::sarpro_band {
        sys::sarpro_band { data: dn.as_ptr() as *const _, dtype: sys::SARPRO_DT_U16, location: sys::SARPRO_LOC_HOST, rows: rows as u64, cols: cols as u64 }
    }
    fn image(buf: *mut u8, bytes: usize, d: i32) -> sys::sarpro_image {
        sys::sarpro_image {
            data: buf as *mut _, location: sys::SARPRO_LOC_HOST, bit_depth: d, capacity_bytes: bytes as u64,
            cols: 0, rows: 0, channels: 0, reserved: 0, meta: sys::sarpro_resize_meta::default(),
        }
    }
    fn out_dims(&self, cols: usize, rows: usize, target_size: Option<usize>, pad: bool) -> DynResult<(usize, usize)> {
        let (mut oc, mut or_) = (0usize, 0usize);
        self.check(unsafe { sys::sarpro_resize_output_dims(cols, rows, target_size.is_some() as i32, target_size.unwrap_or(0), pad as i32, &mut oc, &mut or_) })?;
        Ok((oc, or_))
    }

    // ---------------------------------------------------------------------------------------------------------------
    // stage level: one entry per reference function
    // ---------------------------------------------------------------------------------------------------------------

    /// `process_scalar_data_inplace` (pipeline.rs:8-40).
    pub fn process_scalar_data_inplace(&self, processed: &Array2<f32>) -> DynResult<(Array2<f64>, Vec<bool>)> {
        let (rows, cols) = processed.dim();
        let mut db = vec![0f64; rows * cols];
        let mut mask = vec![0u8; rows * cols];
        let p = processed.as_standard_layout();
        self.check(unsafe { sys::sarpro_process_scalar_data_inplace(self.0, p.as_ptr(), rows, cols, db.as_mut_ptr(), mask.as_mut_ptr()) })?;
        Ok((Array2::from_shape_vec((rows, cols), db)?, mask.into_iter().map(|m| m != 0).collect()))
    }

    /// `process_scalar_data_pipeline` (pipeline.rs:42-66) without the dB plane and mask (callers use them for `.dim()` only,
    /// save.rs:53,122,201,329 — except the Tamed band step, which has its own entry below).
    pub fn process_scalar_data_pipeline(&self, processed: &Array2<f32>, bit_depth: BitDepth, strategy: AutoscaleStrategy)
                                        -> DynResult<(Vec<u8>, Option<Vec<u16>>, sys::sarpro_stats)> {
        let (rows, cols) = processed.dim();
        let n = rows * cols;
        let mut u8v = if matches!(bit_depth, BitDepth::U8) { vec![0u8; n] } else { Vec::new() };
        let mut u16v = if matches!(bit_depth, BitDepth::U16) { vec![0u16; n] } else { Vec::new() };
        let mut st = sys::sarpro_stats::default();
        let p = processed.as_standard_layout();
        self.check(unsafe {
            sys::sarpro_process_scalar_data_pipeline(self.0, p.as_ptr(), rows, cols, depth(bit_depth), strategy as i32,
                                                     if u8v.is_empty() { std::ptr::null_mut() } else { u8v.as_mut_ptr() },
                                                     if u16v.is_empty() { std::ptr::null_mut() } else { u16v.as_mut_ptr() }, &mut st)
        })?;
        Ok((u8v, if matches!(bit_depth, BitDepth::U16) { Some(u16v) } else { None }, st))
    }

    /// `autoscale_db_image_tamed_synrgb_u8` (autoscale.rs:710-742), from the linear band the dB plane was derived from.
    pub fn autoscale_tamed_synrgb_u8(&self, processed: &Array2<f32>, is_copol: bool) -> DynResult<Vec<u8>> {
        let (rows, cols) = processed.dim();
        let mut out = vec![0u8; rows * cols];
        let p = processed.as_standard_layout();
        self.check(unsafe { sys::sarpro_autoscale_tamed_synrgb_u8(self.0, p.as_ptr(), rows, cols, is_copol as i32, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// `scale_u16_to_u8` (autoscale.rs:348-364).
    pub fn scale_u16_to_u8(&self, data: &[u16]) -> DynResult<Vec<u8>> {
        let mut out = vec![0u8; data.len()];
        self.check(unsafe { sys::sarpro_scale_u16_to_u8(self.0, data.as_ptr(), data.len(), out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// `sum_arrays` .. `log_ratio_arrays` (ops.rs:4-44).
    pub fn pol_op(&self, op: PolarizationOperation, a: &Array2<f32>, b: &Array2<f32>) -> DynResult<Array2<f32>> {
        let (rows, cols) = a.dim();
        let mut out = vec![0f32; rows * cols];
        let (pa, pb) = (a.as_standard_layout(), b.as_standard_layout());
        self.check(unsafe { sys::sarpro_pol_op(self.0, op_code(op), pa.as_ptr(), pb.as_ptr(), rows, cols, out.as_mut_ptr()) })?;
        Ok(Array2::from_shape_vec((rows, cols), out)?)
    }

    /// `resize_image_data_with_meta` (resize.rs:91-235): same tuple, same `"U16 data required for U16 bit depth"` error.
    pub fn resize_image_data_with_meta(&self, u8_data: &[u8], u16_data: Option<&[u16]>, original_cols: usize, original_rows: usize,
                                       target_size: Option<usize>, bit_depth: BitDepth, pad: bool) -> DynResult<ResizedWithMeta> {
        let (oc, or_) = self.out_dims(original_cols, original_rows, target_size, pad)?;
        let mut o8 = if matches!(bit_depth, BitDepth::U8) { vec![0u8; oc * or_] } else { Vec::new() };
        let mut o16 = if matches!(bit_depth, BitDepth::U16) { vec![0u16; oc * or_] } else { Vec::new() };
        let mut meta = sys::sarpro_resize_meta::default();
        self.check(unsafe {
            sys::sarpro_resize_image_data_with_meta(self.0, if u8_data.is_empty() { std::ptr::null() } else { u8_data.as_ptr() },
                                                    u16_data.map_or(std::ptr::null(), |d| d.as_ptr()), original_cols, original_rows,
                                                    target_size.is_some() as i32, target_size.unwrap_or(0), depth(bit_depth), pad as i32,
                                                    if o8.is_empty() { std::ptr::null_mut() } else { o8.as_mut_ptr() },
                                                    if o16.is_empty() { std::ptr::null_mut() } else { o16.as_mut_ptr() }, &mut meta)
        })?;
        Ok((meta.cols as usize, meta.rows as usize, o8, if matches!(bit_depth, BitDepth::U16) { Some(o16) } else { None },
            meta.scale_x, meta.scale_y, meta.pad_left as usize, meta.pad_top as usize))
    }

    /// `resize_image_data` (resize.rs:238-257).
    pub fn resize_image_data(&self, u8_data: &[u8], u16_data: Option<&[u16]>, original_cols: usize, original_rows: usize,
                             target_size: Option<usize>, bit_depth: BitDepth, pad: bool) -> DynResult<(usize, usize, Vec<u8>, Option<Vec<u16>>)> {
        let (c, r, a, b, _, _, _, _) = self.resize_image_data_with_meta(u8_data, u16_data, original_cols, original_rows, target_size, bit_depth, pad)?;
        Ok((c, r, a, b))
    }

    /// `add_padding_to_square` (padding.rs:5-49).
    pub fn add_padding_to_square(&self, u8_data: &[u8], u16_data: Option<&[u16]>, cols: usize, rows: usize, bit_depth: BitDepth)
                                 -> DynResult<(Vec<u8>, Option<Vec<u16>>)> {
        let m = cols.max(rows);
        let mut o8 = if matches!(bit_depth, BitDepth::U8) { vec![0u8; m * m] } else { Vec::new() };
        let mut o16 = if matches!(bit_depth, BitDepth::U16) { vec![0u16; m * m] } else { Vec::new() };
        self.check(unsafe {
            sys::sarpro_add_padding_to_square(self.0, if u8_data.is_empty() { std::ptr::null() } else { u8_data.as_ptr() },
                                              u16_data.map_or(std::ptr::null(), |d| d.as_ptr()), cols, rows, depth(bit_depth),
                                              if o8.is_empty() { std::ptr::null_mut() } else { o8.as_mut_ptr() },
                                              if o16.is_empty() { std::ptr::null_mut() } else { o16.as_mut_ptr() })
        })?;
        Ok((o8, if matches!(bit_depth, BitDepth::U16) { Some(o16) } else { None }))
    }

    /// `create_synthetic_rgb_by_mode_and_strategy` (synthetic_rgb.rs:182-197).
    pub fn create_synthetic_rgb_by_mode_and_strategy(&self, mode: SyntheticRgbMode, strategy: AutoscaleStrategy, band1: &[u8], band2: &[u8])
                                                     -> DynResult<Vec<u8>> {
        let n = band1.len().min(band2.len());
        let mut rgb = vec![0u8; n * 3];
        self.check(unsafe {
            sys::sarpro_create_synthetic_rgb_by_mode_and_strategy(self.0, mode as i32, strategy as i32, band1.as_ptr(), band2.as_ptr(), n, rgb.as_mut_ptr())
        })?;
        Ok(rgb)
    }

    // ---------------------------------------------------------------------------------------------------------------
    // fused pipelines
    // ---------------------------------------------------------------------------------------------------------------

    /// `save_processed_image` arms (save.rs:49-65 TIFF, :119-134 JPEG) and `api/mod.rs:84-130, 249-292`; with `op` the
    /// polarization-operation arms `api/mod.rs:284-369` (`band2` required).
    pub fn single(&self, processed: &Array2<f32>, band2: Option<&Array2<f32>>, op: Option<PolarizationOperation>, fmt: OutputFormat,
                  bit_depth: BitDepth, strategy: AutoscaleStrategy, target_size: Option<usize>, pad: bool) -> DynResult<Processed> {
        let (rows, cols) = processed.dim();
        let (oc, or_) = self.out_dims(cols, rows, target_size, pad)?;
        let d = if matches!(fmt, OutputFormat::JPEG) { sys::SARPRO_U8 } else { depth(bit_depth) }; // save.rs:121
        let mut o8 = if d == sys::SARPRO_U8 { vec![0u8; oc * or_] } else { Vec::new() };
        let mut o16 = if d == sys::SARPRO_U16 { vec![0u16; oc * or_] } else { Vec::new() };
        let mut img = if d == sys::SARPRO_U8 { Self::image(o8.as_mut_ptr(), o8.len(), d) } else { Self::image(o16.as_mut_ptr() as *mut u8, o16.len() * 2, d) };
        let b1 = Self::band(processed);
        let b2 = band2.map(Self::band);
        let mut stats = [sys::sarpro_stats::default(); 2];
        self.check(unsafe {
            sys::sarpro_pipeline_single(self.0, &b1, b2.as_ref().map_or(std::ptr::null(), |b| b as *const _), op.map_or(-1, op_code), format(fmt), d,
                                        strategy as i32, target_size.is_some() as i32, target_size.unwrap_or(0), pad as i32, &mut img, stats.as_mut_ptr())
        })?;
        Ok(Processed { cols: img.cols as usize, rows: img.rows as usize, u8_data: o8, u16_data: if d == sys::SARPRO_U16 { Some(o16) } else { None },
                       band2_u8: Vec::new(), band2_u16: None, meta: img.meta, stats })
    }

    /// The two-band TIFF arm of `save_processed_multiband_image_sequential` (save.rs:199-316) / `api/mod.rs:133-200`.
    pub fn multiband_tiff(&self, band1: &Array2<f32>, band2: &Array2<f32>, bit_depth: BitDepth, strategy: AutoscaleStrategy,
                          target_size: Option<usize>, pad: bool) -> DynResult<Processed> {
        let (rows, cols) = band1.dim();
        let (oc, or_) = self.out_dims(cols, rows, target_size, pad)?;
        let d = depth(bit_depth);
        let n = oc * or_;
        let (mut a8, mut b8) = if d == sys::SARPRO_U8 { (vec![0u8; n], vec![0u8; n]) } else { (Vec::new(), Vec::new()) };
        let (mut a16, mut b16) = if d == sys::SARPRO_U16 { (vec![0u16; n], vec![0u16; n]) } else { (Vec::new(), Vec::new()) };
        let (mut i1, mut i2) = if d == sys::SARPRO_U8 {
            (Self::image(a8.as_mut_ptr(), n, d), Self::image(b8.as_mut_ptr(), n, d))
        } else {
            (Self::image(a16.as_mut_ptr() as *mut u8, n * 2, d), Self::image(b16.as_mut_ptr() as *mut u8, n * 2, d))
        };
        let (s1, s2) = (Self::band(band1), Self::band(band2));
        let mut stats = [sys::sarpro_stats::default(); 2];
        self.check(unsafe {
            sys::sarpro_pipeline_multiband_tiff(self.0, &s1, &s2, d, strategy as i32, target_size.is_some() as i32, target_size.unwrap_or(0), pad as i32,
                                                &mut i1, &mut i2, stats.as_mut_ptr())
        })?;
        Ok(Processed { cols: i1.cols as usize, rows: i1.rows as usize, u8_data: a8, u16_data: if d == sys::SARPRO_U16 { Some(a16) } else { None },
                       band2_u8: b8, band2_u16: if d == sys::SARPRO_U16 { Some(b16) } else { None }, meta: i1.meta, stats })
    }

    /// The synthetic-RGB arm of `save_processed_multiband_image_sequential` (save.rs:317-368, `tamed_band_step = true`) /
    /// `api/mod.rs:203-247, 394-438` (`false`): per band dB -> autoscale (-> Tamed band step) -> resize + pad, then
    /// `create_synthetic_rgb_by_mode_and_strategy`. Returns (cols, rows, rgb, meta).
    pub fn synthetic_rgb(&self, band1: &Array2<f32>, band2: &Array2<f32>, strategy: AutoscaleStrategy, mode: SyntheticRgbMode,
                         target_size: Option<usize>, pad: bool, tamed_band_step: bool)
                         -> DynResult<(usize, usize, Vec<u8>, sys::sarpro_resize_meta)> {
        let (rows, cols) = band1.dim();
        let (oc, or_) = self.out_dims(cols, rows, target_size, pad)?;
        let mut rgb = vec![0u8; oc * or_ * 3];
        let (b1, b2) = (Self::band(band1), Self::band(band2));
        let mut img = Self::image(rgb.as_mut_ptr(), rgb.len(), sys::SARPRO_U8);
        self.check(unsafe {
            sys::sarpro_pipeline_synrgb(self.0, &b1, &b2, strategy as i32, mode as i32, target_size.is_some() as i32, target_size.unwrap_or(0),
                                        pad as i32, tamed_band_step as i32, &mut img, std::ptr::null_mut())
        })?;
        Ok((img.cols as usize, img.rows as usize, rgb, img.meta))
    }

    /// The same arm when the writer is the JPEG one (save.rs:370): the RGB image stays in HBM (`SARPRO_LOC_NONE`), nvJPEG
    /// encodes it there and only the stream comes back (`write_rgb_jpeg`, io/writers/jpeg.rs:19-30, then just writes bytes).
    pub fn synthetic_rgb_jpeg_stream(&self, band1: &Array2<f32>, band2: &Array2<f32>, strategy: AutoscaleStrategy, mode: SyntheticRgbMode,
                                     target_size: Option<usize>, pad: bool) -> DynResult<(usize, usize, Vec<u8>, sys::sarpro_resize_meta)> {
        let (b1, b2) = (Self::band(band1), Self::band(band2));
        let mut img = Self::image(std::ptr::null_mut(), 0, sys::SARPRO_U8);
        img.location = sys::SARPRO_LOC_NONE;
        self.check(unsafe {
            sys::sarpro_pipeline_synrgb(self.0, &b1, &b2, strategy as i32, mode as i32, target_size.is_some() as i32, target_size.unwrap_or(0),
                                        pad as i32, 1, &mut img, std::ptr::null_mut())
        })?;
        let mut len = 0usize;
        self.check(unsafe { sys::sarpro_encode_last_jpeg(self.0, 0, 100, std::ptr::null_mut(), 0, &mut len) })?;
        let mut stream = vec![0u8; len];
        self.check(unsafe { sys::sarpro_encode_last_jpeg(self.0, 0, 100, stream.as_mut_ptr() as *mut _, stream.len(), &mut len) })?;
        stream.truncate(len);
        Ok((img.cols as usize, img.rows as usize, stream, img.meta))
    }

    /// Two polarization operations over one pair, each autoscaled into a full-resolution band (the two calls of
    /// sentinel1.rs:1497-1579 -> pipeline.rs:42-66 with the operands read once per pass).
    pub fn polops(&self, a: &Array2<f32>, b: &Array2<f32>, ops: &[PolarizationOperation], bit_depth: BitDepth, strategy: AutoscaleStrategy)
                  -> DynResult<Vec<(Vec<u8>, Option<Vec<u16>>)>> {
        let (rows, cols) = a.dim();
        let n = rows * cols;
        let d = depth(bit_depth);
        let codes: Vec<i32> = ops.iter().map(|o| op_code(*o)).collect();
        let mut bufs8: Vec<Vec<u8>> = ops.iter().map(|_| if d == sys::SARPRO_U8 { vec![0u8; n] } else { Vec::new() }).collect();
        let mut bufs16: Vec<Vec<u16>> = ops.iter().map(|_| if d == sys::SARPRO_U16 { vec![0u16; n] } else { Vec::new() }).collect();
        let mut imgs: Vec<sys::sarpro_image> = (0..ops.len())
            .map(|k| if d == sys::SARPRO_U8 { Self::image(bufs8[k].as_mut_ptr(), n, d) } else { Self::image(bufs16[k].as_mut_ptr() as *mut u8, n * 2, d) })
            .collect();
        let (ba, bb) = (Self::band(a), Self::band(b));
        self.check(unsafe {
            sys::sarpro_pipeline_polops(self.0, &ba, &bb, 0, ops.len() as i32, codes.as_ptr(), d, strategy as i32, imgs.as_mut_ptr(), std::ptr::null_mut())
        })?;
        Ok(bufs8.into_iter().zip(bufs16.into_iter()).map(|(x, y)| (x, if d == sys::SARPRO_U16 { Some(y) } else { None })).collect())
    }

    /// `GdalSarReader::read_band_resampled` (gdal.rs:145-177) for a raster already in memory as raw DN: the downsample-on-read
    /// of sentinel1.rs:1074-1109 without GDAL's resampler (the reader then only decodes the TIFF strips).
    pub fn read_band_resampled(&self, dn: &[u16], rows: usize, cols: usize, target: usize) -> DynResult<Array2<f32>> {
        let (mut oc, mut or_, mut alg) = (0usize, 0usize, 0i32);
        self.check(unsafe { sys::sarpro_read_dims_for_target(cols, rows, target, &mut oc, &mut or_, &mut alg) })?;
        let mut out = vec![0f32; oc * or_];
        let b = Self::band_dn(dn, rows, cols);
        self.check(unsafe { sys::sarpro_read_band_resampled(self.0, &b, oc, or_, alg, out.as_mut_ptr(), sys::SARPRO_LOC_HOST) })?;
        Ok(Array2::from_shape_vec((or_, oc), out)?)
    }

    /// The scene loop of `process_directory_to_path` (api/mod.rs:474-536) over decoded pairs: synthetic-RGB images, next
    /// scene's upload beside the current scene's kernels. Returns the images and the `BatchReport` counters.
    pub fn batch_synthetic_rgb(&self, scenes: &[(Array2<f32>, Array2<f32>)], strategy: AutoscaleStrategy, mode: SyntheticRgbMode,
                               target_size: Option<usize>, pad: bool, continue_on_error: bool)
                               -> DynResult<(Vec<Vec<u8>>, sys::sarpro_batch_report)> {
        let descs: Vec<sys::sarpro_scene> = scenes.iter().map(|(a, b)| sys::sarpro_scene { b1: Self::band(a), b2: Self::band(b) }).collect();
        let mut bufs: Vec<Vec<u8>> = Vec::with_capacity(scenes.len());
        let mut imgs: Vec<sys::sarpro_image> = Vec::with_capacity(scenes.len());
        for (a, _) in scenes {
            let (rows, cols) = a.dim();
            let (oc, or_) = self.out_dims(cols, rows, target_size, pad)?;
            bufs.push(vec![0u8; oc * or_ * 3]);
            let last = bufs.last_mut().unwrap();
            imgs.push(Self::image(last.as_mut_ptr(), last.len(), sys::SARPRO_U8));
        }
        let mut report = sys::sarpro_batch_report::default();
        let mut statuses = vec![0i32; scenes.len()];
        self.check(unsafe {
            sys::sarpro_pipeline_batch(self.0, descs.as_ptr(), descs.len(), sys::SARPRO_BATCH_SYNRGB, sys::SARPRO_U8, strategy as i32, mode as i32,
                                       target_size.is_some() as i32, target_size.unwrap_or(0), pad as i32, 1, continue_on_error as i32,
                                       imgs.as_mut_ptr(), std::ptr::null_mut(), statuses.as_mut_ptr(), &mut report)
        })?;
        Ok((bufs, report))
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { sys::sarpro_ctx_destroy(self.0) }
    }
}

thread_local! {
    pub static GPU: Gpu = Gpu::new(0).expect("sarpro-gpu: no B200 (the GPU build has no CPU fallback)");
}

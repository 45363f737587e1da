//! `src/core/processing/gpu.rs` of SARPRO, `#[cfg(feature = "gpu")]`: safe wrapper over `sarpro-gpu-sys` for the call sites
//! listed in INTEGRATION.md §3. Source only — not compiled in the sarpro-b200 tree (no Rust toolchain there).
use ndarray::Array2;
use sarpro_gpu_sys as sys;
use std::ffi::CStr;

use crate::types::{AutoscaleStrategy, SyntheticRgbMode};

/// One per thread: a `sarpro_ctx` is single-threaded, like the reference path (CLI main thread, GUI worker).
pub struct Gpu(*mut sys::sarpro_ctx);

fn last_error(ctx: *const sys::sarpro_ctx) -> String {
    unsafe { CStr::from_ptr(sys::sarpro_last_error(ctx)).to_string_lossy().into_owned() }
}

impl Gpu {
    pub fn new(device: i32) -> Result<Self, Box<dyn std::error::Error>> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { sys::sarpro_ctx_create(&mut p, device) };
        if rc != sys::SARPRO_OK {
            return Err(last_error(std::ptr::null()).into()); // no CUDA device: the GPU build has no CPU fallback
        }
        Ok(Gpu(p))
    }

    fn check(&self, rc: i32) -> Result<(), Box<dyn std::error::Error>> {
        if rc == sys::SARPRO_OK { Ok(()) } else { Err(last_error(self.0).into()) } // same Box<dyn Error> as resize.rs / padding.rs
    }

    fn band(a: &Array2<f32>) -> sys::sarpro_band {
        let (rows, cols) = a.dim();
        sys::sarpro_band { data: a.as_ptr() as *const _, dtype: sys::SARPRO_DT_F32, location: sys::SARPRO_LOC_HOST, rows: rows as u64, cols: cols as u64 }
    }

    /// The synthetic-RGB arm of `save_processed_multiband_image_sequential` (save.rs:317-368) / `api/mod.rs:203-247`:
    /// per band dB -> autoscale (-> Tamed band step) -> resize + pad, then `create_synthetic_rgb_by_mode_and_strategy`.
    /// Returns (cols, rows, rgb, meta).
    pub fn synthetic_rgb(&self, band1: &Array2<f32>, band2: &Array2<f32>, strategy: AutoscaleStrategy, mode: SyntheticRgbMode,
                         target_size: Option<usize>, pad: bool, tamed_band_step: bool)
                         -> Result<(usize, usize, Vec<u8>, sys::sarpro_resize_meta), Box<dyn std::error::Error>> {
        let (rows, cols) = band1.dim();
        let (mut oc, mut or_) = (0usize, 0usize);
        self.check(unsafe { sys::sarpro_resize_output_dims(cols, rows, target_size.is_some() as i32, target_size.unwrap_or(0), pad as i32, &mut oc, &mut or_) })?;
        let mut rgb = vec![0u8; oc * or_ * 3];
        let (b1, b2) = (Self::band(band1), Self::band(band2));
        let mut img = sys::sarpro_image {
            data: rgb.as_mut_ptr() as *mut _, location: sys::SARPRO_LOC_HOST, bit_depth: 0, capacity_bytes: rgb.len() as u64,
            cols: 0, rows: 0, channels: 0, reserved: 0, meta: sys::sarpro_resize_meta::default(),
        };
        self.check(unsafe {
            sys::sarpro_pipeline_synrgb(self.0, &b1, &b2, strategy as i32, mode as i32, target_size.is_some() as i32, target_size.unwrap_or(0),
                                        pad as i32, tamed_band_step as i32, &mut img, std::ptr::null_mut())
        })?;
        Ok((img.cols as usize, img.rows as usize, rgb, img.meta))
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

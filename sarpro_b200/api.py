"""Host-side mirror of the reference's function surface for the raster path, over the C ABI.

Names, argument meaning and error behaviour follow src/core/processing/*.rs and the buffer variants
of src/api/mod.rs (cited per function), so parity tests read like tests of the reference itself.
Arrays are numpy (host) or torch CUDA tensors (device-resident; used by bench.py's kernel-only leg).
Nothing here computes pixels: every function is one call into libsarpro_gpu.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _ffi as F
from ._ffi import (ADAPTIVE, CLAHE, DEFAULT, EQUALIZED, JPEG, OP_DIFF, OP_LOGRATIO, OP_NDIFF, OP_NONE, OP_RATIO,  # noqa: F401
                   OP_SUM, ROBUST, STANDARD, TAMED, TIFF, U8, U16)


class SarproError(RuntimeError):
    """Error::External(String) of the reference (src/error.rs:43-47) plus the status code."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _host(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


@dataclass
class ProcessedImage:
    """api/mod.rs:51-62 (metadata omitted: SAFE metadata never enters the raster path)."""
    width: int
    height: int
    bit_depth: int
    format: int
    gray: np.ndarray | None = None
    gray16: np.ndarray | None = None
    rgb: np.ndarray | None = None
    gray_band2: np.ndarray | None = None
    gray16_band2: np.ndarray | None = None
    scale_x: float = 1.0
    scale_y: float = 1.0
    pad_left: int = 0
    pad_top: int = 0
    stats: list = field(default_factory=list)


class Context:
    """Owns one sarpro_ctx (one GPU, one stream). Not thread-safe, like a reference call chain."""

    def __init__(self, device: int = 0):
        self._lib = F.lib()
        h = C.c_void_p()
        rc = self._lib.sarpro_ctx_create(C.byref(h), int(device))
        if rc != F.OK:
            raise SarproError(rc, (self._lib.sarpro_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sarpro_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- helpers ------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != F.OK:
            raise SarproError(rc, (self._lib.sarpro_last_error(self._h) or b"").decode())

    def timing(self) -> F.Timing:
        t = F.Timing()
        self._check(self._lib.sarpro_last_timing(self._h, C.byref(t)))
        return t

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._lib.sarpro_ctx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def _band(self, a):
        """sarpro_band for a numpy array (host) or torch CUDA tensor (device). Keeps `a` alive via the return."""
        if _is_torch(a):
            import torch
            if not a.is_cuda or not a.is_contiguous():
                raise ValueError("device bands must be contiguous CUDA tensors")
            if a.dtype == torch.float32:
                dt = F.DT_F32
            elif a.dtype in (torch.uint16, torch.int16):
                dt = F.DT_U16
            else:
                raise ValueError(f"unsupported band dtype {a.dtype}")
            rows, cols = a.shape
            return F.Band(a.data_ptr(), dt, F.LOC_DEVICE, rows, cols), a
        if a.dtype == np.uint16:
            a = _host(a, np.uint16)
            dt = F.DT_U16
        else:
            a = _host(a, np.float32)
            dt = F.DT_F32
        rows, cols = a.shape
        return F.Band(a.ctypes.data, dt, F.LOC_HOST, rows, cols), a

    @staticmethod
    def resize_output_dims(cols, rows, target, pad):
        oc, orr = C.c_size_t(), C.c_size_t()
        F.lib().sarpro_resize_output_dims(cols, rows, int(target is not None), int(target or 0), int(bool(pad)),
                                          C.byref(oc), C.byref(orr))
        return oc.value, orr.value

    KEEP = "keep"  # as `out=`: the result stays in the context's device buffers (F.LOC_NONE), for encode_last_jpeg

    def _image(self, cols, rows, channels, bit_depth, out=None):
        dt = np.uint8 if bit_depth == U8 else np.uint16
        shape = (rows, cols) if channels == 1 else (rows, cols, channels)
        if isinstance(out, str) and out == Context.KEEP:
            return F.Image(None, F.LOC_NONE, bit_depth, 0, 0, 0, 0, 0, F.ResizeMeta()), None
        if out is None:
            out = np.empty(shape, dt)
            img = F.Image(out.ctypes.data, F.LOC_HOST, bit_depth, out.nbytes, 0, 0, 0, 0, F.ResizeMeta())
        elif _is_torch(out):
            img = F.Image(out.data_ptr(), F.LOC_DEVICE, bit_depth, out.numel() * out.element_size(), 0, 0, 0, 0, F.ResizeMeta())
        else:
            img = F.Image(out.ctypes.data, F.LOC_HOST, bit_depth, out.nbytes, 0, 0, 0, 0, F.ResizeMeta())
        return img, out

    # -- stage level (src/core/processing) ------------------------------------------------------
    def process_scalar_data_pipeline(self, processed, bit_depth, strategy):
        """pipeline.rs:42-66. Returns (scaled_u8 | None, scaled_u16 | None, stats).
        The (db_data, valid_mask) members of the reference's tuple are not materialised
        (see process_scalar_data_inplace)."""
        if processed.dtype == np.uint16:
            v = _host(processed, np.uint16)
            fn = self._lib.sarpro_process_dn_pipeline
        else:
            v = _host(processed, np.float32)
            fn = self._lib.sarpro_process_scalar_data_pipeline
        rows, cols = v.shape
        u8 = np.empty(v.shape, np.uint8) if bit_depth == U8 else None
        u16 = np.empty(v.shape, np.uint16) if bit_depth == U16 else None
        st = F.Stats()
        self._check(fn(self._h, _ptr(v), rows, cols, bit_depth, strategy, _ptr(u8), _ptr(u16), C.byref(st)))
        return u8, u16, st

    def process_scalar_data_inplace(self, processed):
        """pipeline.rs:8-40 -> (db f64, valid_mask u8)."""
        v = _host(processed, np.float32)
        rows, cols = v.shape
        db = np.empty(v.shape, np.float64)
        mask = np.empty(v.shape, np.uint8)
        self._check(self._lib.sarpro_process_scalar_data_inplace(self._h, _ptr(v), rows, cols, _ptr(db), _ptr(mask)))
        return db, mask

    def autoscale_db_image_tamed_synrgb_u8(self, processed, is_copol):
        """autoscale.rs:710-742 (takes the linear band the dB plane came from)."""
        v = _host(processed, np.float32)
        rows, cols = v.shape
        out = np.empty(v.shape, np.uint8)
        self._check(self._lib.sarpro_autoscale_tamed_synrgb_u8(self._h, _ptr(v), rows, cols, int(bool(is_copol)), _ptr(out)))
        return out

    def plan_on_device(self, hist65536, bit_depth, strategy, plan_kind=0):
        """Test hook: the device planner (kernels_plan.cu) on a DN histogram: (stats, lut, (hot, hot_top))."""
        h = np.ascontiguousarray(hist65536, dtype=np.uint32)
        assert h.size == 65536
        lut = np.zeros(65536, np.uint16)
        hot = np.zeros(2, np.uint32)
        st = F.Stats()
        self._check(self._lib.sarpro_plan_on_device(self._h, h.ctypes.data, bit_depth, strategy, plan_kind, C.byref(st),
                                                    lut.ctypes.data, hot.ctypes.data))
        return st, lut, (int(hot[0]), int(hot[1]))

    def scale_u16_to_u8(self, data):
        """autoscale.rs:348-364"""
        d = _host(data, np.uint16)
        out = np.empty(d.shape, np.uint8)
        self._check(self._lib.sarpro_scale_u16_to_u8(self._h, _ptr(d), d.size, _ptr(out)))
        return out

    def _pol(self, op, a, b):
        a = _host(a, np.float32)
        b = _host(b, np.float32)
        if a.shape != b.shape:
            raise ValueError("shape mismatch")
        rows, cols = a.shape
        out = np.empty(a.shape, np.float32)
        self._check(self._lib.sarpro_pol_op(self._h, op, _ptr(a), _ptr(b), rows, cols, _ptr(out)))
        return out

    def sum_arrays(self, a, b):              # ops.rs:4
        return self._pol(OP_SUM, a, b)

    def difference_arrays(self, a, b):       # ops.rs:7
        return self._pol(OP_DIFF, a, b)

    def ratio_arrays(self, a, b):            # ops.rs:10-20
        return self._pol(OP_RATIO, a, b)

    def normalized_diff_arrays(self, a, b):  # ops.rs:22-33
        return self._pol(OP_NDIFF, a, b)

    def log_ratio_arrays(self, a, b):        # ops.rs:35-44
        return self._pol(OP_LOGRATIO, a, b)

    @staticmethod
    def calculate_resize_dimensions(original_cols, original_rows, target_size):
        """resize.rs:6-30 (through the same C entry the pipelines use)."""
        if target_size > max(original_cols, original_rows):
            return original_cols, original_rows
        return Context.resize_output_dims(original_cols, original_rows, target_size, False)

    def resize_image_data_with_meta(self, data, target_size, bit_depth, pad):
        """resize.rs:91-236. `data` is the u8 plane (U8) or the u16 plane (U16); None for U16 raises
        the reference's "U16 data required for U16 bit depth"."""
        if data is None:
            if bit_depth == U16:
                meta = F.ResizeMeta()
                self._check(self._lib.sarpro_resize_image_data_with_meta(
                    self._h, None, None, 0, 0, int(target_size is not None), int(target_size or 0), U16, int(bool(pad)),
                    None, None, C.byref(meta)))
            raise ValueError("data is None")
        dt = np.uint8 if bit_depth == U8 else np.uint16
        d = _host(data, dt)
        rows, cols = d.shape
        oc, orr = self.resize_output_dims(cols, rows, target_size, pad)
        out = np.empty((orr, oc), dt)
        meta = F.ResizeMeta()
        self._check(self._lib.sarpro_resize_image_data_with_meta(
            self._h, _ptr(d) if bit_depth == U8 else None, _ptr(d) if bit_depth == U16 else None, cols, rows,
            int(target_size is not None), int(target_size or 0), bit_depth, int(bool(pad)),
            _ptr(out) if bit_depth == U8 else None, _ptr(out) if bit_depth == U16 else None, C.byref(meta)))
        return out, meta

    def resize_image_data(self, data, target_size, bit_depth, pad):
        """resize.rs:238-257"""
        out, meta = self.resize_image_data_with_meta(data, target_size, bit_depth, pad)
        return meta.cols, meta.rows, out

    def add_padding_to_square(self, data, bit_depth):
        """padding.rs:5-49"""
        if data is None and bit_depth == U16:
            self._check(self._lib.sarpro_add_padding_to_square(self._h, None, None, 0, 0, U16, None, None))
        dt = np.uint8 if bit_depth == U8 else np.uint16
        d = _host(data, dt)
        rows, cols = d.shape
        m = max(rows, cols)
        out = np.empty((m, m), dt)
        self._check(self._lib.sarpro_add_padding_to_square(
            self._h, _ptr(d) if bit_depth == U8 else None, _ptr(d) if bit_depth == U16 else None, cols, rows, bit_depth,
            _ptr(out) if bit_depth == U8 else None, _ptr(out) if bit_depth == U16 else None))
        return out

    def create_synthetic_rgb_by_mode_and_strategy(self, mode, strategy, band1_data, band2_data):
        """synthetic_rgb.rs:182-197"""
        b1 = _host(band1_data, np.uint8)
        b2 = _host(band2_data, np.uint8)
        if b1.shape != b2.shape:
            raise ValueError("shape mismatch")
        rgb = np.empty(b1.shape + (3,), np.uint8)
        self._check(self._lib.sarpro_create_synthetic_rgb_by_mode_and_strategy(
            self._h, mode, strategy, _ptr(b1), _ptr(b2), b1.size, _ptr(rgb)))
        return rgb

    # -- fused pipelines (save.rs / api/mod.rs call orders) ---------------------------------------
    def process_single(self, band, fmt, bit_depth, strategy, target_size=None, pad=False, op=OP_NONE, band2=None,
                       out=None) -> ProcessedImage:
        """save.rs:49-65/119-134; api/mod.rs:84-130, 250-281, 284-369 (op: sentinel1.rs:1497-1579)."""
        ba, keep_a = self._band(band)
        bb, keep_b = (self._band(band2) if band2 is not None else (None, None))
        if fmt == JPEG:
            bit_depth = U8
        rows, cols = keep_a.shape
        oc, orr = self.resize_output_dims(cols, rows, target_size, pad)
        img, out = self._image(oc, orr, 1, bit_depth, out)
        st = F.Stats()
        self._check(self._lib.sarpro_pipeline_single(
            self._h, C.byref(ba), C.byref(bb) if bb is not None else None, op, fmt, bit_depth, strategy,
            int(target_size is not None), int(target_size or 0), int(bool(pad)), C.byref(img), C.byref(st)))
        del keep_b
        return ProcessedImage(img.cols, img.rows, bit_depth, fmt,
                              gray=out if bit_depth == U8 else None, gray16=out if bit_depth == U16 else None,
                              scale_x=img.meta.scale_x, scale_y=img.meta.scale_y, pad_left=img.meta.pad_left,
                              pad_top=img.meta.pad_top, stats=[st])

    # -- multi-GPU: one scene row-band-sharded over the ranks of a communicator -----------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        self._check(self._lib.sarpro_comm_init(self._h, unique_id, int(rank), int(world)))

    def comm_destroy(self):
        self._check(self._lib.sarpro_comm_destroy(self._h))

    def process_synrgb_sharded(self, band1_rows, band2_rows, scene_rows, strategy, target_size, pad=False,
                               mode=F.SYNRGB_DEFAULT, tamed_band_step=True, out=None, want_output=True) -> ProcessedImage:
        """save.rs:317-368 on THIS rank's rows [h0,h1) of the scene (see shard_halo_rows); every rank gets the image."""
        b1, k1 = self._band(band1_rows)
        b2, k2 = self._band(band2_rows)
        _, cols = k1.shape
        oc, orr = self.resize_output_dims(cols, scene_rows, target_size, pad)
        if want_output:
            img, out = self._image(oc, orr, 3, U8, out)
        else:
            img, out = F.Image(None, F.LOC_HOST, U8, 0, 0, 0, 0, 0, F.ResizeMeta()), None
        self._check(self._lib.sarpro_pipeline_synrgb_sharded(
            self._h, C.byref(b1), C.byref(b2), int(scene_rows), strategy, mode, int(target_size is not None),
            int(target_size or 0), int(bool(pad)), int(bool(tamed_band_step)), C.byref(img)))
        del k2
        return ProcessedImage(img.cols, img.rows, U8, JPEG, rgb=out, scale_x=img.meta.scale_x, scale_y=img.meta.scale_y,
                              pad_left=img.meta.pad_left, pad_top=img.meta.pad_top)

    def process_polops(self, band_a, band_b, ops, bit_depth, strategy, scene_rows=None, outs=None):
        """One or two polarization operations over the same pair, each autoscaled into a full-resolution band
        (sentinel1.rs:1497-1579 -> ops.rs:4-44 -> pipeline.rs:42-66; BASELINE config 4 with ops = (OP_LOGRATIO, OP_NDIFF)).
        scene_rows: the bands are this rank's row band of a scene of that many rows (comm_init first).
        Returns ([plane per op], [stats per op])."""
        ba, ka = self._band(band_a)
        bb, kb = self._band(band_b)
        rows, cols = ka.shape
        ops = list(ops)
        n = len(ops)
        imgs = (F.Image * n)()
        keep = []
        for o in range(n):
            img, out = self._image(cols, rows, 1, bit_depth, None if outs is None else outs[o])
            imgs[o] = img
            keep.append(out)
        st = (F.Stats * n)()
        arr = (C.c_int * n)(*ops)
        self._check(self._lib.sarpro_pipeline_polops(self._h, C.byref(ba), C.byref(bb), int(scene_rows or 0), n, arr, bit_depth, strategy,
                                                     imgs, st))
        del kb
        return keep, [st[o] for o in range(n)]

    def encode_jpeg(self, image, quality=100):
        """write_gray_jpeg / write_rgb_jpeg (io/writers/jpeg.rs:6-30) without the file: u8 (rows, cols) or (rows, cols, 3) array
        (numpy or torch CUDA) -> bytes of a baseline JPEG stream (nvJPEG on the GPU)."""
        if _is_torch(image):
            shape, loc, keep = tuple(image.shape), F.LOC_DEVICE, image
        else:
            keep = np.ascontiguousarray(image, dtype=np.uint8)
            shape, loc = keep.shape, F.LOC_HOST
        ch = 1 if len(shape) == 2 else shape[2]
        img = F.Image(_ptr(keep), loc, U8, shape[0] * shape[1] * ch, shape[1], shape[0], ch, 0, F.ResizeMeta())
        n = C.c_size_t()
        self._check(self._lib.sarpro_encode_jpeg(self._h, C.byref(img), quality, None, 0, C.byref(n)))
        buf = (C.c_ubyte * n.value)()
        self._check(self._lib.sarpro_encode_jpeg(self._h, C.byref(img), quality, buf, n.value, C.byref(n)))
        return bytes(buf[: n.value])

    def encode_last_jpeg(self, which=0, quality=100, capacity=None):
        """JPEG of the u8 result the last pipeline call left on the device (0 = RGB, 1 / 2 = gray bands): only the stream
        crosses PCIe. capacity: size of the receiving buffer (default: queried first)."""
        n = C.c_size_t()
        if capacity is None:
            self._check(self._lib.sarpro_encode_last_jpeg(self._h, which, quality, None, 0, C.byref(n)))
            capacity = n.value
        buf = (C.c_ubyte * capacity)()
        self._check(self._lib.sarpro_encode_last_jpeg(self._h, which, quality, buf, capacity, C.byref(n)))
        return bytes(buf[: n.value])

    @staticmethod
    def read_dims_for_target(cols, rows, target):
        """sentinel1.rs:1083-1102 -> (out_cols, out_rows, resampler)."""
        oc, orr, alg = C.c_size_t(), C.c_size_t(), C.c_int()
        rc = F.lib().sarpro_read_dims_for_target(cols, rows, target, C.byref(oc), C.byref(orr), C.byref(alg))
        if rc != F.OK:
            raise SarproError(rc, "sarpro_read_dims_for_target: invalid argument")
        return oc.value, orr.value, alg.value

    def read_band_resampled(self, band, out_cols, out_rows, alg, out=None):
        """GdalSarReader::read_band_resampled (gdal.rs:145-177) on the GPU: u16 / f32 raster (numpy or torch CUDA tensor) ->
        f32 (out_rows, out_cols); `out` may be a torch CUDA f32 tensor (stays on the device for the pipelines)."""
        b, keep = self._band(band)
        if out is None:
            out = np.empty((out_rows, out_cols), np.float32)
        loc = F.LOC_DEVICE if _is_torch(out) else F.LOC_HOST
        self._check(self._lib.sarpro_read_band_resampled(self._h, C.byref(b), out_cols, out_rows, alg, _ptr(out), loc))
        del keep
        return out

    def process_batch(self, scenes, kind, bit_depth, strategy, target_size=None, pad=False, mode=F.SYNRGB_DEFAULT,
                      tamed_band_step=True, continue_on_error=True, outs=None):
        """process_directory_to_path's scene loop (api/mod.rs:474-536) over decoded band pairs: `scenes` is a list of
        (band1, band2) or None (a product the reader skips). kind: F.BATCH_MULTIBAND (two gray bands per scene) or
        F.BATCH_SYNRGB (one RGB image per scene). The next scene's upload runs beside the current scene's kernels.
        outs: optional list of preallocated outputs (numpy / torch), 2 per scene for MULTIBAND, 1 for SYNRGB.
        Returns (list of outputs per scene | None for skipped / failed, statuses, report, stats)."""
        n = len(scenes)
        per = 2 if kind == F.BATCH_MULTIBAND else 1
        arr = (F.Scene * max(n, 1))()
        imgs = (F.Image * max(per * n, 1))()
        keep, results = [], []
        depth = bit_depth if kind == F.BATCH_MULTIBAND else U8
        for k, sc in enumerate(scenes):
            if sc is None:
                arr[k] = F.Scene(F.Band(None, F.DT_U16, F.LOC_HOST, 0, 0), F.Band(None, F.DT_U16, F.LOC_HOST, 0, 0))
                results.append(None)
                continue
            b1, k1 = self._band(sc[0])
            b2, k2 = self._band(sc[1])
            keep += [k1, k2]
            arr[k] = F.Scene(b1, b2)
            rows, cols = k1.shape
            oc, orr = self.resize_output_dims(cols, rows, target_size, pad)
            res = []
            for j in range(per):
                img, o = self._image(oc, orr, 1 if kind == F.BATCH_MULTIBAND else 3, depth, None if outs is None else outs[per * k + j])
                imgs[per * k + j] = img
                res.append(o)
            results.append(res)
        st = (F.Stats * max(2 * n, 1))()
        status = (C.c_int * max(n, 1))()
        rep = F.BatchReport()
        self._check(self._lib.sarpro_pipeline_batch(self._h, arr, n, kind, bit_depth, strategy, mode, int(target_size is not None),
                                                    int(target_size or 0), int(bool(pad)), int(bool(tamed_band_step)),
                                                    int(bool(continue_on_error)), imgs, st, status, C.byref(rep)))
        del keep
        statuses = [status[k] for k in range(n)]
        for k in range(n):
            if statuses[k] != F.OK:
                results[k] = None
        return results, statuses, rep, [st[i] for i in range(2 * n)]

    def process_multiband_tiff(self, band1, band2, bit_depth, strategy, target_size=None, pad=False) -> ProcessedImage:
        """save.rs:199-316; api/mod.rs:133-200."""
        b1, k1 = self._band(band1)
        b2, k2 = self._band(band2)
        rows, cols = k1.shape
        oc, orr = self.resize_output_dims(cols, rows, target_size, pad)
        i1, o1 = self._image(oc, orr, 1, bit_depth)
        i2, o2 = self._image(oc, orr, 1, bit_depth)
        st = (F.Stats * 2)()
        self._check(self._lib.sarpro_pipeline_multiband_tiff(
            self._h, C.byref(b1), C.byref(b2), bit_depth, strategy, int(target_size is not None), int(target_size or 0),
            int(bool(pad)), C.byref(i1), C.byref(i2), st))
        del k2
        u8 = bit_depth == U8
        return ProcessedImage(i1.cols, i1.rows, bit_depth, TIFF, gray=o1 if u8 else None, gray16=None if u8 else o1,
                              gray_band2=o2 if u8 else None, gray16_band2=None if u8 else o2,
                              scale_x=i1.meta.scale_x, scale_y=i1.meta.scale_y, pad_left=i1.meta.pad_left,
                              pad_top=i1.meta.pad_top, stats=[st[0], st[1]])

    def process_synrgb_jpeg(self, band1, band2, strategy, target_size=None, pad=False, mode=F.SYNRGB_DEFAULT,
                            tamed_band_step=True, out=None) -> ProcessedImage:
        """save.rs:317-368 (tamed_band_step=True) / api/mod.rs:203-247 (False)."""
        b1, k1 = self._band(band1)
        b2, k2 = self._band(band2)
        rows, cols = k1.shape
        oc, orr = self.resize_output_dims(cols, rows, target_size, pad)
        img, out = self._image(oc, orr, 3, U8, out)
        st = (F.Stats * 2)()
        self._check(self._lib.sarpro_pipeline_synrgb(
            self._h, C.byref(b1), C.byref(b2), strategy, mode, int(target_size is not None), int(target_size or 0),
            int(bool(pad)), int(bool(tamed_band_step)), C.byref(img), st))
        del k2
        return ProcessedImage(img.cols, img.rows, U8, JPEG, rgb=out, scale_x=img.meta.scale_x, scale_y=img.meta.scale_y,
                              pad_left=img.meta.pad_left, pad_top=img.meta.pad_top, stats=[st[0], st[1]])


def comm_unique_id() -> bytes:
    """128-byte ncclUniqueId (rank 0 creates it; ship it to the other ranks with any transport)."""
    buf = C.create_string_buffer(128)
    rc = F.lib().sarpro_comm_unique_id(buf)
    if rc != F.OK:
        raise SarproError(rc, "sarpro_comm_unique_id failed (libnccl.so.2 not loadable? set SARPRO_NCCL_LIB)")
    return buf.raw


def shard_halo_rows(rows, cols, target, world, rank, clahe):
    """Scene rows [h0,h1) a rank must hold: its own band plus the vertical Lanczos halo."""
    h0, h1 = C.c_size_t(), C.c_size_t()
    rc = F.lib().sarpro_shard_halo_rows(rows, cols, int(target is not None), int(target or 0), world, rank, int(bool(clahe)),
                                        C.byref(h0), C.byref(h1))
    if rc != F.OK:
        raise SarproError(rc, "sarpro_shard_halo_rows: invalid argument")
    return h0.value, h1.value


def plan_from_dn_histogram(hist65536, bit_depth, strategy):
    """Host-only planner entry (no GPU needed): stats + DN->sample LUT from a DN histogram."""
    h = np.ascontiguousarray(hist65536, dtype=np.uint64)
    assert h.size == 65536
    lut = np.zeros(65536, np.uint16)
    st = F.Stats()
    rc = F.lib().sarpro_plan_from_dn_histogram(h.ctypes.data, bit_depth, strategy, C.byref(st), lut.ctypes.data)
    if rc != F.OK:
        raise SarproError(rc, "sarpro_plan_from_dn_histogram failed")
    return st, lut


def plan_kind_from_dn_histogram(hist65536, bit_depth, strategy, plan_kind=0):
    """Host planner with the plan kind (0 autoscale, 1 / 2 Tamed-synRGB co- / cross-pol): (stats, lut, (hot, hot_top))."""
    h = np.ascontiguousarray(hist65536, dtype=np.uint64)
    assert h.size == 65536
    lut = np.zeros(65536, np.uint16)
    hot = np.zeros(2, np.uint32)
    st = F.Stats()
    rc = F.lib().sarpro_plan_kind_from_dn_histogram(h.ctypes.data, bit_depth, strategy, plan_kind, C.byref(st), lut.ctypes.data,
                                                    hot.ctypes.data)
    if rc != F.OK:
        raise SarproError(rc, "sarpro_plan_kind_from_dn_histogram failed")
    return st, lut, (int(hot[0]), int(hot[1]))


def present_list_from_histogram(hist65536, cap, order=None):
    """Host emulation of the list k_hist_total writes for the planner: (blocks[256, 2], pairs[cap, 2]) of uint32. `order` is
    the order in which the 256 blocks allocate their pairs (any permutation; the device's is arbitrary)."""
    h = np.asarray(hist65536).astype(np.uint64).reshape(256, 256)
    blocks = np.zeros((256, 2), np.uint32)
    pairs = np.full((cap, 2), 0xDEADBEEF, np.uint32)
    alloc = 0
    for b in (range(256) if order is None else order):
        (nz,) = np.nonzero(h[b])
        blocks[b] = (alloc if nz.size else 0, nz.size)
        keep = nz[: max(0, min(nz.size, cap - alloc))]  # pairs beyond cap are dropped
        pairs[alloc:alloc + keep.size, 0] = b * 256 + keep
        pairs[alloc:alloc + keep.size, 1] = h[b][keep]
        alloc += nz.size
    return blocks, pairs


def plan_from_present_list(blocks, pairs, bit_depth, strategy):
    """Host-only planner entry on the device-compacted list of present DNs. Returns (stats, lut) or None when the list
    overflowed its capacity."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint32)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32)
    assert blocks.shape == (256, 2) and pairs.ndim == 2 and pairs.shape[1] == 2
    lut = np.zeros(65536, np.uint16)
    st = F.Stats()
    rc = F.lib().sarpro_plan_from_present_list(blocks.ctypes.data, pairs.ctypes.data, pairs.shape[0], bit_depth, strategy,
                                               C.byref(st), lut.ctypes.data)
    if rc < 0:
        raise SarproError(rc, "sarpro_plan_from_present_list failed")
    return (st, lut) if rc == 1 else None


def shard_rows(rows, world, rank, clahe):
    r0, r1 = C.c_size_t(), C.c_size_t()
    rc = F.lib().sarpro_shard_rows(rows, world, rank, int(bool(clahe)), C.byref(r0), C.byref(r1))
    if rc != F.OK:
        raise SarproError(rc, "sarpro_shard_rows: invalid argument")
    return r0.value, r1.value

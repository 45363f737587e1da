"""ctypes loader for the CPU ORACLE (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY. May be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, and from nowhere else. The product package
(sarpro_b200) never imports this module.

PARITY UNPINNED: see oracle/oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

STANDARD, ROBUST, ADAPTIVE, EQUALIZED, CLAHE, TAMED, DEFAULT = range(7)
STRATEGY_NAMES = ["standard", "robust", "adaptive", "equalized", "clahe", "tamed", "default"]
U8, U16 = 0, 1
OP_SUM, OP_DIFF, OP_RATIO, OP_NDIFF, OP_LOGRATIO = range(5)
TIFF, JPEG = 0, 1


class Stats(C.Structure):
    _fields_ = [("valid_count", C.c_uint64)] + [
        (n, C.c_double)
        for n in (
            "min_db", "max_db", "mean_db", "std_db", "median_db",
            "p01", "p02", "p05", "p10", "p25", "p75", "p90", "p95", "p98", "p99",
            "low_clip", "high_clip", "gamma",
        )
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ResizeMeta(C.Structure):
    _fields_ = [
        ("cols", C.c_uint64), ("rows", C.c_uint64),
        ("scale_x", C.c_double), ("scale_y", C.c_double),
        ("pad_left", C.c_uint64), ("pad_top", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle_resize.cpp", "oracle_read.cpp", "oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_resize_u8_image.restype = C.c_int
        _lib.oracle_resize_u16_image.restype = C.c_int
        _lib.oracle_add_padding_to_square.restype = C.c_int
        _lib.oracle_resize_image_data_with_meta.restype = C.c_int
        _lib.oracle_pipeline_single.restype = C.c_int
        _lib.oracle_pipeline_multiband_tiff.restype = C.c_int
        _lib.oracle_pipeline_synrgb_jpeg.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _sz(x):
    return C.c_size_t(int(x))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# --------------------------------------------------------------------------- stage level
def process_scalar_data_inplace(v):
    v = _c(v, np.float32)
    db = np.empty(v.shape, np.float64)
    mask = np.empty(v.shape, np.uint8)
    lib().oracle_process_scalar_data_inplace(_p(v), _sz(v.size), _p(db), _p(mask))
    return db, mask


def compute_histogram_stats(db, mask):
    db = _c(db, np.float64)
    mask = _c(mask, np.uint8)
    st = Stats()
    hist = np.zeros(4096, np.uint64)
    rows, cols = db.shape
    lib().oracle_compute_histogram_stats(_p(db), _p(mask), _sz(rows), _sz(cols), C.byref(st), _p(hist))
    return st, hist


def clahe_equalize_normalized(norm, mask, tiles_x=8, tiles_y=8, clip_limit=2.0, num_bins=256):
    norm = _c(norm, np.float64)
    mask = _c(mask, np.uint8)
    rows, cols = norm.shape
    out = np.empty_like(norm)
    cdfs = np.zeros((tiles_y * tiles_x, num_bins), np.float64)
    lib().oracle_clahe_equalize_normalized(
        _p(norm), _p(mask), _sz(rows), _sz(cols), _sz(tiles_x), _sz(tiles_y),
        C.c_double(clip_limit), _sz(num_bins), _p(out), _p(cdfs))
    return out, cdfs


def scale_u16_to_u8(data):
    data = _c(data, np.uint16)
    out = np.empty(data.shape, np.uint8)
    lib().oracle_scale_u16_to_u8(_p(data), _sz(data.size), _p(out))
    return out


def autoscale_db_image(db, mask, bit_depth):
    db = _c(db, np.float64); mask = _c(mask, np.uint8)
    rows, cols = db.shape
    out = np.empty(db.shape, np.uint16)
    st = Stats()
    lib().oracle_autoscale_db_image(_p(db), _p(mask), _sz(rows), _sz(cols), C.c_int(bit_depth), _p(out), C.byref(st))
    return out, st


def autoscale_db_image_advanced(db, mask, bit_depth, strategy):
    db = _c(db, np.float64); mask = _c(mask, np.uint8)
    rows, cols = db.shape
    out = np.empty(db.shape, np.uint16)
    st = Stats()
    lib().oracle_autoscale_db_image_advanced(
        _p(db), _p(mask), _sz(rows), _sz(cols), C.c_int(bit_depth), C.c_int(strategy), _p(out), C.byref(st))
    return out, st


def autoscale_db_image_tamed_synrgb_u8(db, mask, is_copol):
    db = _c(db, np.float64); mask = _c(mask, np.uint8)
    rows, cols = db.shape
    out = np.empty(db.shape, np.uint8)
    lib().oracle_autoscale_db_image_tamed_synrgb_u8(_p(db), _p(mask), _sz(rows), _sz(cols), C.c_int(int(is_copol)), _p(out))
    return out


@dataclass
class PipelineOut:
    db: np.ndarray
    mask: np.ndarray
    u8: np.ndarray | None
    u16: np.ndarray | None
    stats: Stats


def process_scalar_data_pipeline(v, bit_depth, strategy, want_db=True):
    """pipeline.rs:42-66 — returns (db, mask, scaled_u8, scaled_u16)."""
    v = _c(v, np.float32)
    rows, cols = v.shape
    db = np.empty(v.shape, np.float64) if want_db else None
    mask = np.empty(v.shape, np.uint8) if want_db else None
    u8 = np.empty(v.shape, np.uint8) if bit_depth == U8 else None
    u16 = np.empty(v.shape, np.uint16) if bit_depth == U16 else None
    st = Stats()
    lib().oracle_process_scalar_data_pipeline(
        _p(v), _sz(rows), _sz(cols), C.c_int(bit_depth), C.c_int(strategy), _p(db), _p(mask), _p(u8), _p(u16), C.byref(st))
    return PipelineOut(db, mask, u8, u16, st)


def pol_op(op, a, b):
    a = _c(a, np.float32); b = _c(b, np.float32)
    out = np.empty(a.shape, np.float32)
    lib().oracle_pol_op(C.c_int(op), _p(a), _p(b), _sz(a.size), _p(out))
    return out


def calculate_resize_dimensions(cols, rows, target):
    nc, nr = C.c_size_t(), C.c_size_t()
    lib().oracle_calculate_resize_dimensions(_sz(cols), _sz(rows), _sz(target), C.byref(nc), C.byref(nr))
    return nc.value, nr.value


def resize_u8_image(data, tcols, trows):
    data = _c(data, np.uint8)
    rows, cols = data.shape
    out = np.empty((trows, tcols), np.uint8)
    rc = lib().oracle_resize_u8_image(_p(data), _sz(cols), _sz(rows), _sz(tcols), _sz(trows), _p(out))
    if rc:
        raise RuntimeError("oracle resize_u8_image failed")
    return out


def resize_u16_image(data, tcols, trows):
    data = _c(data, np.uint16)
    rows, cols = data.shape
    out = np.empty((trows, tcols), np.uint16)
    rc = lib().oracle_resize_u16_image(_p(data), _sz(cols), _sz(rows), _sz(tcols), _sz(trows), _p(out))
    if rc:
        raise RuntimeError("oracle resize_u16_image failed")
    return out


def resize_output_dims(cols, rows, target, pad):
    oc, orr = C.c_size_t(), C.c_size_t()
    lib().oracle_resize_output_dims(
        _sz(cols), _sz(rows), C.c_int(target is not None), _sz(target or 0), C.c_int(int(pad)), C.byref(oc), C.byref(orr))
    return oc.value, orr.value


def add_padding_to_square(data, bit_depth):
    rows, cols = data.shape
    m = max(rows, cols)
    if bit_depth == U8:
        data = _c(data, np.uint8)
        out = np.empty((m, m), np.uint8)
        rc = lib().oracle_add_padding_to_square(_p(data), None, _sz(cols), _sz(rows), C.c_int(U8), _p(out), None)
    else:
        data = _c(data, np.uint16)
        out = np.empty((m, m), np.uint16)
        rc = lib().oracle_add_padding_to_square(None, _p(data), _sz(cols), _sz(rows), C.c_int(U16), None, _p(out))
    if rc:
        raise RuntimeError("U16 data required for U16 bit depth")
    return out


def resize_image_data_with_meta(data, target, bit_depth, pad):
    """resize.rs:91-236 — data is a 2-D u8 (U8) or u16 (U16) array."""
    rows, cols = data.shape
    oc, orr = resize_output_dims(cols, rows, target, pad)
    meta = ResizeMeta()
    if bit_depth == U8:
        data = _c(data, np.uint8)
        out = np.empty((orr, oc), np.uint8)
        rc = lib().oracle_resize_image_data_with_meta(
            _p(data), None, _sz(cols), _sz(rows), C.c_int(target is not None), _sz(target or 0),
            C.c_int(U8), C.c_int(int(pad)), _p(out), None, C.byref(meta))
    else:
        data = _c(data, np.uint16)
        out = np.empty((orr, oc), np.uint16)
        rc = lib().oracle_resize_image_data_with_meta(
            None, _p(data), _sz(cols), _sz(rows), C.c_int(target is not None), _sz(target or 0),
            C.c_int(U16), C.c_int(int(pad)), None, _p(out), C.byref(meta))
    if rc:
        raise RuntimeError(f"oracle resize_image_data_with_meta failed rc={rc}")
    return out, meta


def create_synthetic_rgb(b1, b2):
    b1 = _c(b1, np.uint8); b2 = _c(b2, np.uint8)
    rgb = np.empty(b1.shape + (3,), np.uint8)
    lib().oracle_create_synthetic_rgb(_p(b1), _p(b2), _sz(b1.size), _p(rgb))
    return rgb


def create_synthetic_rgb_suppressed(b1, b2):
    b1 = _c(b1, np.uint8); b2 = _c(b2, np.uint8)
    rgb = np.empty(b1.shape + (3,), np.uint8)
    lib().oracle_create_synthetic_rgb_suppressed(_p(b1), _p(b2), _sz(b1.size), _p(rgb))
    return rgb


def create_synthetic_rgb_by_mode_and_strategy(mode, strategy, b1, b2):
    b1 = _c(b1, np.uint8); b2 = _c(b2, np.uint8)
    rgb = np.empty(b1.shape + (3,), np.uint8)
    lib().oracle_create_synthetic_rgb_by_mode_and_strategy(C.c_int(mode), C.c_int(strategy), _p(b1), _p(b2), _sz(b1.size), _p(rgb))
    return rgb


# --------------------------------------------------------------------------- orchestration
def pipeline_single(v, fmt, bit_depth, strategy, target, pad):
    v = _c(v, np.float32)
    rows, cols = v.shape
    if fmt == JPEG:
        bit_depth = U8
    oc, orr = resize_output_dims(cols, rows, target, pad)
    out = np.empty((orr, oc), np.uint8 if bit_depth == U8 else np.uint16)
    meta = ResizeMeta()
    rc = lib().oracle_pipeline_single(
        _p(v), _sz(rows), _sz(cols), C.c_int(fmt), C.c_int(bit_depth), C.c_int(strategy),
        C.c_int(target is not None), _sz(target or 0), C.c_int(int(pad)),
        _p(out) if bit_depth == U8 else None, _p(out) if bit_depth == U16 else None, C.byref(meta))
    if rc:
        raise RuntimeError(f"oracle pipeline_single rc={rc}")
    return out, meta


def pipeline_multiband_tiff(v1, v2, bit_depth, strategy, target, pad):
    v1 = _c(v1, np.float32); v2 = _c(v2, np.float32)
    rows, cols = v1.shape
    oc, orr = resize_output_dims(cols, rows, target, pad)
    dt = np.uint8 if bit_depth == U8 else np.uint16
    o1 = np.empty((orr, oc), dt); o2 = np.empty((orr, oc), dt)
    meta = ResizeMeta()
    u8 = bit_depth == U8
    rc = lib().oracle_pipeline_multiband_tiff(
        _p(v1), _p(v2), _sz(rows), _sz(cols), C.c_int(bit_depth), C.c_int(strategy),
        C.c_int(target is not None), _sz(target or 0), C.c_int(int(pad)),
        _p(o1) if u8 else None, None if u8 else _p(o1), _p(o2) if u8 else None, None if u8 else _p(o2), C.byref(meta))
    if rc:
        raise RuntimeError(f"oracle pipeline_multiband_tiff rc={rc}")
    return o1, o2, meta


def pipeline_synrgb_jpeg(v1, v2, strategy, target, pad, mode=0, tamed_band_step=True):
    v1 = _c(v1, np.float32); v2 = _c(v2, np.float32)
    rows, cols = v1.shape
    oc, orr = resize_output_dims(cols, rows, target, pad)
    rgb = np.empty((orr, oc, 3), np.uint8)
    meta = ResizeMeta()
    rc = lib().oracle_pipeline_synrgb_jpeg(
        _p(v1), _p(v2), _sz(rows), _sz(cols), C.c_int(strategy), C.c_int(mode),
        C.c_int(target is not None), _sz(target or 0), C.c_int(int(pad)), C.c_int(int(tamed_band_step)),
        _p(rgb), C.byref(meta))
    if rc:
        raise RuntimeError(f"oracle pipeline_synrgb_jpeg rc={rc}")
    return rgb, meta


def set_resize_threads(n):
    lib().oracle_set_resize_threads(C.c_int(int(n)))


RESAMPLE_AVERAGE, RESAMPLE_LANCZOS = 0, 1


def read_dims_for_target(cols, rows, target):
    """sentinel1.rs:1083-1102 -> (out_cols, out_rows, alg)."""
    oc, orr, alg = C.c_size_t(), C.c_size_t(), C.c_int()
    lib().oracle_read_dims_for_target(_sz(cols), _sz(rows), _sz(target), C.byref(oc), C.byref(orr), C.byref(alg))
    return oc.value, orr.value, alg.value


def read_band_resampled(src, out_cols, out_rows, alg):
    """gdal.rs:145-177 (restated GDAL RasterIO resampling, parity unpinned): u16 or f32 raster -> f32 (out_rows, out_cols)."""
    rows, cols = src.shape
    out = np.empty((out_rows, out_cols), np.float32)
    if src.dtype == np.uint16:
        s = _c(src, np.uint16)
        lib().oracle_read_band_resampled_u16(_p(s), _sz(rows), _sz(cols), _sz(out_cols), _sz(out_rows), C.c_int(alg), _p(out))
    else:
        s = _c(src, np.float32)
        lib().oracle_read_band_resampled_f32(_p(s), _sz(rows), _sz(cols), _sz(out_cols), _sz(out_rows), C.c_int(alg), _p(out))
    return out

"""ctypes binding of libsarpro_gpu.so (include/sarpro_gpu.h). Loads the in-tree library only.

There is no fallback: if the library has not been built (python -m sarpro_b200.build or
__graft_entry__.build()) importing the symbols raises, and without a CUDA device every context
creation raises SarproError(NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SARPRO_GPU_LIB: an experimental build of the same library (sarpro_b200.build --variant=...), for A/B measurements
LIB_PATH = os.environ.get("SARPRO_GPU_LIB") or os.path.join(_HERE, "libsarpro_gpu.so")

# enums (Rust declaration order, src/types.rs)
STANDARD, ROBUST, ADAPTIVE, EQUALIZED, CLAHE, TAMED, DEFAULT = range(7)
STRATEGY_NAMES = ["standard", "robust", "adaptive", "equalized", "clahe", "tamed", "default"]
U8, U16 = 0, 1
OP_NONE, OP_SUM, OP_DIFF, OP_RATIO, OP_NDIFF, OP_LOGRATIO = -1, 0, 1, 2, 3, 4
TIFF, JPEG = 0, 1
DT_F32, DT_U16 = 0, 1
LOC_HOST, LOC_DEVICE, LOC_NONE = 0, 1, 2
SYNRGB_DEFAULT, SYNRGB_RGB_RATIO, SYNRGB_SAR_URBAN, SYNRGB_ENHANCED = range(4)

OK = 0
ERR_INVALID_ARGUMENT, ERR_NO_DEVICE, ERR_CUDA, ERR_OUT_OF_MEMORY = -1, -2, -3, -4
ERR_U16_REQUIRED, ERR_TOO_LARGE, ERR_COMM, ERR_INTERNAL = -5, -6, -7, -8


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


class Band(C.Structure):
    _fields_ = [
        ("data", C.c_void_p), ("dtype", C.c_int32), ("location", C.c_int32),
        ("rows", C.c_uint64), ("cols", C.c_uint64),
    ]


class Image(C.Structure):
    _fields_ = [
        ("data", C.c_void_p), ("location", C.c_int32), ("bit_depth", C.c_int32),
        ("capacity_bytes", C.c_uint64), ("cols", C.c_uint64), ("rows", C.c_uint64),
        ("channels", C.c_int32), ("reserved", C.c_int32), ("meta", ResizeMeta),
    ]


class Scene(C.Structure):
    _fields_ = [("b1", Band), ("b2", Band)]


class BatchReport(C.Structure):
    _fields_ = [("processed", C.c_uint64), ("skipped", C.c_uint64), ("errors", C.c_uint64)]


BATCH_MULTIBAND, BATCH_SYNRGB = 0, 1
RESAMPLE_AVERAGE, RESAMPLE_LANCZOS = 0, 1


class Timing(C.Structure):
    _fields_ = [
        ("total_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("kernel_ms", C.c_float),
        ("kernel_launches", C.c_uint32), ("host_syncs", C.c_uint32),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("stage_ms", C.c_float * 8), ("stage_launches", C.c_uint32 * 8),
    ]


STAGE_NAMES = ["hist", "plan", "apply", "vresize", "rgb", "convert", "comm", "other"]


# every symbol include/sarpro_gpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SZ = C.c_size_t
_I = C.c_int
SYMBOLS = {
    "sarpro_abi_version": (_I, []),
    "sarpro_ctx_create": (_I, [C.POINTER(_P), _I]),
    "sarpro_ctx_destroy": (None, [_P]),
    "sarpro_last_error": (C.c_char_p, [_P]),
    "sarpro_ctx_set_stream": (_I, [_P, _P]),
    "sarpro_ctx_synchronize": (_I, [_P]),
    "sarpro_last_timing": (_I, [_P, C.POINTER(Timing)]),
    "sarpro_host_alloc": (_I, [C.POINTER(_P), _SZ]),
    "sarpro_host_free": (None, [_P]),
    "sarpro_host_register": (_I, [_P, _SZ]),
    "sarpro_host_unregister": (None, [_P]),
    "sarpro_pol_op": (_I, [_P, _I, _P, _P, _SZ, _SZ, _P]),
    "sarpro_process_scalar_data_pipeline": (_I, [_P, _P, _SZ, _SZ, _I, _I, _P, _P, C.POINTER(Stats)]),
    "sarpro_process_dn_pipeline": (_I, [_P, _P, _SZ, _SZ, _I, _I, _P, _P, C.POINTER(Stats)]),
    "sarpro_process_scalar_data_inplace": (_I, [_P, _P, _SZ, _SZ, _P, _P]),
    "sarpro_autoscale_tamed_synrgb_u8": (_I, [_P, _P, _SZ, _SZ, _I, _P]),
    "sarpro_scale_u16_to_u8": (_I, [_P, _P, _SZ, _P]),
    "sarpro_resize_output_dims": (_I, [_SZ, _SZ, _I, _SZ, _I, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "sarpro_resize_image_data_with_meta": (_I, [_P, _P, _P, _SZ, _SZ, _I, _SZ, _I, _I, _P, _P, C.POINTER(ResizeMeta)]),
    "sarpro_add_padding_to_square": (_I, [_P, _P, _P, _SZ, _SZ, _I, _P, _P]),
    "sarpro_create_synthetic_rgb_by_mode_and_strategy": (_I, [_P, _I, _I, _P, _P, _SZ, _P]),
    "sarpro_pipeline_single": (_I, [_P, C.POINTER(Band), C.POINTER(Band), _I, _I, _I, _I, _I, _SZ, _I, C.POINTER(Image), C.POINTER(Stats)]),
    "sarpro_pipeline_multiband_tiff": (_I, [_P, C.POINTER(Band), C.POINTER(Band), _I, _I, _I, _SZ, _I, C.POINTER(Image), C.POINTER(Image), C.POINTER(Stats)]),
    "sarpro_pipeline_synrgb": (_I, [_P, C.POINTER(Band), C.POINTER(Band), _I, _I, _I, _SZ, _I, _I, C.POINTER(Image), C.POINTER(Stats)]),
    "sarpro_comm_unique_id": (_I, [_P]),
    "sarpro_comm_init": (_I, [_P, _P, _I, _I]),
    "sarpro_comm_destroy": (_I, [_P]),
    "sarpro_shard_rows": (_I, [_SZ, _I, _I, _I, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "sarpro_shard_halo_rows": (_I, [_SZ, _SZ, _I, _SZ, _I, _I, _I, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "sarpro_pipeline_synrgb_sharded": (_I, [_P, C.POINTER(Band), C.POINTER(Band), _SZ, _I, _I, _I, _SZ, _I, _I, C.POINTER(Image)]),
    "sarpro_pipeline_polops": (_I, [_P, C.POINTER(Band), C.POINTER(Band), _SZ, _I, C.POINTER(C.c_int), _I, _I, C.POINTER(Image), C.POINTER(Stats)]),
    "sarpro_pipeline_single_sharded": (_I, [_P, C.POINTER(Band), C.POINTER(Band), _SZ, _I, _I, _I, C.POINTER(Image), C.POINTER(Stats)]),
    "sarpro_read_dims_for_target": (_I, [_SZ, _SZ, _SZ, C.POINTER(_SZ), C.POINTER(_SZ), C.POINTER(C.c_int)]),
    "sarpro_read_band_resampled": (_I, [_P, C.POINTER(Band), _SZ, _SZ, _I, _P, _I]),
    "sarpro_encode_jpeg": (_I, [_P, C.POINTER(Image), _I, _P, _SZ, C.POINTER(_SZ)]),
    "sarpro_encode_last_jpeg": (_I, [_P, _I, _I, _P, _SZ, C.POINTER(_SZ)]),
    "sarpro_pipeline_batch": (_I, [_P, C.POINTER(Scene), _SZ, _I, _I, _I, _I, _I, _SZ, _I, _I, _I, C.POINTER(Image), C.POINTER(Stats),
                                   C.POINTER(C.c_int), C.POINTER(BatchReport)]),
    "sarpro_plan_from_dn_histogram": (_I, [_P, _I, _I, C.POINTER(Stats), _P]),
    "sarpro_plan_from_present_list": (_I, [_P, _P, C.c_uint32, _I, _I, C.POINTER(Stats), _P]),
    "sarpro_plan_on_device": (_I, [_P, _P, _I, _I, _I, C.POINTER(Stats), _P, _P]),
    "sarpro_plan_kind_from_dn_histogram": (_I, [_P, _I, _I, _I, C.POINTER(Stats), _P, _P]),
    "sarpro_read_row_plan_check": (_I, [_P, _SZ, _SZ, _I, _P]),
    "sarpro_f32_guard_params": (_I, [C.c_double, C.c_double, C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "sarpro_lanczos_row_plan_check": (_I, [_P, _SZ, _SZ, _SZ, _SZ, _P, _P]),
    "sarpro_lanczos_row_check_u16": (_I, [_P, _SZ, _SZ, _P]),
    "sarpro_synrgb_lut_check": (_I, [_I, _P, C.c_uint64, C.POINTER(C.c_int), _P, _P, _P]),
    "sarpro_plan_from_stat_histogram": (_I, [_P, C.c_uint64, C.c_float, C.c_float, C.c_double, C.c_double, _I, _I, C.POINTER(Stats)]),
    "sarpro_narrow_f32_check": (_I, [_P, _SZ, _P, C.POINTER(C.c_int)]),
    "sarpro_f32_edges_check": (_I, [_I, C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_uint32), _P]),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded library. Raises if it has not been built — there is no Python/CPU substitute."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m sarpro_b200.build` "
                "(sarpro_b200 has no CPU or pure-Python fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib

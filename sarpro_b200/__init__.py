"""sarpro_b200 — B200 (sm_100a) implementation of SARPRO's per-pixel raster hot path.

The compute lives in libsarpro_gpu.so (sarpro_b200/csrc, C ABI in include/sarpro_gpu.h); this package
is the ctypes binding plus a host-side mirror of the reference's function names. There is no CPU or
pure-Python path: importing works anywhere, using it needs the built library and a B200.
"""
from . import _ffi
from ._ffi import (ADAPTIVE, BATCH_MULTIBAND, BATCH_SYNRGB, RESAMPLE_AVERAGE, RESAMPLE_LANCZOS, CLAHE, DEFAULT, EQUALIZED, JPEG, OP_DIFF, OP_LOGRATIO, OP_NDIFF, OP_NONE, OP_RATIO,
                   OP_SUM, ROBUST, STANDARD, STRATEGY_NAMES, TAMED, TIFF, U8, U16)
import os as _os

from .api import (Context, ProcessedImage, SarproError, comm_unique_id, plan_from_dn_histogram, plan_kind_from_dn_histogram, plan_from_present_list, present_list_from_histogram,
                  shard_halo_rows,
                  shard_rows)


def _locate_nccl():
    """Point the library at the NCCL torch ships (the library dlopen()s it lazily; no link-time dependency)."""
    if _os.environ.get("SARPRO_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = _os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if _os.path.exists(cand):
                _os.environ["SARPRO_NCCL_LIB"] = cand
    except Exception:
        pass


_locate_nccl()

__all__ = [
    "Context", "ProcessedImage", "SarproError", "plan_from_dn_histogram", "plan_kind_from_dn_histogram", "plan_from_present_list", "present_list_from_histogram", "shard_rows", "shard_halo_rows",
    "comm_unique_id", "_ffi",
    "STANDARD", "ROBUST", "ADAPTIVE", "EQUALIZED", "CLAHE", "TAMED", "DEFAULT", "STRATEGY_NAMES",
    "U8", "U16", "TIFF", "JPEG", "BATCH_MULTIBAND", "BATCH_SYNRGB", "RESAMPLE_AVERAGE", "RESAMPLE_LANCZOS", "OP_NONE", "OP_SUM", "OP_DIFF", "OP_RATIO", "OP_NDIFF", "OP_LOGRATIO",
]

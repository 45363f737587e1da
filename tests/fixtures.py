"""Branch-forcing synthetic rasters shared by the CPU and GPU tests (SURVEY.md §8a 'branch coverage')."""
from __future__ import annotations

import numpy as np

from sarpro_b200.synth import synth_band


def speckle(rows, cols, seed, **kw):
    kw.setdefault("block", 16)  # small rasters still get every terrain class
    return synth_band(rows, cols, seed, **kw)


def low_contrast(rows, cols, seed):
    """dynamic range < 15 dB -> Standard branch 1 (median-based, gamma 1.1)."""
    rng = np.random.default_rng(seed)
    return rng.integers(100, 400, (rows, cols)).astype(np.uint16)  # 12 dB


def homogeneous(rows, cols, seed):
    """range >= 15 dB but IQR < 5 dB -> Standard branch 2."""
    rng = np.random.default_rng(seed)
    dn = np.rint(200 * np.sqrt(rng.gamma(30.0, 1 / 30.0, (rows, cols)))).astype(np.uint16)
    dn[0, :8] = [10, 20, 3000, 4000, 15, 25, 2500, 12]
    return dn


def high_dynamic(rows, cols, seed):
    """range > 40 dB and IQR >= 5 -> Standard branch 3 (gamma 0.9)."""
    dn = synth_band(rows, cols, seed, point_targets=1e-3, block=16)
    dn[1, 50:60] = 1
    return dn


def all_equal(rows, cols, value=77):
    return np.full((rows, cols), value, np.uint16)


def all_zero(rows, cols):
    return np.zeros((rows, cols), np.uint16)


def skewed(rows, cols, seed, positive=True):
    """Two-level raster: 70 % of the pixels near one level, 30 % near another 30 dB away, so that
    (mean - median) / std = +-0.65 -> the skewed Adaptive branches (autoscale.rs:506-512)."""
    rng = np.random.default_rng(seed)
    minority = rng.random((rows, cols)) < 0.3
    lo = rng.integers(9, 12, (rows, cols))
    hi = rng.integers(9990, 10011, (rows, cols))
    dn = np.where(minority, hi, lo) if positive else np.where(minority, lo, hi)
    return dn.astype(np.uint16)


def heavy_tail(rows, cols, seed):
    """|skew| <= 0.5 with (p99-p95)/(p95-p75) > 2 for Adaptive branch 3."""
    rng = np.random.default_rng(seed)
    dn = rng.integers(95, 106, (rows, cols)).astype(np.uint16)
    m = rng.random((rows, cols)) < 0.03
    dn[m] = rng.integers(800, 5000, int(m.sum())).astype(np.uint16)
    m2 = rng.random((rows, cols)) < 0.03
    dn[m2] = rng.integers(2, 12, int(m2.sum())).astype(np.uint16)
    return dn


def narrow_output(rows, cols, seed):
    """U8 result whose max < 255 before scale_u16_to_u8 (so the re-stretch really rescales)."""
    rng = np.random.default_rng(seed)
    dn = rng.integers(1000, 1010, (rows, cols)).astype(np.uint16)
    dn[0, 0] = 1
    dn[0, 1] = 60000
    return dn


CASES = {
    "speckle": lambda r, c: speckle(r, c, 11),
    "speckle_vh": lambda r, c: speckle(r, c, 12, cross_pol=True),
    "low_contrast": lambda r, c: low_contrast(r, c, 13),
    "homogeneous": lambda r, c: homogeneous(r, c, 14),
    "high_dynamic": lambda r, c: high_dynamic(r, c, 11),
    "all_equal": lambda r, c: all_equal(r, c),
    "all_zero": lambda r, c: all_zero(r, c),
    "skew_pos": lambda r, c: skewed(r, c, 16, True),
    "skew_neg": lambda r, c: skewed(r, c, 17, False),
    "heavy_tail": lambda r, c: heavy_tail(r, c, 18),
    "narrow_output": lambda r, c: narrow_output(r, c, 19),
    "no_invalid": lambda r, c: speckle(r, c, 20, black_cols=0) + np.uint16(1),
}

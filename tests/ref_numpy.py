"""Independent numpy re-derivation of the reference's raster path (vectorised, written from the Rust
source separately from oracle/*.cpp) used to cross-check the C++ oracle on small arrays.
Citations are file:line in the reference tree."""
from __future__ import annotations

import math

import numpy as np

STANDARD, ROBUST, ADAPTIVE, EQUALIZED, CLAHE, TAMED, DEFAULT = range(7)


def db_and_mask(v):  # pipeline.rs:19-22
    db = 10.0 * np.log10(np.fmax(v.astype(np.float64), 1e-10))
    return db, db > -50.0


def rs_round(x):  # f64::round: half away from zero
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def stats(db, mask):  # autoscale.rs:35-160
    vals = db[mask]
    n = vals.size
    if n == 0:
        return None
    mn, mx = float(vals.min()), float(vals.max())
    out = {"n": n, "min": mn, "max": mx, "mean": float(vals.mean()), "std": float(vals.std()) if n > 1 else 0.0}
    names = ["median", "p01", "p02", "p05", "p10", "p25", "p75", "p90", "p95", "p98", "p99"]
    ps = [0.5, 0.01, 0.02, 0.05, 0.10, 0.25, 0.75, 0.90, 0.95, 0.98, 0.99]
    if abs(mx - mn) < np.finfo(np.float64).eps:
        for k in names[:6]:
            out[k] = mn
        for k in names[6:]:
            out[k] = mx
        out["median"] = mn
        return out
    span = mx - mn
    inv = 1.0 / span
    t = np.clip((vals - mn) * inv, 0.0, 1.0)
    idx = np.minimum((t * 4096.0).astype(np.int64), 4095)
    hist = np.bincount(idx, minlength=4096).astype(np.int64)
    csum = np.cumsum(hist)
    out["hist"] = hist
    for k, p in zip(names, ps):
        target = min(int(math.floor(p * n)), n - 1)
        b = int(np.searchsorted(csum, target, side="right"))
        before = int(csum[b - 1]) if b > 0 else 0
        frac = (target - before) / float(hist[b]) if hist[b] > 0 else 0.0
        bw = span / 4096.0
        out[k] = mn + b * bw + frac * bw
    return out


def window(st, strategy, kind="autoscale"):  # autoscale.rs:404-428, 491-562, 721-727
    mn, mx = st["min"], st["max"]
    iqr = st["p75"] - st["p25"]
    if kind == "tamed_copol":
        return min(st["p02"], st["p05"]), st["p99"], 1.0
    if kind == "tamed_cross":
        return st["p05"], st["p99"], 1.0
    if strategy == STANDARD:
        dr = mx - mn
        if dr < 15.0:
            r = max(20.0, dr * 0.8)
            lo, hi, g = st["median"] - r / 2.0, st["median"] + r / 2.0, 1.1
        elif iqr < 5.0:
            lo, hi, g = st["p25"] - 2.5 * iqr, st["p75"] + 2.5 * iqr, 1.0
        elif dr > 40.0:
            lo, hi, g = max(st["p02"], mn + 0.02 * dr), min(st["p98"], mx - 0.02 * dr), 0.9
        else:
            lo, hi, g = st["p02"], st["p98"], 1.0
        return max(lo, mn), min(hi, mx), g
    if strategy == ROBUST:
        thr = 2.5 * iqr
        return max(max(st["p25"] - thr, st["p01"]), mn), min(min(st["p75"] + thr, st["p99"]), mx), 1.0
    if strategy == ADAPTIVE:
        skew = (st["mean"] - st["median"]) / max(abs(st["std"]), 1.0)
        tail = (st["p99"] - st["p95"]) / max(st["p95"] - st["p75"], 1.0)
        if abs(skew) > 0.5:
            return (st["p02"], st["p98"], 0.9) if skew > 0 else (st["p05"], st["p95"], 1.1)
        if tail > 2.0:
            return st["p10"], st["p90"], 0.8
        return st["p05"], st["p95"], 1.0
    if strategy in (EQUALIZED, CLAHE):
        return st["p01"], st["p99"], 1.0
    if strategy == TAMED:
        return st["p25"], st["p99"], 1.0
    return st["p05"], st["p95"], 1.0


def quantize(db, mask, lo, hi, g, max_val):  # autoscale.rs:437-446
    rng = max(hi - lo, 1.0)
    clipped = np.fmin(np.fmax(db, lo), hi)
    with np.errstate(invalid="ignore"):
        norm = np.power((clipped - lo) / rng, g)
    q = np.clip(norm * max_val, 0.0, max_val)
    q = np.where(np.isnan(q), 0.0, q)
    return np.where(mask, np.floor(q), 0).astype(np.uint16)


def scale_u16_to_u8(d):  # autoscale.rs:348-364
    if d.size == 0:
        return d.astype(np.uint8)
    mn, mx = np.float32(d.min()), np.float32(d.max())
    scale = np.float32(255.0) / (mx - mn) if mx > mn else np.float32(1.0)
    val = (d.astype(np.float32) - mn) * scale
    val = np.sign(val) * np.floor(np.abs(val) + np.float32(0.5))
    return np.clip(val, 0, 255).astype(np.uint8)


def clahe(norm, mask, tiles=8, clip_limit=2.0, bins=256):  # autoscale.rs:220-345
    rows, cols = norm.shape
    th, tw = -(-rows // tiles), -(-cols // tiles)
    binimg = rs_round(np.clip(norm, 0.0, 1.0) * (bins - 1.0)).astype(np.int64)
    cdfs = np.zeros((tiles, tiles, bins))
    for ty in range(tiles):
        for tx in range(tiles):
            r0, r1, c0, c1 = ty * th, min((ty + 1) * th, rows), tx * tw, min((tx + 1) * tw, cols)
            if r1 <= r0 or c1 <= c0:
                continue
            m = mask[r0:r1, c0:c1]
            h = np.bincount(binimg[r0:r1, c0:c1][m], minlength=bins).astype(np.float64)
            thr = max(clip_limit * ((r1 - r0) * (c1 - c0) / float(bins)), 1.0)
            over = h > thr
            excess = float(np.sum(h[over] - thr))
            h[over] = math.floor(thr)
            add = math.floor(excess / bins)
            rem = int(math.floor(excess - add * bins + 0.5))
            h = np.floor(h + add)
            h += rem // bins
            h[: rem % bins] += 1
            total = max(h.sum(), 1.0)
            cdfs[ty, tx] = np.clip(np.cumsum(h) / total, 0.0, 1.0)
    rr = np.arange(rows) / float(th) - 0.5
    cc = np.arange(cols) / float(tw) - 0.5
    ty = np.maximum(np.floor(rr), 0).astype(np.int64)
    tx = np.maximum(np.floor(cc), 0).astype(np.int64)
    dy = (rr - ty)[:, None]
    dx = (cc - tx)[None, :]
    ty0, ty1 = np.clip(ty, 0, tiles - 1), np.clip(ty + 1, 0, tiles - 1)
    tx0, tx1 = np.clip(tx, 0, tiles - 1), np.clip(tx + 1, 0, tiles - 1)
    c00 = cdfs[ty0[:, None], tx0[None, :], binimg]
    c01 = cdfs[ty0[:, None], tx1[None, :], binimg]
    c10 = cdfs[ty1[:, None], tx0[None, :], binimg]
    c11 = cdfs[ty1[:, None], tx1[None, :], binimg]
    top = c00 * (1.0 - dx) + c01 * dx
    bottom = c10 * (1.0 - dx) + c11 * dx
    out = top * (1.0 - dy) + bottom * dy
    return np.where(mask, out, 0.0), cdfs.reshape(tiles * tiles, bins)


def autoscale(v, bit_depth_u8, strategy, kind="autoscale"):
    """process_scalar_data_pipeline (pipeline.rs:42-66) -> final u8 or u16 plane."""
    db, mask = db_and_mask(v)
    st = stats(db, mask)
    tamed_rgb = kind != "autoscale"
    if st is None:
        return np.zeros(v.shape, np.uint8 if (bit_depth_u8 or tamed_rgb) else np.uint16), None
    lo, hi, g = window(st, strategy, kind)
    max_val = 255.0 if (bit_depth_u8 or tamed_rgb) else 65535.0
    if tamed_rgb:
        return quantize(db, mask, lo, hi, 1.0, 255.0).astype(np.uint8), st
    if strategy == CLAHE:
        rng = max(hi - lo, 1.0)
        norm = np.where(mask, (np.fmin(np.fmax(db, lo), hi) - lo) / rng, 0.0)
        eq, _ = clahe(norm, mask)
        q = np.where(mask, np.floor(np.clip(eq, 0.0, 1.0) * max_val), 0).astype(np.uint16)
    else:
        q = quantize(db, mask, lo, hi, g, max_val)
    return (scale_u16_to_u8(q) if bit_depth_u8 else q), st


# ---- Lanczos3 (fast_image_resize 5.x restated independently with numpy) ---------------------
def _lanczos3(x):
    x = np.asarray(x, np.float64)
    out = np.sinc(x) * np.sinc(x / 3.0)  # np.sinc(x) = sin(pi x)/(pi x)
    return np.where((x >= -3.0) & (x < 3.0), out, 0.0)


def lanczos_axis(in_size, out_size, wide):
    scale = in_size / out_size
    fs = max(scale, 1.0)
    radius = 3.0 * fs
    rows = []
    maxw = 0.0
    for ox in range(out_size):
        center = (ox + 0.5) * scale
        x0 = int(max(math.floor(center - radius), 0))
        x1 = int(min(math.ceil(center + radius), in_size))
        xs = np.arange(x0, x1)
        w = _lanczos3((xs - (center - 0.5)) / fs)
        while w.size and w[0] == 0.0:
            w, xs = w[1:], xs[1:]
        full = w.copy()
        while w.size and w[-1] == 0.0:
            w, xs = w[:-1], xs[:-1]
        s = full.sum()
        if s != 0.0:
            w = w / s
        rows.append((int(xs[0]) if xs.size else x0, w))
        maxw = max(maxw, float(w.max()) if w.size else 0.0)
    limit, top = (22, 1 << 15) if not wide else (46, 1 << 31)
    p = 0
    for cur in range(limit):
        p = cur
        if math.floor(maxw * (1 << (p + 1)) + 0.5) >= top:
            break
    coefs = [(s, np.array([int(math.floor(abs(c) * (1 << p) + 0.5)) * (1 if c >= 0 else -1) for c in w], dtype=np.int64)) for s, w in rows]
    return p, coefs


def _convolve_rows(img, out_size, wide):
    in_size = img.shape[1]
    p, coefs = lanczos_axis(in_size, out_size, wide)
    out = np.empty((img.shape[0], out_size), np.int64)
    src = img.astype(np.int64)
    half = (1 << (p - 1)) if p > 0 else 0
    for ox, (s, k) in enumerate(coefs):
        out[:, ox] = (src[:, s:s + k.size] @ k + half) >> p
    hi = 65535 if wide else 255
    return np.clip(out, 0, hi).astype(img.dtype)


def resize_lanczos3(img, tcols, trows):
    wide = img.dtype == np.uint16
    tmp = _convolve_rows(img, tcols, wide)                 # horizontal first
    return _convolve_rows(tmp.T.copy(), trows, wide).T.copy()  # then vertical


def resize_dims(cols, rows, target):  # resize.rs:6-30
    if target > max(cols, rows):
        return cols, rows
    short = int(math.floor(min(cols, rows) * (target / max(cols, rows)) + 0.5))
    return (target, short) if cols > rows else (short, target)


def pad_square(img):  # padding.rs:5-49
    rows, cols = img.shape
    m = max(rows, cols)
    out = np.zeros((m, m), img.dtype)
    out[(m - rows) // 2:(m - rows) // 2 + rows, (m - cols) // 2:(m - cols) // 2 + cols] = img
    return out


# ---- synthetic RGB ---------------------------------------------------------------------------------
def _round_half_away32(x):
    x = x.astype(np.float32)
    return np.sign(x) * np.floor(np.abs(x) + np.float32(0.5))


def synrgb_default(b1, b2):  # synthetic_rgb.rs:10-67
    v = np.arange(256, dtype=np.float32) / np.float32(255.0)
    lr = np.clip(_round_half_away32(np.power(v, np.float32(0.7)) * np.float32(255.0)), 0, 255).astype(np.uint8)
    lg = np.clip(_round_half_away32(np.power(v, np.float32(0.9)) * np.float32(255.0)), 0, 255).astype(np.uint8)
    r = lr[b1].astype(np.float32)
    g = lg[b2].astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = r / g
        blue = np.power(ratio, np.float32(0.1)) * np.float32(255.0) * np.float32(0.24)
    blue = _round_half_away32(np.clip(blue, 0, 255))
    blue = np.where(np.isnan(blue), 0, blue).astype(np.uint8)
    blue = np.where(b2 == 0, 0, blue).astype(np.uint8)
    return np.stack([lr[b1], lg[b2], blue], axis=-1)


def synrgb_suppressed(b1, b2):  # synthetic_rgb.rs:88-178
    hist = np.bincount(b1.ravel(), minlength=256) + np.bincount(b2.ravel(), minlength=256)
    total = (b1.size + b2.size) & 0xFFFFFFFF
    target = int(math.floor(total * 0.05 + 0.5))
    csum = np.cumsum(hist)
    hit = np.nonzero(csum >= target)[0]
    floor_value = int(hit[0]) if hit.size else 0
    fwc = min(floor_value + 3, 40)
    f = np.float32(fwc)
    denom = max(np.float32(255.0) - f, np.float32(1.0))
    v = np.arange(256, dtype=np.float32)
    shifted = (v - f) / denom
    with np.errstate(invalid="ignore"):
        lr = np.clip(_round_half_away32(np.power(shifted, np.float32(1.15)) * np.float32(255.0)), 0, 255)
        lg = np.clip(_round_half_away32(np.power(shifted, np.float32(1.10)) * np.float32(255.0)), 0, 255)
    lr = np.where(np.arange(256) <= fwc, 0, np.nan_to_num(lr)).astype(np.uint8)
    lg = np.where(np.arange(256) <= fwc, 0, np.nan_to_num(lg)).astype(np.uint8)
    r = lr[b1].astype(np.float32)
    g = lg[b2].astype(np.float32)
    ratio = (r + np.float32(8.0)) / (g + np.float32(8.0))
    blue = _round_half_away32(np.clip(np.power(ratio, np.float32(0.1)) * np.float32(255.0) * np.float32(0.18), 0, 255)).astype(np.uint8)
    rgb = np.stack([lr[b1], lg[b2], blue], axis=-1)
    water = (b1 <= fwc) & (b2 <= fwc)
    rgb[water] = 0
    return rgb

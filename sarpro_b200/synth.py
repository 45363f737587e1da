"""Synthetic Sentinel-1-GRD-like u16 DN rasters: gamma speckle (ENL 4.4) over a
64x64-block class map, black-fill left margin (invalid pixels), optional bright point targets.
numpy version for tests; torch version (same recipe, device RNG) for full-size bench inputs."""
from __future__ import annotations

import numpy as np

SEED_VV = 20251017
SEED_VH = 20251018
ENL = 4.4


def synth_band(rows: int, cols: int, seed: int, cross_pol: bool = False, black_cols: int = 40,
               point_targets: float = 1e-5, block: int = 64) -> np.ndarray:
    rng = np.random.default_rng(seed)
    br, bc = -(-rows // block), -(-cols // block)
    u = rng.random((br, bc))
    mean = np.where(u < 0.3, 30.0, np.where(u < 0.9, 150.0, 600.0))
    if cross_pol:
        mean = mean * 0.35
    mean = np.repeat(np.repeat(mean, block, axis=0), block, axis=1)[:rows, :cols]
    inten = mean * mean * rng.gamma(ENL, 1.0 / ENL, size=(rows, cols))
    dn = np.clip(np.rint(np.sqrt(inten)), 0, 65535).astype(np.uint16)
    if point_targets > 0:
        n = int(rows * cols * point_targets)
        if n:
            rr = rng.integers(0, rows, n)
            cc = rng.integers(0, cols, n)
            dn[rr, cc] = rng.integers(20000, 60001, n).astype(np.uint16)
    if black_cols > 0:
        dn[:, : min(black_cols, cols)] = 0
    return dn


def synth_pair(rows: int, cols: int, scene: int = 0, **kw):
    return (synth_band(rows, cols, SEED_VV + 2 * scene, False, **kw),
            synth_band(rows, cols, SEED_VH + 2 * scene, True, **kw))


def synth_band_torch(rows: int, cols: int, seed: int, device, cross_pol: bool = False, black_cols: int = 40,
                     point_targets: float = 1e-5, block: int = 64, chunk_rows: int = 2000):
    """Same recipe on the GPU (device RNG, so values differ from the numpy generator)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((rows, cols), dtype=torch.int16, device=device)
    br, bc = -(-rows // block), -(-cols // block)
    u = torch.rand((br, bc), generator=g, device=device)
    mean_b = torch.where(u < 0.3, 30.0, torch.where(u < 0.9, 150.0, 600.0))
    if cross_pol:
        mean_b = mean_b * 0.35
    conc = torch.tensor(ENL, device=device)
    for r0 in range(0, rows, chunk_rows):
        r1 = min(rows, r0 + chunk_rows)
        idx_r = torch.arange(r0, r1, device=device) // block
        idx_c = torch.arange(cols, device=device) // block
        mean = mean_b[idx_r][:, idx_c]
        gam = torch._standard_gamma(conc.expand(r1 - r0, cols), generator=g) / ENL
        dn = torch.clamp(torch.round(torch.sqrt(mean * mean * gam)), 0, 65535)
        if point_targets > 0:
            m = torch.rand((r1 - r0, cols), generator=g, device=device) < point_targets
            bright = torch.randint(20000, 60001, (r1 - r0, cols), generator=g, device=device).to(dn.dtype)
            dn = torch.where(m, bright, dn)
        out[r0:r1] = dn.to(torch.int32).to(torch.int16)  # u16 bit pattern in an int16 tensor
    if black_cols > 0:
        out[:, : min(black_cols, cols)] = 0
    return out

"""Device time of the shapes the tensor-core pass B (k_hmma) does not take, through the generic exact kernel (k_hresize), next
to the k_hmma shape: a source width that is not a multiple of 8, a U16 output, a scale factor below 6.5, and a DN table of more
than 2000 hot values. Inputs resident in HBM; median of 5 calls after 2 warm-ups; CUDA events of the library (timing().total_ms)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import json
import numpy as np
import torch
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch

dev = torch.device("cuda:0")
ROWS = int(os.environ.get("ROWS", 16000))


def med(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        fn()
        ts.append(ctx.timing().total_ms)
    return float(np.median(ts))


out = {}
with S.Context(0) as ctx:
    for cols in (25000, 25004):
        vv = synth_band_torch(ROWS, cols, SEED_VV, dev)
        vh = synth_band_torch(ROWS, cols, SEED_VH, dev, cross_pol=True)
        torch.cuda.synchronize()
        for name, strat in (("clahe", S.CLAHE), ("robust", S.ROBUST)):
            out[f"synrgb_{name}_2048_cols{cols}"] = med(lambda: ctx.process_synrgb_jpeg(vv, vh, strat, 2048, True, out=S.Context.KEEP))
        if cols == 25000:
            out["multiband_u16_clahe_2048"] = med(lambda: ctx.process_multiband_tiff(vv, vh, S.U16, S.CLAHE, 2048, True))
            out["multiband_u16_robust_2048"] = med(lambda: ctx.process_multiband_tiff(vv, vh, S.U16, S.ROBUST, 2048, True))
            out["synrgb_clahe_6250_scale4"] = med(lambda: ctx.process_synrgb_jpeg(vv, vh, S.CLAHE, 6250, True, out=S.Context.KEEP))
            # more than 2000 hot DNs: uniform DNs over [1, 40000]
            g = torch.Generator(device=dev); g.manual_seed(5)
            wide = torch.randint(1, 40000, (ROWS, cols), generator=g, device=dev, dtype=torch.int32).to(torch.int16)
            torch.cuda.synchronize()
            out["synrgb_clahe_2048_wide_table"] = med(lambda: ctx.process_synrgb_jpeg(wide, vh, S.CLAHE, 2048, True, out=S.Context.KEEP))
            del wide
        del vv, vh
        torch.cuda.empty_cache()
print(json.dumps({"rows": ROWS, "ms": {k: round(v, 3) for k, v in out.items()}}))

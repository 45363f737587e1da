"""Exploration timing (not the contract bench): per-variant device times of the pipelines."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch

rows = int(os.environ.get("ROWS", 16000))
cols = int(os.environ.get("COLS", 25000))
dev = torch.device("cuda:0")
t0 = time.time()
vv = synth_band_torch(rows, cols, SEED_VV, dev)
vh = synth_band_torch(rows, cols, SEED_VH, dev, cross_pol=True)
torch.cuda.synchronize()
print(f"synth {rows}x{cols} x2 in {time.time()-t0:.1f}s; VV max {int(vv.to(torch.int32).bitwise_and(0xffff).max())}", flush=True)

for variant in os.environ.get("VARIANTS", "0,1,2,3,4,5,6").split(","):
    os.environ["SARPRO_HIST_VARIANT"] = variant
    ctx = S.Context(0)
    for strat, name in ((S.ROBUST, "robust"), (S.CLAHE, "clahe")):
        out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
        times = []
        for it in range(int(os.environ.get("ITERS", 4))):
            ctx.process_synrgb_jpeg(vv, vh, strat, 2048, True, out=out)
            t = ctx.timing()
            times.append(t.total_ms)
        print(f"hist_variant={variant} {name}: ms={['%.3f' % x for x in times]} launches={t.kernel_launches} syncs={t.host_syncs} "
              f"Mpx/s={rows*cols/min(times)/1e3:.0f}", flush=True)
    ctx.close()

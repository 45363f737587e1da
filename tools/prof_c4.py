"""A few C4 calls (sarpro_pipeline_polops: log-ratio + n-diff -> Equalized -> two u16 bands) for ncu captures of the general-path kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
dev = torch.device("cuda:0")
rows, cols = int(os.environ.get("ROWS", 16000)), int(os.environ.get("COLS", 25000))
vv = synth_band_torch(rows, cols, SEED_VV, dev); vh = synth_band_torch(rows, cols, SEED_VH, dev, cross_pol=True)
torch.cuda.synchronize()
outs = [torch.empty((rows, cols), dtype=torch.uint16, device=dev) for _ in range(2)]
with S.Context(0) as ctx:
    for _ in range(int(os.environ.get("ITERS", 2))):
        ctx.process_polops(vv, vh, (S.OP_LOGRATIO, S.OP_NDIFF), S.U16, S.EQUALIZED, outs=outs)
        t = ctx.timing()
        print("total_ms", t.total_ms, "stage_ms", list(t.stage_ms), flush=True)

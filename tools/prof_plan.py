"""One C3 call (for ncu captures of the small kernels: planner, CLAHE statistics)."""
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
out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
with S.Context(0) as ctx:
    for _ in range(int(os.environ.get("ITERS", 2))):
        ctx.process_synrgb_jpeg(vv, vh, S.CLAHE, 2048, True, out=out)

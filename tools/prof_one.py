"""Runs the synRGB pipeline a few times for ncu captures (strategy from argv)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
rows = int(os.environ.get("ROWS", 16000)); cols = int(os.environ.get("COLS", 25000))
strat = S.STRATEGY_NAMES.index(sys.argv[1] if len(sys.argv) > 1 else "clahe")
dev = torch.device("cuda:0")
vv = synth_band_torch(rows, cols, SEED_VV, dev); vh = synth_band_torch(rows, cols, SEED_VH, dev, cross_pol=True)
torch.cuda.synchronize()
ctx = S.Context(0)
out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
for _ in range(int(os.environ.get("ITERS", 2))):
    ctx.process_synrgb_jpeg(vv, vh, strat, 2048, True, out=out)
print("ok", ctx.timing().total_ms)

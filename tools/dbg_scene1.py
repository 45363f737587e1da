import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
dev = torch.device("cuda:0")
k = int(os.environ.get("SCENE", 1))
vh = synth_band_torch(16000, 25000, SEED_VH + 2 * k, dev, cross_pol=True)
torch.cuda.synchronize()
ctx = S.Context(0)
for it in range(2):
    img = ctx.process_single(vh, S.TIFF, S.U8, S.CLAHE, 2048, False)
    t = ctx.timing()
    print("apply ms", t.stage_ms[2], "total", t.total_ms, flush=True)
print("stages", [round(x, 3) for x in t.stage_ms], "launches", list(t.stage_launches), flush=True)
import numpy as np
g = img.gray
print("out min/max", int(g.min()), int(g.max()), g.shape, flush=True)

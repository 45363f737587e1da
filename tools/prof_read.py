"""One downsample-on-read call (Average, 25000 x 16000 u16 -> 2048 x 1311 f32) for ncu captures of k_read_average."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VV, synth_band_torch
dev = torch.device("cuda:0")
rows, cols = int(os.environ.get("ROWS", 16000)), int(os.environ.get("COLS", 25000))
vv = synth_band_torch(rows, cols, SEED_VV, dev)
torch.cuda.synchronize()
oc, orr, alg = S.Context.read_dims_for_target(cols, rows, 2048)
out = torch.empty((orr, oc), dtype=torch.float32, device=dev)
with S.Context(0) as ctx:
    for _ in range(int(os.environ.get("ITERS", 2))):
        ctx.read_band_resampled(vv, oc, orr, alg, out=out)

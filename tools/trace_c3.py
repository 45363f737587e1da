"""Device timeline of one C3 / C2 call with every launch timed (SARPRO_TRACE=1 from context creation on). STRATEGY=clahe|robust."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SARPRO_TRACE"] = "1"
import torch
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
dev = torch.device("cuda:0")
rows, cols = int(os.environ.get("ROWS", 16000)), int(os.environ.get("COLS", 25000))
vv = synth_band_torch(rows, cols, SEED_VV, dev); vh = synth_band_torch(rows, cols, SEED_VH, dev, cross_pol=True)
torch.cuda.synchronize()
strategy = S.STRATEGY_NAMES.index(os.environ.get("STRATEGY", "clahe"))
out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
with S.Context(0) as ctx:
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)  # warm-up calls: no trace output
    for _ in range(5):
        ctx.process_synrgb_jpeg(vv, vh, strategy, 2048, True, out=out)
    os.dup2(saved, 2)
    ctx.process_synrgb_jpeg(vv, vh, strategy, 2048, True, out=out)
    t = ctx.timing()
    print("stage ms:", {S._ffi.STAGE_NAMES[i]: round(t.stage_ms[i], 4) for i in range(8) if t.stage_launches[i]}, "total", round(t.total_ms, 4),
          "launches", t.kernel_launches, "host syncs", t.host_syncs, flush=True)

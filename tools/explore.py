"""GPU exploration (not the contract bench): full-size C3/C2 pipelines under kernel variants selected by env,
bitwise comparison of the outputs between variants, per-stage device times."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import sarpro_b200 as S
from sarpro_b200._ffi import STAGE_NAMES
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch

rows = int(os.environ.get("ROWS", 16000))
cols = int(os.environ.get("COLS", 25000))
iters = int(os.environ.get("ITERS", 4))
dev = torch.device("cuda:0")
t0 = time.time()
vv = synth_band_torch(rows, cols, SEED_VV, dev)
vh = synth_band_torch(rows, cols, SEED_VH, dev, cross_pol=True)
torch.cuda.synchronize()
print(f"synth {rows}x{cols} x2 in {time.time()-t0:.1f}s", flush=True)

# (label, env) pairs
configs = []
for spec in os.environ.get("CONFIGS", "old:SARPRO_HPIPE=0,SARPRO_HIST_VARIANT=0;new:SARPRO_HPIPE=1,SARPRO_HIST_VARIANT=0").split(";"):
    label, envs = spec.split(":")
    configs.append((label, dict(kv.split("=") for kv in envs.split(",") if kv)))

ref = {}
for label, env in configs:
    for k, v in env.items():
        os.environ[k] = v
    ctx = S.Context(0)
    for strat, name in ((S.CLAHE, "clahe"), (S.ROBUST, "robust")):
        out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
        best = None
        for it in range(iters):
            ctx.process_synrgb_jpeg(vv, vh, strat, 2048, True, out=out)
            t = ctx.timing()
            if best is None or t.total_ms < best.total_ms:
                best = t
        stages = " ".join(f"{STAGE_NAMES[i]}={best.stage_ms[i]:.3f}" for i in range(8) if best.stage_ms[i] > 0)
        key = name
        same = ""
        if key in ref:
            nd = int((ref[key] != out).sum())
            same = f" diff_vs_first={nd}"
        else:
            ref[key] = out.clone()
        print(f"[{label}] {name}: total={best.total_ms:.3f} ms {stages} launches={best.kernel_launches} syncs={best.host_syncs} "
              f"Mpx/s={rows*cols/best.total_ms/1e3:.0f}{same}", flush=True)
    ctx.close()

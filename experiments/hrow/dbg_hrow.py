"""Debug: k_hrow against the previous kernel (SARPRO_HROW=0) on one raster; prints where they differ and the device times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sarpro_b200 as S
from sarpro_b200.synth import synth_pair
rows, cols, target = (int(x) for x in os.environ.get("SHAPE", "1211x2048x640").split("x"))
strategy = S.STRATEGY_NAMES.index(os.environ.get("STRATEGY", "clahe"))
vv, vh = synth_pair(rows, cols, point_targets=1e-4)
if os.environ.get("ZERO"):
    vv[300:500, -120:] = 0
outs = {}
for env in ("0", "1"):
    os.environ["SARPRO_HROW"] = env
    with S.Context(0) as c:
        t0 = time.time()
        img = c.process_single(vv, S.TIFF, S.U8, strategy, target, True)
        img = c.process_single(vv, S.TIFF, S.U8, strategy, target, True)
        t = c.timing()
        print("HROW", env, "wall", round(time.time() - t0, 3), "device ms", round(t.total_ms, 3), "apply", round(t.stage_ms[2], 3), flush=True)
        outs[env] = img.gray.copy()
d = outs["0"] != outs["1"]
print("diff", int(d.sum()), "of", d.size)
if d.any():
    ys, xs = np.nonzero(d)
    print("rows", ys.min(), ys.max(), "cols", xs.min(), xs.max(), "pad_top", img.pad_top, "pad_left", img.pad_left)
    print("distinct rows", np.unique(ys)[:40], "n", len(np.unique(ys)))
    print("distinct cols", np.unique(xs)[:40], "n", len(np.unique(xs)))
    dv = outs["1"].astype(int) - outs["0"].astype(int)
    print("delta hist", np.unique(dv[d], return_counts=True))

"""Compares the tensor-core pass B (kernels_hmma.cu) with the second-generation / exact kernels on one band and
reports where they differ; then times the C3 scene with each (exploration, not the contract bench)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sarpro_b200 as S
from sarpro_b200.synth import synth_band

def run(dn, strat, target, env):
    for k in ("SARPRO_HMMA", "SARPRO_HPIPE", "SARPRO_FORCE_EXACT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with S.Context(0) as c:
        img = c.process_single(dn, S.TIFF, S.U8, strat, target, False)
        t = c.timing()
        return img.gray.copy(), t

cases = [(900, 1400, 512), (1031, 2048, 300), (2500, 9000, 700), (3001, 5000, 1024), (600, 4096, 2048), (517, 25000, 2048)]
if len(sys.argv) > 3:
    cases = [tuple(int(x) for x in sys.argv[1:4])]
bad = 0
for rows, cols, target in cases:
    dn = synth_band(rows, cols, 11, block=16)
    dn[rows // 3: rows // 2, -77:] = 0
    dn[5:9, 100:300] = 60000
    for strat, name in ((S.CLAHE, "clahe"), (S.ROBUST, "robust")):
        ref, _ = run(dn, strat, target, {"SARPRO_HMMA": "0"})
        got, t = run(dn, strat, target, {})
        d = ref != got
        print(f"{rows}x{cols}->{target} {name}: diff {int(d.sum())} of {d.size}; launches {t.kernel_launches}", flush=True)
        if d.any():
            bad += 1
            ys, xs = np.nonzero(d)
            print("  bbox rows", ys.min(), ys.max(), "cols", xs.min(), xs.max())
            print("  col histogram (per 8 cols):", np.bincount(xs // 8, minlength=(got.shape[1] + 7) // 8)[:64])
            for y, x in list(zip(ys, xs))[:8]:
                print("  ", y, x, ref[y, x], got[y, x])
print("MISMATCHING CASES:", bad)

if os.environ.get("TIME", "1") == "1":
    import torch
    from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
    dev = torch.device("cuda:0")
    vv = synth_band_torch(16000, 25000, SEED_VV, dev); vh = synth_band_torch(16000, 25000, SEED_VH, dev, cross_pol=True)
    out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
    res = {}
    for env in ({"SARPRO_HMMA": "0"}, {}):
        for k in ("SARPRO_HMMA",):
            os.environ.pop(k, None)
        os.environ.update(env)
        torch.cuda.synchronize()
        ctx = S.Context(0)
        for strat, name in ((S.CLAHE, "clahe"), (S.ROBUST, "robust")):
            ts = []
            for it in range(5):
                ctx.process_synrgb_jpeg(vv, vh, strat, 2048, True, out=out)
                t = ctx.timing()
                ts.append((t.total_ms, t.stage_ms[2]))
            res[(name, str(env))] = out.cpu().numpy().copy()
            print(f"env={env} {name}: total/apply ms = {[('%.3f' % a, '%.3f' % b) for a, b in ts]}", flush=True)
        ctx.close()
    for name in ("clahe", "robust"):
        a, b = res[(name, str({"SARPRO_HMMA": "0"}))], res[(name, str({}))]
        print(f"C3-size {name}: rgb diff {int((a != b).sum())}")

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sarpro_b200 as S
from sarpro_b200.synth import synth_band
rows, cols, target = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (900, 1400, 512))]
dn = synth_band(rows, cols, 11, block=16)
outs = {}
for hp in ("0", "1"):
    os.environ["SARPRO_HPIPE"] = hp
    with S.Context(0) as c:
        img = c.process_single(dn, S.TIFF, S.U8, S.CLAHE, target, False)
        outs[hp] = img.gray.copy()
        t = c.timing()
        print("hpipe", hp, "launches", t.kernel_launches, "shape", img.gray.shape, flush=True)
a, b = outs["0"].astype(int), outs["1"].astype(int)
d = a != b
print("diff count", int(d.sum()), "of", d.size)
if d.any():
    ys, xs = np.nonzero(d)
    print("bbox rows", ys.min(), ys.max(), "cols", xs.min(), xs.max())
    print("col histogram (per 32 cols):", np.bincount(xs // 32, minlength=(b.shape[1] + 31) // 32))
    print("row histogram (per 32 rows):", np.bincount(ys // 32, minlength=(b.shape[0] + 31) // 32))
    for y, x in list(zip(ys, xs))[:12]:
        print(y, x, a[y, x], b[y, x])
    print("delta stats", np.bincount((b - a)[d] + 255)[200:311])

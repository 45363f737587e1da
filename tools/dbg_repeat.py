"""Poisons freed device memory, then checks call 1 / call 3 of the same scene on one context against the new path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
dev = torch.device("cuda:0")
vv = synth_band_torch(16000, 25000, SEED_VV, dev); vh = synth_band_torch(16000, 25000, SEED_VH, dev, cross_pol=True)
# poison: cudaMalloc'ed (not torch-cached) buffers filled with garbage, then freed
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
ptrs = []
for _ in range(40):
    p = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(64 << 20)) == 0
    rt.cudaMemset(p, 0xA7, ctypes.c_size_t(64 << 20))
    ptrs.append(p)
for p in ptrs:
    rt.cudaFree(p)
def run(env):
    for k in ("SARPRO_HMMA", "SARPRO_TWO_STREAM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    res = []
    with S.Context(0) as c:
        out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
        c.process_synrgb_jpeg(vv, vh, S.CLAHE, 2048, True, out=out); res.append(out.cpu().numpy().copy())
        c.process_synrgb_jpeg(vh, vv, S.CLAHE, 2048, True, out=out)
        c.process_synrgb_jpeg(vv, vh, S.CLAHE, 2048, True, out=out); res.append(out.cpu().numpy().copy())
    return res
new = run({})
old = run({"SARPRO_HMMA": "0", "SARPRO_TWO_STREAM": "0"})
old2 = run({"SARPRO_HMMA": "0"})
for name, r in (("new", new), ("old", old), ("old+2stream", old2)):
    for i, a in enumerate(r):
        d = a != new[0]
        print(name, "call", 1 + 2 * i, "diff vs new call 1:", int(d.sum()), "channels", [int(d[..., ch].sum()) for ch in range(3)],
              ("rows %d-%d cols %d-%d" % (np.nonzero(d.any(axis=(1, 2)))[0].min(), np.nonzero(d.any(axis=(1, 2)))[0].max(),
                                          np.nonzero(d.any(axis=(0, 2)))[0].min(), np.nonzero(d.any(axis=(0, 2)))[0].max())) if d.any() else "")

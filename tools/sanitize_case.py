"""One small CLAHE + one LUT scene through the tensor-core pass B, the general f32 kernels (two polarization operations), the
downsample-on-read kernels and the batch entry (for compute-sanitizer runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sarpro_b200 as S
from sarpro_b200.synth import synth_pair
vv, vh = synth_pair(700, 2048, point_targets=1e-4)
vv[100:200, -90:] = 0
with S.Context(0) as c:
    for strat in (S.CLAHE, S.ROBUST):
        img = c.process_synrgb_jpeg(vv, vh, strat, 256, True)
        print(strat, img.rgb.shape, int(img.rgb.sum()), flush=True)
    planes, _ = c.process_polops(vv, vh, (S.OP_LOGRATIO, S.OP_NDIFF), S.U16, S.EQUALIZED)
    print("polops", int(planes[0].sum()), int(planes[1].sum()), flush=True)
    for alg, shape in ((S.RESAMPLE_AVERAGE, (256, 88)), (S.RESAMPLE_LANCZOS, (1024, 350))):
        r = c.read_band_resampled(vv, shape[0], shape[1], alg)
        print("read", alg, float(r.sum()), flush=True)
    res, st, rep, _ = c.process_batch([(vv, vh), None, (vh, vv)], S.BATCH_MULTIBAND, S.U8, S.CLAHE, 256, True)
    print("batch", rep.processed, rep.skipped, rep.errors, flush=True)

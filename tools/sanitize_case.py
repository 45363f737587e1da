"""One small CLAHE + one LUT scene through the tensor-core pass B (for compute-sanitizer runs)."""
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

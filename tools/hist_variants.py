"""Pass-A stage time of the C3 scene per table shape (SARPRO_HIST_VARIANT)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
dev = torch.device("cuda:0")
vv = synth_band_torch(16000, 25000, SEED_VV, dev); vh = synth_band_torch(16000, 25000, SEED_VH, dev, cross_pol=True)
torch.cuda.synchronize()
out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
for variant in ("", "20", "21", "22"):
    os.environ.pop("SARPRO_HIST_VARIANT", None)
    if variant:
        os.environ["SARPRO_HIST_VARIANT"] = variant
    os.environ["SARPRO_STAGE_TIMING"] = "all"
    with S.Context(0) as c:
        for band, name in ((vv, "VV"), (vh, "VH")):
            ts = []
            for _ in range(4):
                c.process_single(band, S.TIFF, S.U8, S.CLAHE, 2048, False)
                ts.append(c.timing().stage_ms[0])
            print(f"variant {variant or 'auto'} {name}: hist ms {min(ts):.4f}", flush=True)

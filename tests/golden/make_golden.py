"""Regenerates tests/golden/golden_small.npz from the CPU oracle (the reference itself cannot be run in
this image: no Rust toolchain, no GDAL). Run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402
from sarpro_b200.synth import synth_pair  # noqa: E402

vv, vh = synth_pair(120, 168, scene=3, point_targets=1e-3, block=16)
out = {"vv": vv, "vh": vh}
for s in range(7):
    for bd, key in ((O.U8, "u8"), (O.U16, "u16")):
        po = O.process_scalar_data_pipeline(vv.astype(np.float32), bd, s, want_db=False)
        out[f"autoscale_{s}_{key}"] = po.u8 if bd == O.U8 else po.u16
    rgb, _ = O.pipeline_synrgb_jpeg(vv.astype(np.float32), vh.astype(np.float32), s, 96, True)
    out[f"synrgb_{s}"] = rgb
rng = np.random.default_rng(99)
out["img8"] = rng.integers(0, 256, (100, 141)).astype(np.uint8)
out["img16"] = rng.integers(0, 65536, (141, 100)).astype(np.uint16)
out["resize8"] = O.resize_u8_image(out["img8"], 57, 41)
out["resize16"] = O.resize_u16_image(out["img16"], 41, 57)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"), **out)
print("wrote golden_small.npz", {k: v.shape for k, v in out.items() if k.startswith("synrgb")})

# A scene wide enough for the tensor-core pass B (kernels_hmma.cu: source width a multiple of 8, scale factor 8): only the
# outputs and a checksum of the (seeded, regenerated) inputs are stored.
import hashlib  # noqa: E402
wvv, wvh = synth_pair(640, 2048, scene=5, point_targets=1e-4)
wide = {"input_sha256": np.frombuffer(hashlib.sha256(wvv.tobytes() + wvh.tobytes()).digest(), np.uint8)}
for s, name in ((O.CLAHE, "clahe"), (O.ROBUST, "robust"), (O.TAMED, "tamed")):
    rgb, _ = O.pipeline_synrgb_jpeg(wvv.astype(np.float32), wvh.astype(np.float32), s, 256, True)
    wide[f"synrgb_{name}"] = rgb
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_wide.npz"), **wide)
print("wrote golden_wide.npz", {k: v.shape for k, v in wide.items()})

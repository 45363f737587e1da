"""torchrun worker: row-band-sharded synRGB of ONE scene across the ranks vs the single-GPU result (and the
oracle on rank 0). Prints one line per rank; exit code 1 on mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import sarpro_b200 as S
from sarpro_b200.synth import synth_pair

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rows, cols, target = int(os.environ.get("ROWS", 2003)), int(os.environ.get("COLS", 3011)), int(os.environ.get("TARGET", 512))
vv, vh = synth_pair(rows, cols, block=32)
ctx = S.Context(lr)
uid = [S.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(uid[0], rank, world)
bad = 0
for strategy in (S.CLAHE, S.ROBUST, S.TAMED, S.STANDARD):
    clahe = strategy == S.CLAHE
    h0, h1 = S.shard_halo_rows(rows, cols, target, world, rank, clahe)
    img = ctx.process_synrgb_sharded(vv[h0:h1], vh[h0:h1], rows, strategy, target, True)
    single = ctx.process_synrgb_jpeg(vv, vh, strategy, target, True)
    same = np.array_equal(img.rgb, single.rgb)
    msg = f"rank {rank}/{world} strategy {S.STRATEGY_NAMES[strategy]} rows[{h0},{h1}) sharded==single: {same}"
    if rank == 0:
        from oracle import pyoracle as O
        ref, _ = O.pipeline_synrgb_jpeg(vv.astype(np.float32), vh.astype(np.float32), strategy, target, True)
        ok = np.array_equal(img.rgb, ref)
        msg += f" sharded==oracle: {ok}"
        same = same and ok
    print(msg, flush=True)
    bad += 0 if same else 1
t = torch.tensor([bad], device="cuda")
dist.all_reduce(t)
ctx.comm_destroy()
ctx.close()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)

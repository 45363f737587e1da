"""torchrun worker: device timeline (SARPRO_TRACE=1 on rank 0) of the row-band-sharded C3 scene."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
os.environ.setdefault("SARPRO_STAGE_TIMING", "all")  # every launch gets an event pair (the context reads this when it is created)
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rows, cols = 16000, 25000
vv = synth_band_torch(rows, cols, SEED_VV, dev); vh = synth_band_torch(rows, cols, SEED_VH, dev, cross_pol=True)
ctx = S.Context(lr)
uid = [S.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(uid[0], rank, world)
h0, h1 = S.shard_halo_rows(rows, cols, 2048, world, rank, True)
out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
for it in range(6):
    if it == 5 and rank == 0:
        os.environ["SARPRO_TRACE"] = "1"
    dist.barrier(); torch.cuda.synchronize()
    ctx.process_synrgb_sharded(vv[h0:h1], vh[h0:h1], rows, S.CLAHE, 2048, True, out=out)
    if rank == 0:
        print("iter", it, "total_ms", ctx.timing().total_ms, flush=True)
ctx.comm_destroy(); ctx.close(); dist.destroy_process_group()

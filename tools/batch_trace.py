"""torchrun worker: per-rank step times of the scene-per-GPU mode (and the device timeline with SARPRO_TRACE=1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows, cols = 16000, 25000
vv = synth_band_torch(rows, cols, SEED_VV + 2 * rank, dev); vh = synth_band_torch(rows, cols, SEED_VH + 2 * rank, dev, cross_pol=True)
ctx = S.Context(lr)
out = torch.empty((2048, 2048, 3), dtype=torch.uint8, device=dev)
for it in range(4):
    ctx.process_synrgb_jpeg(vv, vh, S.CLAHE, 2048, True, out=out)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
dev_ms = []
for it in range(n):
    ctx.process_synrgb_jpeg(vv, vh, S.CLAHE, 2048, True, out=out)
    dev_ms.append(ctx.timing().total_ms)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / n
print(f"rank {rank}: wall {wall:.3f} ms/step, device total_ms min {min(dev_ms):.3f} median {sorted(dev_ms)[n//2]:.3f} max {max(dev_ms):.3f}; cpus {len(os.sched_getaffinity(0))}", flush=True)
if os.environ.get("TRACE_RANK") == str(rank):
    os.environ["SARPRO_TRACE"] = "1"
    ctx.process_synrgb_jpeg(vv, vh, S.CLAHE, 2048, True, out=out)
ctx.close()
if world > 1:
    dist.destroy_process_group()

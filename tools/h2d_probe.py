"""torchrun worker: pinned host -> device bandwidth of every rank at once (the end-to-end bound of config 5 on one host) and of
rank 0 alone. 1.6 GB per copy (one scene), 5 copies each."""
import os, sys, time
import torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1_600_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device=dev)
def run(k=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return k * n / (time.perf_counter() - t0) / 1e9
run(1)
if world > 1:
    dist.barrier()
together = run()
if world > 1:
    t = torch.tensor([together], device=dev, dtype=torch.float64)
    dist.all_reduce(t)
    total = float(t.item())
    dist.barrier()
    alone = run() if rank == 0 else 0.0
    dist.barrier()
    if rank == 0:
        print(f"h2d probe: {world} ranks at once: {total:.1f} GB/s aggregate ({total / world:.1f} per rank); rank 0 alone: {alone:.1f} GB/s; cpus {os.cpu_count()}", flush=True)
    dist.destroy_process_group()
else:
    print(f"h2d probe: 1 rank: {together:.1f} GB/s", flush=True)

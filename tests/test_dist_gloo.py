"""CPU, world_size 2 over gloo: the multi-rank host logic. Every rank histograms its own row band, the
integer histograms are all-reduced, and each rank's planner must produce the same LUT (bit for bit) as a
single process that saw the whole raster — the property that makes the sharded GPU result identical to
the 1-GPU result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sarpro_b200 as S
    from tests.fixtures import CASES
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dn = CASES["speckle"](403, 517)
        ok = True
        for clahe in (False, True):
            r0, r1 = S.shard_rows(dn.shape[0], world, rank, clahe)
            local = np.bincount(dn[r0:r1].ravel(), minlength=65536).astype(np.int64)
            t = torch.from_numpy(local)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            merged = t.numpy().astype(np.uint64)
            whole = np.bincount(dn.ravel(), minlength=65536).astype(np.uint64)
            ok &= bool(np.array_equal(merged, whole))
            for strategy in (S.ROBUST, S.CLAHE, S.STANDARD):
                st_m, lut_m = S.plan_from_dn_histogram(merged, S.U8, strategy)
                st_w, lut_w = S.plan_from_dn_histogram(whole, S.U8, strategy)
                ok &= bool(np.array_equal(lut_m, lut_w)) and st_m.p99 == st_w.p99 and st_m.low_clip == st_w.low_clip
            # min/max all-reduce of the per-band sample extrema (scale_u16_to_u8 over the whole raster)
            _, lut = S.plan_from_dn_histogram(whole, S.U16, S.DEFAULT)
            q16 = lut[dn[r0:r1]]
            mm = torch.tensor([int(q16.min()), -int(q16.max())])
            dist.all_reduce(mm, op=dist.ReduceOp.MIN)
            ok &= (int(mm[0]), -int(mm[1])) == (int(lut[dn].min()), int(lut[dn].max()))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_histogram_allreduce_gives_identical_plan(lib_built):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, True), (1, True)]

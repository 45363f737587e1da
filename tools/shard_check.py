"""torchrun worker: row-band-sharded synRGB of ONE scene across the ranks vs the single-GPU result (and, for shapes the
oracle finishes in seconds, the oracle on rank 0). Prints one line per rank and shape; exit code 1 on mismatch.

SHAPES="rows x cols x target,..." overrides the default list. The defaults cover: a raster whose width is not a multiple of 8
(generic exact kernel), shapes with cols % 8 == 0 and at least 8 row groups of 16 rows per rank (tensor-core pass B, k_hmma,
on every rank), and with FULL=1 the 16000 x 25000 C3 scene of bench.py (single-GPU result as the reference: that one is
pinned to the oracle by tests/test_gpu_parity.py::test_full_size_scene_against_oracle)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import sarpro_b200 as S
from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch, synth_pair

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
shapes = [(2003, 3011, 512), (4000, 8192, 1024), (1100 * world, 4096, 512), (2100, 8197, 1024)]  # (the last: an odd width on the tensor-core kernel)
if os.environ.get("SHAPES"):
    shapes = [tuple(int(x) for x in s.split("x")) for s in os.environ["SHAPES"].split(",")]
if os.environ.get("FULL"):
    shapes.append((16000, 25000, 2048))
ctx = S.Context(lr)
uid = [S.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(uid[0], rank, world)
bad = 0
for rows, cols, target in shapes:
    big = rows * cols > 40_000_000
    if big:  # device generator (same seeds on every rank -> same scene)
        vv = synth_band_torch(rows, cols, SEED_VV, dev)
        vh = synth_band_torch(rows, cols, SEED_VH, dev, cross_pol=True)
        torch.cuda.synchronize(dev)
    else:
        vv, vh = synth_pair(rows, cols, block=32, point_targets=1e-4)
    for strategy in (S.CLAHE, S.ROBUST, S.TAMED, S.STANDARD):
        clahe = strategy == S.CLAHE
        h0, h1 = S.shard_halo_rows(rows, cols, target, world, rank, clahe)
        img = ctx.process_synrgb_sharded(vv[h0:h1], vh[h0:h1], rows, strategy, target, True)
        t = ctx.timing()
        single = ctx.process_synrgb_jpeg(vv, vh, strategy, target, True)
        same = np.array_equal(img.rgb, single.rgb)
        msg = (f"rank {rank}/{world} {rows}x{cols}->{target} {S.STRATEGY_NAMES[strategy]} rows[{h0},{h1}) sharded==single: {same} "
               f"(host syncs {t.host_syncs}, launches {t.kernel_launches})")
        if rank == 0 and not big:
            from oracle import pyoracle as O
            ref, _ = O.pipeline_synrgb_jpeg(np.asarray(vv).astype(np.float32), np.asarray(vh).astype(np.float32), strategy, target, True)
            ok = np.array_equal(img.rgb, ref)
            msg += f" sharded==oracle: {ok}"
            same = same and ok
        print(msg, flush=True)
        bad += 0 if same else 1
    if not big:
        # scenes whose CLAHE sample extrema are NOT (0, 255): scale_u16_to_u8 re-stretches (autoscale.rs:348-364), which the
        # sharded pipeline only learns from the merged extrema that travel with the rows (repair path of comm.cu)
        flat = np.full((rows, cols), 1000, np.uint16)                                   # one bin: every sample 255 -> 0 after the re-stretch
        two = np.where((np.arange(rows)[:, None] // 97 + np.arange(cols)[None, :] // 131) % 2 == 0, 300, 900).astype(np.uint16)
        tri = np.asarray(vv).copy()
        tri[tri == 0] = 1                                                               # no invalid pixel: the minimum sample may exceed 0
        # a raster whose DN -> bin table does not fit the tensor-core kernel (> 2000 distinct hot DNs): the device planner sets
        # plan->use_generic, the sharded pipeline notices after the exchange and re-runs the band with the generic exact kernel
        wide = np.random.default_rng(rows + cols).integers(1, 40000, (rows, cols)).astype(np.uint16)
        for name, a, b in (("flat", flat, two), ("two", two, tri), ("wide", wide, tri)):
            h0, h1 = S.shard_halo_rows(rows, cols, target, world, rank, True)
            img = ctx.process_synrgb_sharded(a[h0:h1], b[h0:h1], rows, S.CLAHE, target, True)
            t = ctx.timing()
            single = ctx.process_synrgb_jpeg(a, b, S.CLAHE, target, True)
            same = np.array_equal(img.rgb, single.rgb)
            msg = f"rank {rank}/{world} {rows}x{cols}->{target} clahe re-stretch case '{name}' sharded==single: {same} (host syncs {t.host_syncs})"
            if rank == 0:
                from oracle import pyoracle as O
                ref, _ = O.pipeline_synrgb_jpeg(a.astype(np.float32), b.astype(np.float32), S.CLAHE, target, True)
                ok = np.array_equal(img.rgb, ref)
                msg += f" sharded==oracle: {ok}"
                same = same and ok
            print(msg, flush=True)
            bad += 0 if same else 1
    # config 4's call: two polarization operations at full resolution, any contiguous row split, merged scan / stat histograms
    if not big or os.environ.get("POLOPS_BIG"):
        r0, r1 = S.shard_rows(rows, world, rank, False)
        for strategy, bd in ((S.EQUALIZED, S.U16), (S.ROBUST, S.U8), (S.STANDARD, S.U16)):
            ops = (S.OP_LOGRATIO, S.OP_NDIFF)
            mine, st = ctx.process_polops(vv[r0:r1], vh[r0:r1], ops, bd, strategy, scene_rows=rows)
            whole, st1 = ctx.process_polops(vv, vh, ops, bd, strategy)
            same = all(np.array_equal(np.asarray(mine[k]), np.asarray(whole[k])[r0:r1]) for k in range(2))
            same = same and all(st[k].low_clip == st1[k].low_clip and st[k].high_clip == st1[k].high_clip and
                                st[k].valid_count == st1[k].valid_count for k in range(2))
            msg = f"rank {rank}/{world} {rows}x{cols} polops {S.STRATEGY_NAMES[strategy]} u{8 if bd == S.U8 else 16} rows[{r0},{r1}) sharded==single: {same}"
            if rank == 0 and not big:
                from oracle import pyoracle as O
                a32, b32 = np.asarray(vv).astype(np.float32), np.asarray(vh).astype(np.float32)
                ok = True
                for k, op in enumerate(ops):
                    po = O.process_scalar_data_pipeline(O.pol_op(op, a32, b32), bd, strategy, want_db=False)
                    ok = ok and np.array_equal(np.asarray(mine[k]), (po.u8 if bd == S.U8 else po.u16)[r0:r1])
                msg += f" sharded==oracle: {ok}"
                same = same and ok
            print(msg, flush=True)
            bad += 0 if same else 1
    del vv, vh
t = torch.tensor([bad], device="cuda")
dist.all_reduce(t)
ctx.comm_destroy()
ctx.close()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)

"""bench_extra.py — the BASELINE configurations 4 and 5 of bench.py (kept apart from the headline config's code).

c4: full-resolution log-ratio and normalised difference of the VV / VH pair -> Equalized -> two u16 bands, no downsample
    (sentinel1.rs:1497-1579 -> ops.rs:4-44 -> pipeline.rs:42-66 -> save.rs:199-316 without a target size). One call,
    sarpro_pipeline_polops: three passes over the u16 pair (scan, 4096-bin stat histograms, quantise), both operations per pass.
    N > 1: the pair row-sharded over the ranks (any contiguous split; merged scan / histograms by NCCL), "strong".
c5: batch of dual-pol scenes -> CLAHE -> 1024 px padded multiband u8 tiles (params.rs:33 defaults; api/mod.rs:474-536's loop),
    scenes distributed over the GPUs with no collective on the data path ("weak"): sarpro_pipeline_batch.
"""
from __future__ import annotations

import json
import sys


def run_c4(args, ctx, S, dev, stream, rank, world, rows, cols, make_scene, pinned_u16, timed, timed_e2e, with_load_window, peak, peak_src,
           names, METRIC, WORKLOADS):
    import numpy as np
    import torch
    import torch.distributed as dist

    ops = (S.OP_LOGRATIO, S.OP_NDIFF)
    sharded = world > 1
    vv, vh = make_scene(0)
    torch.cuda.synchronize(dev)
    parity_ok = None
    if sharded:
        uid = [S.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        r0, r1 = S.shard_rows(rows, world, rank, False)
    else:
        r0, r1 = 0, rows
    mine = r1 - r0
    outs = [torch.empty((mine, cols), dtype=torch.uint16, device=dev) for _ in ops]
    pvv, pvh = vv[r0:r1], vh[r0:r1]
    step = lambda: ctx.process_polops(pvv, pvh, ops, S.U16, S.EQUALIZED, scene_rows=rows if sharded else None, outs=outs)
    if sharded:
        # parity: this rank's rows of the sharded result against the single-GPU result of the whole pair, computed in this run
        whole = [torch.empty((rows, cols), dtype=torch.uint16, device=dev) for _ in ops]
        ctx.process_polops(vv, vh, ops, S.U16, S.EQUALIZED, outs=whole)
        step()
        torch.cuda.synchronize(dev)
        ok = torch.tensor([1 if all(torch.equal(outs[k].view(torch.int16), whole[k][r0:r1].view(torch.int16)) for k in range(2)) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity_ok = bool(int(ok.item()))
        del whole
        torch.cuda.empty_cache()
        if not parity_ok:
            if rank == 0:
                print(json.dumps({"error": "sharded c4 result differs from the single-GPU result", "n_gpus": world}), file=sys.stderr, flush=True)
            sys.exit(1)
    ms_per_step, acc, clocks = with_load_window(step, args.steps)
    # e2e: pinned host u16 bands in, pinned host u16 bands out
    pvv_np, pvh_np = pinned_u16(pvv), pinned_u16(pvh)
    outs_h = [torch.empty((mine, cols), dtype=torch.int16).pin_memory().numpy().view(np.uint16) for _ in ops]
    e2e_steps = max(2, min(args.steps, 3))
    e2e_ms, h2d, d2h = timed_e2e(lambda: ctx.process_polops(pvv_np, pvh_np, ops, S.U16, S.EQUALIZED, scene_rows=rows if sharded else None,
                                                            outs=outs_h), e2e_steps)
    line = None
    if rank == 0:
        stage_ms, stage_n = acc["stage_ms"], acc["stage_n"]
        n_apply = max(stage_n[names.index("apply")], 1)
        apply_ms = stage_ms[names.index("apply")] / n_apply
        px = mine * cols
        alg_apply = px * (2 + 2 + 2 + 2)        # k_f32_quantize: read both u16 operands once, write two u16 bands
        alg_step = px * 16                      # 3 x (2 + 2) B reads + 2 x 2 B writes
        achieved = alg_apply / (apply_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": round(rows * cols / (ms_per_step * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS["c4"].format(rows=rows, cols=cols), "config": "c4",
                       "cache": "inputs (1.6 GB) and outputs (1.6 GB) exceed the 126 MB L2; no flush needed",
                       "parallelism": (f"the pair row-sharded over {world} GPUs (contiguous row split); NCCL: all-reduce of the scan "
                                       "(min / max / count) and of the two 4096-bin stat histograms") if sharded else "single GPU"},
            "clocks": clocks, "gpu_launches": acc["launches"], "host_syncs_per_step": acc["syncs"] / args.steps,
            "e2e": {"value": round(rows * cols / (e2e_ms * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms, 3),
                    "input": "pinned host u16 DN bands in, pinned host u16 bands out" + (" (each rank its rows)" if sharded else ""),
                    "steps": e2e_steps},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": None, "kernel": "k_f32_quantize<2 ops> (pass 3: both operations of the u16 pair -> thresholds -> two u16 bands)",
                         "algorithmic_bytes_per_launch": alg_apply, "avg_launch_ms": round(apply_ms, 4), "peak_source": peak_src,
                         "stage_ms_per_step": {names[i]: round(stage_ms[i] / args.steps, 4) for i in range(8) if stage_n[i]},
                         "step": {"algorithmic_bytes": alg_step, "achieved": round(alg_step / (ms_per_step * 1e-3) / 1e9, 1),
                                  "frac": round(alg_step / (ms_per_step * 1e-3) / 1e9 / peak, 4),
                                  "note": "whole step against SURVEY 8d's 16 B/px floor (includes the two host planning round trips)"}},
        }
        if sharded:
            line["parity_ok"] = parity_ok
    return line, clocks


def run_c5(args, ctx, S, dev, stream, rank, world, rows, cols, target, make_scene, pinned_u16, timed, timed_e2e, with_load_window, peak,
           peak_src, names, METRIC, WORKLOADS):
    import numpy as np
    import torch
    import torch.distributed as dist

    # scenes per GPU: distinct device-resident scenes for the kernel-only number (bounded by HBM: 1.6 GB each), and the same
    # count of pinned host scenes for the end-to-end number (bounded by host memory: 2 distinct host scenes, reused round-robin)
    per_gpu = args.scenes // world if args.scenes else 8
    per_gpu = max(1, per_gpu)
    n_dev = min(per_gpu, 4)
    dev_scenes = [make_scene(rank * per_gpu + k) for k in range(n_dev)]
    torch.cuda.synchronize(dev)
    oc, orr = S.Context.resize_output_dims(cols, rows, target, True)
    outs = [torch.empty((orr, oc), dtype=torch.uint8, device=dev) for _ in range(2 * per_gpu)]
    scenes = [dev_scenes[k % n_dev] for k in range(per_gpu)]
    step = lambda: ctx.process_batch(scenes, S.BATCH_MULTIBAND, S.U8, S.CLAHE, target, True, outs=outs)
    # parity inside the run: the batch entry against the single-scene entry on the first scene
    res, statuses, rep, _ = step()
    torch.cuda.synchronize(dev)
    single = ctx.process_multiband_tiff(scenes[0][0], scenes[0][1], S.U8, S.CLAHE, target, True)
    parity_ok = (rep.processed == per_gpu and np.array_equal(outs[0].cpu().numpy(), single.gray)
                 and np.array_equal(outs[1].cpu().numpy(), single.gray_band2))
    if world > 1:
        ok = torch.tensor([1 if parity_ok else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity_ok = bool(int(ok.item()))
    if not parity_ok:
        if rank == 0:
            print(json.dumps({"error": "batch result differs from the single-scene result", "n_gpus": world}), file=sys.stderr, flush=True)
        sys.exit(1)

    # one timed call = one batch of per_gpu scenes; the line reports per SCENE (one bench step = one scene)
    batch_ms, acc, clocks = with_load_window(step, args.steps)
    ms_per_scene = batch_ms / per_gpu
    # e2e: pinned host scenes through the same entry (uploads double-buffered beside the kernels), host tiles out
    host_scenes = [(pinned_u16(s[0]), pinned_u16(s[1])) for s in dev_scenes[:2]]
    hs = [host_scenes[k % len(host_scenes)] for k in range(per_gpu)]
    outs_h = [torch.empty((orr, oc), dtype=torch.uint8).pin_memory().numpy() for _ in range(2 * per_gpu)]
    e2e_ms, h2d, d2h = timed_e2e(lambda: ctx.process_batch(hs, S.BATCH_MULTIBAND, S.U8, S.CLAHE, target, True, outs=outs_h), 2)
    # the same scenes one call at a time (no overlap of the next upload): what the batch entry buys
    def one_by_one():
        for k in range(per_gpu):
            ctx.process_multiband_tiff(hs[k][0], hs[k][1], S.U8, S.CLAHE, target, True)
    serial_ms, _, _ = timed_e2e(one_by_one, 2)
    line = None
    if rank == 0:
        stage_ms, stage_n = acc["stage_ms"], acc["stage_n"]
        n_apply = max(stage_n[names.index("apply")], 1)
        apply_ms = stage_ms[names.index("apply")] / n_apply
        alg_bytes = rows * cols * 2 + rows * oc
        achieved = alg_bytes / (apply_ms * 1e-3) / 1e9
        scenes_s = world * 1e3 / ms_per_scene
        line = {
            "metric": METRIC, "value": round(world * rows * cols / (ms_per_scene * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "n_gpus": world,
            "steps": args.steps * per_gpu, "warmup": args.warmup * per_gpu, "ms_per_step": round(ms_per_scene, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": WORKLOADS["c5"].format(rows=rows, cols=cols), "config": "c5", "step": "one scene",
                       "scenes_per_gpu": per_gpu, "scenes_per_s": round(scenes_s, 1),
                       "cache": "inputs (1.6 GB per scene) exceed the 126 MB L2; no flush needed",
                       "parallelism": ("single GPU" if world == 1 else
                                       f"scenes distributed over {world} GPUs ({per_gpu} per GPU), no collective on the data path: replicas only")},
            "clocks": clocks, "gpu_launches": acc["launches"], "host_syncs_per_step": acc["syncs"] / (args.steps * per_gpu),
            "e2e": {"value": round(world * per_gpu * rows * cols / (e2e_ms * 1e-3) / 1e6, 1), "unit": "Mpixel/s",
                    "h2d_bytes_per_step": h2d // per_gpu, "d2h_bytes_per_step": d2h // per_gpu, "ms_per_step": round(e2e_ms / per_gpu, 3),
                    "scenes_per_s": round(world * per_gpu * 1e3 / e2e_ms, 2),
                    "input": "pinned host u16 DN scenes through sarpro_pipeline_batch (next scene's upload beside the kernels), host u8 tiles out",
                    "one_call_per_scene_ms": round(serial_ms / per_gpu, 3), "steps": 2 * per_gpu},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": None, "kernel": "k_hmma<CLAHE> (pass B), one launch per band",
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": round(apply_ms, 4), "peak_source": peak_src,
                         "stage_ms_per_step": {names[i]: round(stage_ms[i] / (args.steps * per_gpu), 4) for i in range(8) if stage_n[i]}},
            "parity_ok": parity_ok,
        }
    return line, clocks


def run_read(args, ctx, S, dev, stream, rank, world, rows, cols, target, make_scene, pinned_u16, timed, timed_e2e, with_load_window, peak,
             peak_src, names, METRIC, WORKLOADS):
    """The CLI's --size flow on one GPU (replicas at N > 1): sarpro_read_band_resampled (Average) for both bands, the f32 rasters
    stay on the device, then sarpro_pipeline_synrgb at the same target. The oracle of the same chain is timed beside it on a crop."""
    import time

    import numpy as np
    import torch
    vv, vh = make_scene(rank)
    torch.cuda.synchronize(dev)
    oc, orr, alg = S.Context.read_dims_for_target(cols, rows, target)
    small = [torch.empty((orr, oc), dtype=torch.float32, device=dev) for _ in range(2)]
    poc, porr = S.Context.resize_output_dims(oc, orr, target, True)
    out_dev = torch.empty((porr, poc, 3), dtype=torch.uint8, device=dev)
    read_ms = [0.0, 0]

    def step():
        for k, band in enumerate((vv, vh)):
            ctx.read_band_resampled(band, oc, orr, alg, out=small[k])
            t = ctx.timing()
            read_ms[0] += t.stage_ms[names.index("convert")]
            read_ms[1] += 1
        ctx.process_synrgb_jpeg(small[0], small[1], S.CLAHE, target, True, out=out_dev)

    ms_per_step, acc, clocks = with_load_window(step, args.steps)
    vv_np, vh_np = pinned_u16(vv), pinned_u16(vh)
    out_h = torch.empty((porr, poc, 3), dtype=torch.uint8).pin_memory().numpy()

    def step_host():
        for k, band in enumerate((vv_np, vh_np)):
            ctx.read_band_resampled(band, oc, orr, alg, out=small[k])
        ctx.process_synrgb_jpeg(small[0], small[1], S.CLAHE, target, True, out=out_h)

    t0 = time.perf_counter()
    step_host()
    step_host()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / 2
    line = None
    if rank == 0:
        k_ms = read_ms[0] / max(read_ms[1], 1)
        alg_bytes = rows * cols * 2 + oc * orr * 4
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": round(world * rows * cols / (ms_per_step * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS["read"].format(rows=rows, cols=cols), "config": "read",
                       "cache": "inputs (1.6 GB per scene) exceed the 126 MB L2; no flush needed",
                       "parallelism": "single GPU" if world == 1 else f"replicas only: one scene per GPU x{world}",
                       "resampler": "average" if alg == 0 else "lanczos", "read_shape": [orr, oc]},
            "clocks": clocks, "gpu_launches": None,
            "e2e": {"value": round(world * rows * cols / (e2e_ms * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": rows * cols * 4,
                    "d2h_bytes_per_step": porr * poc * 3, "ms_per_step": round(e2e_ms, 3), "input": "pinned host u16 DN bands", "steps": 2},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": None, "kernel": "k_read_average<u16> (one launch per band: the raster read once, f64 accumulation in GDAL's order)",
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": round(k_ms, 4), "peak_source": peak_src},
        }
        # the oracle of the same chain on a crop (serial, like GDAL's RasterIO resampling and the reference's raster path)
        from oracle import pyoracle as O
        cr, cc = 4000, 6250
        a = vv[:cr, :cc].cpu().numpy().view(np.uint16)
        b = vh[:cr, :cc].cpu().numpy().view(np.uint16)
        coc, corr, calg = O.read_dims_for_target(cc, cr, 512)
        t0 = time.perf_counter()
        r1 = O.read_band_resampled(a, coc, corr, calg)
        r2 = O.read_band_resampled(b, coc, corr, calg)
        O.pipeline_synrgb_jpeg(r1, r2, O.CLAHE, 512, True)
        dt = time.perf_counter() - t0
        import os
        line["cpu_baseline"] = {"value": round(cr * cc / dt / 1e6, 3), "unit": "Mpixel/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": f"{cr}x{cc} crop per band, Average to 512px on read (OpenMP over output rows) then synRGB+CLAHE + pad, {dt:.2f} s"}
        # launches of one step, counted on a separate step: one resampling kernel per band + the pipeline's
        per_step = 0
        for k, band in enumerate((vv, vh)):
            ctx.read_band_resampled(band, oc, orr, alg, out=small[k])
            per_step += ctx.timing().kernel_launches
        ctx.process_synrgb_jpeg(small[0], small[1], S.CLAHE, target, True, out=out_dev)
        per_step += ctx.timing().kernel_launches
        line["gpu_launches"] = per_step * args.steps
    return line, clocks

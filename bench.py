#!/usr/bin/env python
"""bench.py — headline benchmark of the raster hot path (BASELINE.json metric).

Metric: Mpixel/s of the ~400 MP dual-pol synRGB + CLAHE pipeline (25,000 x 16,000 u16 per band ->
CLAHE autoscale -> Lanczos3 to 2048 px long side -> pad to square -> suppressed synthetic RGB),
scene pixels (the pair counted once) per second. One "step" = one pass of that pipeline over one
synthetic scene.

  value   whole-job Mpixel/s with the two bands already resident in HBM (device pointers through the
          C ABI), timed with CUDA events on the stream the library launches on, max over ranks.
  e2e     the same call with HOST (pinned) u16 bands and a host RGB result: H2D + kernels + D2H.
  roofline  the dominant kernel of the step (pass B: CLAHE apply fused with the horizontal Lanczos),
          algorithmic bytes / its mean CUDA-event duration inside the timed region, vs MEASURED_PEAKS.
  cpu_baseline  the CPU oracle (a C++ restatement of the reference's serial path) on a bounded crop of
          the same scene, on this box's host cores.

N > 1 (torchrun, one rank per GPU): `value` / `e2e` are the batch mode of BASELINE config 5 — one whole scene per rank per
step, no collective on the data path ("weak" scaling: per-GPU work fixed). The same run also times ONE scene
row-band-sharded over the ranks (config 3: every rank holds its band plus the Lanczos halo, the library all-reduces the
integer DN / CLAHE-tile histograms, the CLAHE min/max and the resized rows over NCCL; "strong" scaling, result bit-identical
to 1 GPU) and reports it under "sharded_mode".
`--impl reference` times the oracle (the reference cannot be built here: no Rust toolchain) on a
bounded sample with all host threads its threaded stage (the Lanczos resize) can use.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS, COLS, TARGET = 16000, 25000, 2048
METRIC = "Mpixel/s, 400MP dual-pol synRGB+CLAHE 2048px"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [s for (t, s) in self.samples if t0 - 0.05 <= t <= t1 + 0.15] or [s for (_, s) in self.samples]
        for s in rows:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(vv_crop, vh_crop, threads):
    """Oracle (reference restatement) on a crop: save.rs:317-368 order, CLAHE, 2048-proportional target."""
    import numpy as np
    from oracle import pyoracle as O
    O.set_resize_threads(threads)
    rows, cols = vv_crop.shape
    target = max(64, int(round(TARGET * cols / COLS)))
    a = vv_crop.astype(np.float32)
    b = vh_crop.astype(np.float32)
    t0 = time.perf_counter()
    O.pipeline_synrgb_jpeg(a, b, O.CLAHE, target, True)
    dt = time.perf_counter() - t0
    return rows * cols / dt / 1e6, dt, target


def run_reference(args, rank, world):
    if rank != 0:
        return
    import numpy as np
    from sarpro_b200.synth import synth_pair
    cores = os.cpu_count() or 1
    rows, cols = 4000, 6250  # 1/16 of the full scene: 25 MP per band
    vv, vh = synth_pair(rows, cols)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, target = cpu_sample(vv, vh, cores)
        if i >= args.warmup:
            vals.append((v, dt))
    mpx = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(d for _, d in vals) / len(vals)
    sample = f"{rows}x{cols} crop per band (1/16 of the scene), synRGB+CLAHE -> {target}px + pad, per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mpx, 3), "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 dual-pol 25000x16000 u16 -> CLAHE -> 2048px + pad -> synRGB (reference timed on a bounded crop)"},
        "cpu_baseline": {"value": round(mpx, 3), "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "reference path is serial (SURVEY F1); only the Lanczos stage uses the threads"},
        "e2e": {"value": round(mpx, 3), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS)
    ap.add_argument("--cols", type=int, default=COLS)
    ap.add_argument("--strategy", default="clahe")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import sarpro_b200 as S
    from sarpro_b200.synth import SEED_VH, SEED_VV, synth_band_torch

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rows, cols = args.rows, args.cols
    strategy = S.STRATEGY_NAMES.index(args.strategy)

    sharded = world > 1
    ctx = S.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    oc, orr = S.Context.resize_output_dims(cols, rows, TARGET, True)
    out_dev = torch.empty((orr, oc, 3), dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def make_scene(scene):
        return (synth_band_torch(rows, cols, SEED_VV + 2 * scene, dev),
                synth_band_torch(rows, cols, SEED_VH + 2 * scene, dev, cross_pol=True))

    def timed(fn, steps, warmup):
        """K steps of fn() between barriers; CUDA events on the library's stream; max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = {"stage_ms": [0.0] * 8, "stage_n": [0] * 8, "launches": 0, "syncs": 0, "h2d": 0, "d2h": 0}
        w0 = time.time()
        ev0.record(stream)
        for _ in range(steps):
            fn()
            t = ctx.timing()
            acc["launches"] += t.kernel_launches
            acc["syncs"] += t.host_syncs
            acc["h2d"], acc["d2h"] = int(t.h2d_bytes), int(t.d2h_bytes)
            for i in range(8):
                acc["stage_ms"][i] += t.stage_ms[i]
                acc["stage_n"][i] += t.stage_launches[i]
        ev1.record(stream)
        barrier()
        w1 = time.time()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms / steps, acc, (w0, w1)

    sampler = ClockSampler(local_rank)
    # Primary leg at every N: one whole scene per rank (BASELINE config 5 at N > 1: scenes distributed over the GPUs, no
    # collective on the data path) -> "weak" scaling, value = all ranks' scene pixels / max-over-ranks time.
    vv, vh = make_scene(rank)
    torch.cuda.synchronize(dev)  # generated on torch's current stream; the library reads them on `stream`
    h0, h1 = 0, rows
    step = lambda: ctx.process_synrgb_jpeg(vv, vh, strategy, TARGET, True, out=out_dev)

    # ---------------- kernel-only leg: inputs resident in HBM ---------------------------------
    # The timed region is short (K steps of ~2 ms) against nvidia-smi's 100 ms sampling period, so the same step keeps
    # running untimed for ~0.7 s before and ~0.5 s after it: the clock samples are taken under the load that is timed, the
    # timed region in the middle of it. (Every rank runs the same number of load steps.)
    for _ in range(args.warmup):
        step()
    if rank == 0:
        sampler.start()
    load0 = time.time()
    for _ in range(400):
        step()
    ms_per_step, acc, (wall0, wall1) = timed(step, args.steps, 0)
    for _ in range(300):
        step()
    clocks = sampler.stop(load0 + 0.25, time.time()) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "same step looped for %.1f s around the timed region" % (time.time() - load0)
    value = world * rows * cols / (ms_per_step * 1e-3) / 1e6   # one scene per rank per step
    stage_ms, stage_n, launches, syncs = acc["stage_ms"], acc["stage_n"], acc["launches"], acc["syncs"]

    # Secondary leg (N > 1): ONE scene row-band-sharded over the ranks (BASELINE config 3), "strong" scaling: every rank
    # holds its band + the Lanczos halo; the library all-reduces the DN / CLAHE-tile histograms, the CLAHE min/max and the
    # resized rows over NCCL; the result is bit-identical to 1 GPU.
    shard = None
    if sharded:
        uid = [S.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        svv, svh = (vv, vh) if rank == 0 else make_scene(0)   # every rank synthesises the same scene and keeps its rows
        torch.cuda.synchronize(dev)
        sh0, sh1 = S.shard_halo_rows(rows, cols, TARGET, world, rank, strategy == S.CLAHE)
        pvv, pvh = svv[sh0:sh1], svh[sh0:sh1]                 # contiguous row slices (views)
        sstep = lambda: ctx.process_synrgb_sharded(pvv, pvh, rows, strategy, TARGET, True, out=out_dev)
        sms, sacc, _ = timed(sstep, args.steps, args.warmup)
        ph_vv = torch.empty((sh1 - sh0, cols), dtype=torch.int16).pin_memory()
        ph_vh = torch.empty((sh1 - sh0, cols), dtype=torch.int16).pin_memory()
        ph_vv.copy_(pvv)
        ph_vh.copy_(pvh)
        out_hs = torch.empty((orr, oc, 3), dtype=torch.uint8).pin_memory().numpy()
        sest = lambda: ctx.process_synrgb_sharded(ph_vv.numpy().view(np.uint16), ph_vh.numpy().view(np.uint16), rows, strategy, TARGET,
                                                  True, out=out_hs)
        sest()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            sest()
        barrier()
        se2e = torch.tensor([(time.perf_counter() - t0) * 1e3 / 3], device=dev, dtype=torch.float64)
        dist.all_reduce(se2e, op=dist.ReduceOp.MAX)
        shard = {"value": round(rows * cols / (sms * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "ms_per_step": round(sms, 4), "scaling": "strong",
                 "e2e_ms_per_step": round(float(se2e.item()), 3), "e2e_value": round(rows * cols / (float(se2e.item()) * 1e-3) / 1e6, 1),
                 "stage_ms_per_step": {S._ffi.STAGE_NAMES[i]: round(sacc["stage_ms"][i] / args.steps, 4) for i in range(8) if sacc["stage_n"][i]},
                 "parallelism": f"ONE scene row-band-sharded over {world} GPUs; NCCL all-reduce of DN / CLAHE-tile histograms, min/max, resized rows; "
                                "bit-identical to 1 GPU"}
        del pvv, pvh, svv, svh, ph_vv, ph_vh

    # ---------------- end-to-end leg: host buffers through the C ABI ----------------------------
    e2e_steps = max(2, min(args.steps, 5))
    lrows = h1 - h0
    vv_h = torch.empty((lrows, cols), dtype=torch.int16).pin_memory()
    vh_h = torch.empty((lrows, cols), dtype=torch.int16).pin_memory()
    vv_h.copy_(vv)
    vh_h.copy_(vh)
    vv_np = vv_h.numpy().view(np.uint16)
    vh_np = vh_h.numpy().view(np.uint16)
    out_h = torch.empty((orr, oc, 3), dtype=torch.uint8).pin_memory().numpy()
    estep = lambda: ctx.process_synrgb_jpeg(vv_np, vh_np, strategy, TARGET, True, out=out_h)
    estep()  # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(e2e_steps):
        estep()
        t = ctx.timing()
        h2d, d2h = int(t.h2d_bytes), int(t.d2h_bytes)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_ms, float(h2d), float(d2h)], device=dev, dtype=torch.float64)
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        e2e_ms = float(mx[0].item())
        h2d, d2h = int(tt[1].item()), int(tt[2].item())
    e2e_value = world * rows * cols / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = load_peaks()
        # dominant kernel: pass B (apply fused with horizontal Lanczos), one launch per band.
        n_apply = max(stage_n[S._ffi.STAGE_NAMES.index("apply")], 1)
        apply_ms = stage_ms[S._ffi.STAGE_NAMES.index("apply")] / n_apply
        alg_bytes = (h1 - h0) * cols * 2 + (h1 - h0) * oc * 1  # read the held u16 DN rows once, write the h-resized u8 rows
        achieved = alg_bytes / (apply_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at the full C3 size, from the
        # ncu --set full capture summarised in profiles/ (same command line); None for other sizes
        full_c3 = (rows, cols, args.strategy) == (ROWS, COLS, "clahe")
        traffic = 871_000_000 if full_c3 else None  # profiles/r01q_ncu_full_hmma.md: 859.7 MB read + 11.3 MB written
        roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": traffic, "kernel": "k_hmma<CLAHE> (pass B: CLAHE apply fused with the horizontal Lanczos pass on IMMA.16832)",
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": round(apply_ms, 4), "peak_source": peak_src,
                    # what the same ncu capture names as the busiest unit of this kernel (a note, not a second roofline)
                    "limiter": ("L1TEX LSU data pipe (shared-memory table gathers): l1tex__data_pipe_lsu_wavefronts 74 % of peak on "
                                "average, 84 % on the busiest SM; DRAM 19 % (profiles/r01q_ncu_full_hmma.md)") if full_c3 else None,
                    "stage_ms_per_step": {S._ffi.STAGE_NAMES[i]: round(stage_ms[i] / args.steps, 4) for i in range(8) if stage_n[i]}}
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": f"C3: dual-pol VV+VH {cols}x{rows} u16 GRD-like -> {args.strategy} autoscale -> Lanczos3 {TARGET}px "
                                   f"+ pad -> synRGB" + (f"; one scene per GPU per step on {world} GPUs, no collective on the data path (config 5); the row-band-sharded single scene (config 3) is under sharded_mode" if sharded else "; one scene on one GPU"),
                       "cache": "inputs (1.6 GB per scene) exceed the 126 MB L2; no flush needed",
                       "scene_bytes": rows * cols * 4, "parallelism": f"scene-per-GPU x{world}" if sharded else "single GPU"},
            "clocks": clocks, "gpu_launches": launches, "host_syncs_per_step": syncs / args.steps,
            "e2e": {"value": round(e2e_value, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms, 3), "input": "pinned host u16 DN bands", "steps": e2e_steps},
            "roofline": roofline,
        }
        if shard:
            line["sharded_mode"] = shard
        if not args.no_cpu_baseline and world == 1:
            crop_r, crop_c = 8000, 6250
            vvc = vv[:crop_r, :crop_c].cpu().numpy().view(np.uint16)
            vhc = vh[:crop_r, :crop_c].cpu().numpy().view(np.uint16)
            v, dt, target = cpu_sample(vvc, vhc, 1)
            line["cpu_baseline"] = {"value": round(v, 3), "unit": "Mpixel/s", "cores": 1, "kind": "port",
                                    "sample": f"{crop_r}x{crop_c} crop per band of the same scene -> {target}px + pad, {dt:.1f} s, serial like the reference (SURVEY F1)"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

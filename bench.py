#!/usr/bin/env python
"""bench.py — headline benchmark of the raster hot path (BASELINE.json metric), one JSON line per run.

Metric: Mpixel/s of the ~400 MP dual-pol synRGB + CLAHE pipeline (25,000 x 16,000 u16 per band -> CLAHE autoscale ->
Lanczos3 to 2048 px long side -> pad to square -> suppressed synthetic RGB), scene pixels (the pair counted once) per
second. One "step" = one pass of that pipeline over one synthetic scene.

  value   whole-job Mpixel/s with the bands already resident in HBM (device pointers through the C ABI), timed with CUDA
          events on the stream the library launches on, max over ranks.
  e2e     the same call with HOST (pinned) u16 bands and a host RGB result: H2D + kernels + D2H inside the timed region.
  roofline  the dominant kernel of the step (pass B: CLAHE apply fused with the horizontal Lanczos pass), algorithmic bytes /
          its mean CUDA-event duration inside the timed region, vs MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (a C++ restatement of the reference's serial path) on a bounded crop of the same scene.

--config selects the BASELINE.json configuration (default c3, the one the metric is quoted on):
  c1  VV 4096 x 4096 -> Standard -> u8 gray, no resize (the reference's CPU-runnable case)
  c2  as c3 with Robust autoscale (default synRGB map)
  c3  dual-pol 25,000 x 16,000 -> CLAHE -> 2048 px + pad -> suppressed synRGB
  c4  full-resolution log-ratio and n-diff of the pair -> Equalized -> two u16 bands, no downsample
  c5  batch of scenes -> 1024 px padded multiband u8 tiles (reference defaults, params.rs:33), scenes distributed over the GPUs
  read  the CLI's --size flow (not a BASELINE config): both bands averaged down on read, then the synRGB pipeline

N > 1 (torchrun, one rank per GPU). c3 / c2: `value` is ONE scene row-band-sharded over the ranks ("strong" scaling: every
rank holds its band plus the Lanczos halo; integer histogram / CLAHE-tile all-reduces and one all-gather of the owned output rows
inside the library; the result is compared byte for byte with the single-GPU result of the same scene in this run —
"parity_ok"; the line is not printed if that fails). The scenes-per-GPU replicas (config 5's distribution, no collective on the
data path) are reported under "batch_mode". c1: independent replicas. c4: the pair row-sharded. c5: scenes distributed.

`--impl reference` times the oracle (the reference cannot be built here: no Rust toolchain) on a bounded sample with all host
threads its threaded stage (the Lanczos resize) can use.
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


# ------------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (reference restatement) on a bounded sample
# ------------------------------------------------------------------------------------------------------------------------
def cpu_sample(cfg, vv_crop, vh_crop, threads, full_cols):
    """One pass of config `cfg` of the oracle on a crop; returns (Mpixel/s, seconds, note)."""
    import numpy as np
    from oracle import pyoracle as O
    O.set_resize_threads(threads)
    rows, cols = vv_crop.shape
    a = vv_crop.astype(np.float32)
    b = vh_crop.astype(np.float32) if vh_crop is not None else None
    t0 = time.perf_counter()
    if cfg == "c1":
        O.pipeline_single(a, O.TIFF, O.U8, O.STANDARD, None, False)
        note = "Standard -> u8, no resize"
    elif cfg == "c4":
        for op in (O.OP_LOGRATIO, O.OP_NDIFF):
            O.pipeline_single(O.pol_op(op, a, b), O.TIFF, O.U16, O.EQUALIZED, None, False)
        note = "log-ratio + n-diff -> Equalized -> u16, no resize"
    elif cfg == "c5":
        target = max(64, int(round(1024 * cols / full_cols)))
        O.pipeline_multiband_tiff(a, b, O.U8, O.CLAHE, target, True)
        note = f"CLAHE -> {target}px + pad multiband u8"
    else:
        target = max(64, int(round(TARGET * cols / full_cols)))
        O.pipeline_synrgb_jpeg(a, b, O.ROBUST if cfg == "c2" else O.CLAHE, target, True)
        note = f"synRGB+{'Robust' if cfg == 'c2' else 'CLAHE'} -> {target}px + pad"
    dt = time.perf_counter() - t0
    return rows * cols / dt / 1e6, dt, note


WORKLOADS = {
    "c1": "C1: single-band VV 4096x4096 u16 -> dB -> Standard autoscale -> u8 gray (no resize)",
    "c2": "C2: dual-pol VV+VH {cols}x{rows} u16 GRD-like -> Robust autoscale -> Lanczos3 2048px + pad -> synRGB",
    "c3": "C3: dual-pol VV+VH {cols}x{rows} u16 GRD-like -> CLAHE autoscale -> Lanczos3 2048px + pad -> synRGB",
    "c4": "C4: full-resolution log-ratio and n-diff of VV/VH {cols}x{rows} u16 -> Equalized autoscale -> two u16 bands (no downsample)",
    "c5": "C5: batch of dual-pol {cols}x{rows} scenes -> CLAHE -> 1024px padded multiband u8 tiles, scenes distributed over the GPUs",
    "read": "downsample-on-read (SURVEY f2): VV+VH {cols}x{rows} u16 -> GDAL-style Average to 2048px long side (f32) -> CLAHE -> pad -> synRGB",
}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from sarpro_b200.synth import synth_pair
    cores = os.cpu_count() or 1
    cfg = args.config
    rows, cols = (2048, 2048) if cfg == "c1" else (4000, 6250)  # 1/4 of C1; 1/16 of the full scene: 25 MP per band
    full_cols = 4096 if cfg == "c1" else COLS
    vv, vh = synth_pair(rows, cols)
    vals = []
    note = ""
    for i in range(args.warmup + args.steps):
        v, dt, note = cpu_sample(cfg, vv, None if cfg == "c1" else vh, cores, full_cols)
        if i >= args.warmup:
            vals.append((v, dt))
    mpx = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(d for _, d in vals) / len(vals)
    sample = f"{rows}x{cols} crop per band, {note}, per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mpx, 3), "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "strong" if (cfg in ("c2", "c3", "c4") and args.gpus > 1) else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg].format(rows=ROWS, cols=COLS) + " (reference timed on a bounded crop)"},
        "cpu_baseline": {"value": round(mpx, 3), "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "reference path is serial (SURVEY F1); only the Lanczos stage uses the threads"},
        "e2e": {"value": round(mpx, 3), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=globals().get('_REAL_STDOUT', sys.stdout), flush=True)


# ------------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=ROWS)
    ap.add_argument("--cols", type=int, default=COLS)
    ap.add_argument("--scenes", type=int, default=0, help="c5: scenes in the batch (default 8 per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-load-window", action="store_true", help="skip the ~1 s of untimed steps around the timed region (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # ONE JSON line on stdout, whatever the libraries print there (NCCL writes its version banner to stdout): everything else
    # goes to stderr, the line is written to the original stdout at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    globals()["_REAL_STDOUT"] = real_stdout
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

    cfg = args.config
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rows, cols = (4096, 4096) if cfg == "c1" else (args.rows, args.cols)
    strategy = {"c1": S.STANDARD, "c2": S.ROBUST, "c3": S.CLAHE, "c4": S.EQUALIZED, "c5": S.CLAHE, "read": S.CLAHE}[cfg]
    target = 1024 if cfg == "c5" else TARGET

    if cfg == "read":
        os.environ.setdefault("SARPRO_STAGE_TIMING", "all")  # the resampling kernel is timed under the `convert` stage
    ctx = S.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    names = S._ffi.STAGE_NAMES

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def make_scene(scene):
        return (synth_band_torch(rows, cols, SEED_VV + 2 * scene, dev),
                synth_band_torch(rows, cols, SEED_VH + 2 * scene, dev, cross_pol=True))

    def pinned_u16(t):
        h = torch.empty(tuple(t.shape), dtype=torch.int16).pin_memory()
        h.copy_(t)
        return h.numpy().view(np.uint16)

    def timed(fn, steps, warmup):
        """K steps of fn() between barriers; CUDA events on the library's stream; max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = {"stage_ms": [0.0] * 8, "stage_n": [0] * 8, "launches": 0, "syncs": 0}
        ev0.record(stream)
        for _ in range(steps):
            fn()
            t = ctx.timing()
            acc["launches"] += t.kernel_launches
            acc["syncs"] += t.host_syncs
            for i in range(8):
                acc["stage_ms"][i] += t.stage_ms[i]
                acc["stage_n"][i] += t.stage_launches[i]
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms / steps, acc

    def timed_e2e(fn, steps):
        """Wall clock of `steps` calls with host buffers (H2D + kernels + D2H, the call returns after its last copy), max over ranks."""
        fn()  # warm-up (allocations)
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        for _ in range(steps):
            fn()
            t = ctx.timing()
            h2d, d2h = int(t.h2d_bytes), int(t.d2h_bytes)
        barrier()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        if world > 1:
            tt = torch.tensor([ms, float(h2d), float(d2h)], device=dev, dtype=torch.float64)
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            ms, h2d, d2h = float(mx[0].item()), int(tt[1].item()), int(tt[2].item())
        return ms, h2d, d2h

    def with_load_window(step, steps):
        """The timed region is short (K steps of ~1 ms) against nvidia-smi's 100 ms sampling period, so the same step keeps
        running untimed for ~0.7 s before and ~0.5 s after it: the clock samples are taken under the load that is timed."""
        sampler = ClockSampler(local_rank)
        for _ in range(args.warmup):
            step()
        if rank == 0:
            sampler.start()
        load0 = time.time()
        n_pre, n_post = (0, 0) if args.no_load_window else (400, 300)
        for _ in range(n_pre):
            step()
        ms, acc = timed(step, steps, 0)
        for _ in range(n_post):
            step()
        clocks = sampler.stop(load0 + 0.25, time.time()) if rank == 0 else None
        if clocks is not None:
            clocks["window"] = "same step looped for %.1f s around the timed region" % (time.time() - load0)
        return ms, acc, clocks

    peak, peak_src = load_peaks()
    e2e_steps = max(2, min(args.steps, 5))
    line = None

    # ====================================================================================================================
    if cfg in ("c2", "c3"):
        oc, orr = S.Context.resize_output_dims(cols, rows, target, True)
        out_dev = torch.empty((orr, oc, 3), dtype=torch.uint8, device=dev)
        sharded = world > 1
        # every rank synthesises scene 0 (same seeds -> the same bytes); a sharded rank only keeps its rows
        vv, vh = make_scene(0)
        torch.cuda.synchronize(dev)  # generated on torch's current stream; the library reads them on `stream`
        single_step = lambda: ctx.process_synrgb_jpeg(vv, vh, strategy, target, True, out=out_dev)
        parity_ok = None
        if not sharded:
            h0, h1 = 0, rows
            step = single_step
            ms_per_step, acc, clocks = with_load_window(step, args.steps)
            vv_np, vh_np = pinned_u16(vv), pinned_u16(vh)
            out_h = torch.empty((orr, oc, 3), dtype=torch.uint8).pin_memory().numpy()
            e2e_ms, h2d, d2h = timed_e2e(lambda: ctx.process_synrgb_jpeg(vv_np, vh_np, strategy, target, True, out=out_h), e2e_steps)
            batch = None
        else:
            uid = [S.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ctx.comm_init(uid[0], rank, world)
            # reference bytes for the parity check: the single-GPU pipeline on the whole scene (itself pinned to the oracle by
            # tests/test_gpu_parity.py::test_full_size_scene_against_oracle)
            single_step()
            torch.cuda.synchronize(dev)
            ref_rgb = out_dev.clone()
            h0, h1 = S.shard_halo_rows(rows, cols, target, world, rank, strategy == S.CLAHE)
            pvv, pvh = vv[h0:h1], vh[h0:h1]  # contiguous row slices (views)
            step = lambda: ctx.process_synrgb_sharded(pvv, pvh, rows, strategy, target, True, out=out_dev)
            out_dev.zero_()
            step()
            torch.cuda.synchronize(dev)
            ok = torch.tensor([1 if torch.equal(out_dev, ref_rgb) else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            parity_ok = bool(int(ok.item()))
            if not parity_ok:
                if rank == 0:
                    print(json.dumps({"error": "sharded result differs from the single-GPU result", "n_gpus": world}), file=sys.stderr, flush=True)
                ctx.close()
                dist.destroy_process_group()
                sys.exit(1)
            # batch mode (scenes per GPU, no collective on the data path): weak scaling, reported beside the headline
            bms, bacc = timed(single_step, args.steps, args.warmup)
            ms_per_step, acc, clocks = with_load_window(step, args.steps)
            pvv_np, pvh_np = pinned_u16(pvv), pinned_u16(pvh)
            out_h = torch.empty((orr, oc, 3), dtype=torch.uint8).pin_memory().numpy()
            e2e_ms, h2d, d2h = timed_e2e(lambda: ctx.process_synrgb_sharded(pvv_np, pvh_np, rows, strategy, target, True, out=out_h), e2e_steps)
            batch = {"value": round(world * rows * cols / (bms * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "ms_per_step": round(bms, 4),
                     "scaling": "weak", "parallelism": f"one whole scene per GPU per step on {world} GPUs, no collective on the data path",
                     "host_syncs_per_step": bacc["syncs"] / args.steps}
        value = rows * cols / (ms_per_step * 1e-3) / 1e6
        e2e_value = rows * cols / (e2e_ms * 1e-3) / 1e6
        if rank == 0:
            stage_ms, stage_n = acc["stage_ms"], acc["stage_n"]
            n_apply = max(stage_n[names.index("apply")], 1)
            apply_ms = stage_ms[names.index("apply")] / n_apply
            alg_bytes = (h1 - h0) * cols * 2 + (h1 - h0) * oc * 1  # read the held u16 DN rows once, write the h-resized u8 rows
            achieved = alg_bytes / (apply_ms * 1e-3) / 1e9
            full = (rows, cols) == (ROWS, COLS) and world == 1
            kname = "k_hmma<CLAHE>" if cfg == "c3" else "k_hmma<LUT>"
            roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                        "traffic": (872_000_000 if cfg == "c3" else None) if full else None,  # profiles/r02p_ncu_full_hmma.md: 861.3 MB read + 10.7 MB written per launch
                        "kernel": f"{kname} (pass B: per-pixel stage fused with the horizontal Lanczos pass on IMMA.16832), one launch per band",
                        "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": round(apply_ms, 4), "peak_source": peak_src,
                        "stage_ms_per_step": {names[i]: round(stage_ms[i] / args.steps, 4) for i in range(8) if stage_n[i]}}
            par = (f"ONE scene row-band-sharded over {world} GPUs (tile-row aligned bands + Lanczos halo); NCCL: fused all-reduce of both bands' DN "
                   "histograms, of their CLAHE tile histograms, one all-gather of the owned output rows with the sample min/max; planned on the device") \
                if sharded else "single GPU"
            line = {
                "metric": METRIC, "value": round(value, 1), "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
                "scaling": "strong" if sharded else "weak",
                "vs_baseline": None, "dtype": "u16", "data": "synthetic",
                "config": {"workload": WORKLOADS[cfg].format(rows=rows, cols=cols), "config": cfg,
                           "cache": "inputs (1.6 GB per scene) exceed the 126 MB L2; no flush needed",
                           "scene_bytes": rows * cols * 4, "parallelism": par},
                "clocks": clocks, "gpu_launches": acc["launches"], "host_syncs_per_step": acc["syncs"] / args.steps,
                "e2e": {"value": round(e2e_value, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": round(e2e_ms, 3), "input": "pinned host u16 DN bands" + (" (each rank its row band + halo)" if sharded else ""),
                        "steps": e2e_steps},
                "roofline": roofline,
            }
            if sharded:
                line["parity_ok"] = parity_ok
                line["batch_mode"] = batch
        if not sharded and rank == 0 and (rows, cols) == (ROWS, COLS) and not args.no_cpu_baseline:
            # the real drop-in boundary of the Rust caller: host Array2<f32> bands (gdal.rs:123); h2d_bytes is what crossed PCIe
            vv_f = torch.empty((rows, cols), dtype=torch.float32).pin_memory()
            vh_f = torch.empty((rows, cols), dtype=torch.float32).pin_memory()
            vv_f.copy_(vv.to(torch.int32).bitwise_and_(0xffff))
            vh_f.copy_(vh.to(torch.int32).bitwise_and_(0xffff))
            f_ms, f_h2d, f_d2h = timed_e2e(lambda: ctx.process_synrgb_jpeg(vv_f.numpy(), vh_f.numpy(), strategy, target, True, out=out_h), 2)
            line["e2e_f32_boundary"] = {"value": round(rows * cols / (f_ms * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "ms_per_step": round(f_ms, 3),
                                        "h2d_bytes_per_step": f_h2d, "d2h_bytes_per_step": f_d2h,
                                        "input": "pinned host f32 bands (the reference's Array2<f32> boundary, gdal.rs:123)" +
                                                 ("; narrowed to u16 DNs by the library's host threads chunk by chunk while the previous chunk uploads"
                                                  if f_h2d < rows * cols * 8 else "; uploaded as f32 (host narrowing off or slower than PCIe on this host)")}
            del vv_f, vh_f

    # ====================================================================================================================
    elif cfg == "c1":
        vv = synth_band_torch(rows, cols, SEED_VV + 2 * rank, dev)
        torch.cuda.synchronize(dev)
        out_dev = torch.empty((rows, cols), dtype=torch.uint8, device=dev)
        step = lambda: ctx.process_single(vv, S.TIFF, S.U8, strategy, None, False, out=out_dev)
        ms_per_step, acc, clocks = with_load_window(step, args.steps)
        vv_np = pinned_u16(vv)
        out_h = torch.empty((rows, cols), dtype=torch.uint8).pin_memory().numpy()
        e2e_ms, h2d, d2h = timed_e2e(lambda: ctx.process_single(vv_np, S.TIFF, S.U8, strategy, None, False, out=out_h), e2e_steps)
        value = world * rows * cols / (ms_per_step * 1e-3) / 1e6
        if rank == 0:
            stage_ms, stage_n = acc["stage_ms"], acc["stage_n"]
            apply_ms = stage_ms[names.index("apply")] / max(stage_n[names.index("apply")], 1)
            alg_bytes = rows * cols * 3  # k_apply_lut: read u16 DN, write u8
            achieved = alg_bytes / (apply_ms * 1e-3) / 1e9
            line = {
                "metric": METRIC, "value": round(value, 1), "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16",
                "data": "synthetic",
                "config": {"workload": WORKLOADS[cfg], "config": cfg,
                           "cache": "the 33.5 MB raster fits the 126 MB L2: a 256 MB buffer is written between the timed steps' inputs? no - "
                                    "steps run back to back and the raster stays L2-resident; this config is launch-latency-bound (SURVEY 8d: 12.8 us of traffic)",
                           "parallelism": "single GPU" if world == 1 else f"replicas only: one raster per GPU x{world}"},
                "clocks": clocks, "gpu_launches": acc["launches"], "host_syncs_per_step": acc["syncs"] / args.steps,
                "e2e": {"value": round(world * rows * cols / (e2e_ms * 1e-3) / 1e6, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms, 3), "input": "pinned host u16 DN band", "steps": e2e_steps},
                "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                             "traffic": None, "kernel": "k_apply_lut (pass B at full resolution: DN -> u8 through the 65,536-entry table)",
                             "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": round(apply_ms, 4), "peak_source": peak_src,
                             "stage_ms_per_step": {names[i]: round(stage_ms[i] / args.steps, 4) for i in range(8) if stage_n[i]},
                             "note": "L2-resident input: the fraction is against the HBM copy peak and may exceed what DRAM alone would allow"},
            }

    # ====================================================================================================================
    elif cfg == "c4":
        from bench_extra import run_c4
        line, clocks = run_c4(args, ctx, S, dev, stream, rank, world, rows, cols, make_scene, pinned_u16, timed, timed_e2e,
                              with_load_window, peak, peak_src, names, METRIC, WORKLOADS)
    elif cfg == "read":
        from bench_extra import run_read
        line, clocks = run_read(args, ctx, S, dev, stream, rank, world, rows, cols, target, make_scene, pinned_u16, timed, timed_e2e,
                                with_load_window, peak, peak_src, names, METRIC, WORKLOADS)
    elif cfg == "c5":
        from bench_extra import run_c5
        line, clocks = run_c5(args, ctx, S, dev, stream, rank, world, rows, cols, target, make_scene, pinned_u16, timed, timed_e2e,
                              with_load_window, peak, peak_src, names, METRIC, WORKLOADS)

    if rank == 0 and line is not None:
        if not args.no_cpu_baseline and world == 1:
            full_cols = 4096 if cfg == "c1" else COLS
            if cfg == "c1":
                crop_r, crop_c = 4096, 4096
                vvc, vhc = vv.cpu().numpy().view(np.uint16), None
            elif cfg in ("c2", "c3"):
                crop_r, crop_c = 8000, 6250
                vvc = vv[:crop_r, :crop_c].cpu().numpy().view(np.uint16)
                vhc = vh[:crop_r, :crop_c].cpu().numpy().view(np.uint16)
            elif cfg == "read":
                vvc = None
            else:
                from sarpro_b200.synth import synth_pair
                crop_r, crop_c = 4000, 6250
                vvc, vhc = synth_pair(crop_r, crop_c)  # same recipe as the device generator (numpy RNG)
            if vvc is None:
                print(json.dumps(line), file=globals().get('_REAL_STDOUT', sys.stdout), flush=True)
                ctx.close()
                return
            v, dt, note = cpu_sample(cfg, vvc, vhc, 1, full_cols)
            line["cpu_baseline"] = {"value": round(v, 3), "unit": "Mpixel/s", "cores": 1, "kind": "port",
                                    "sample": f"{crop_r}x{crop_c} crop per band of the same scene, {note}, {dt:.1f} s, serial like the reference (SURVEY F1)"}
        print(json.dumps(line), file=globals().get('_REAL_STDOUT', sys.stdout), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

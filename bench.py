#!/usr/bin/env python
"""Benchmark of the per-marker quantification hot path (BASELINE.json metric: ROI-pixels/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE config 3, "chip time series" -- per rank 50 timepoints x
4 channels of 4x4 tiles of 2048^2 uint16 (26.8 GB), overlap 102, flat-field correction on,
1792 buttons (56x32), roi_length 72, per-marker fg/bg sums/means.  One step = the whole hot
path over that stack: flat-field max pass -> all-reduce(MAX) -> flat-field apply fused with
stitch -> ROI gather fused with the masked reductions.  Weak scaling: every rank owns its own
block of 50 timepoints; the only collectives are the 2 x float64 MAX all-reduce and the gather
of the per-marker summaries.

value  = ROI pixels of all ranks / (max-over-ranks device time), inputs resident in HBM.
e2e    = same metric from pinned HOST buffers through magnify_b200.pipeline.HostStagedRunner,
         H2D of the tiles and D2H of image + roi + summaries inside the timed region.
roofline = the dominant kernel (flat-field apply + stitch): algorithmic bytes / CUDA-event time.
cpu_baseline = the NumPy oracle (the reference's arithmetic, in RAM, no dask/zarr spill) on a
         bounded sample (one timepoint), timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "roi_pixels_per_sec"
UNIT = "ROI-px/s"

# BASELINE config 3 geometry ("C3")
C3 = dict(c=4, t=50, r=4, cc=4, h=2048, w=2048, overlap=102, rows=56, cols=32, row_dist=126.1, col_dist=232.9,
          roi_length=72, chamber_radius=30, max_button_radius=15)


# BASELINE config 5 geometry, one rank's shard ("C5"): 20480^2 images as 10x10 tiles of 2048^2 with overlap 0,
# 4 channels, 100k beads, roi_length 50 (beads_pipe default, registry.py:572-573); 12 timepoints per rank
# (100 timepoints over 8 GPUs = 12.5).  Not the default workload: `--config c5`.
C5 = dict(c=4, t=12, r=10, cc=10, h=2048, w=2048, overlap=0, n_beads=100000, min_radius=4, max_radius=12,
          roi_length=50)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference's arithmetic, chunked per tile over a thread
# pool the way dask's threaded scheduler runs the reference's per-tile chunks.
# ----------------------------------------------------------------------------------------------
def cpu_hot_path(tiles, flat, dark, overlap, x, y, fg_radius, roi_length, chamber_radius, max_button_radius,
                 threads):
    from concurrent.futures import ThreadPoolExecutor

    from oracle import flatfield as o_ff, reduce as o_red, rois as o_rois, stitch as o_st

    c, t, r, cc, h, w = tiles.shape
    idx = [(a, b, i, j) for a in range(c) for b in range(t) for i in range(r) for j in range(cc)]
    with ThreadPoolExecutor(max_workers=threads) as pool:
        maxima = list(pool.map(lambda k: o_ff.flatfield_maxima(tiles[k], flat, dark), idx))
        m1 = max(m[0] for m in maxima)
        m2 = max(m[1] for m in maxima)
        out = np.empty_like(tiles)

        def apply(k):
            out[k] = o_ff.flatfield_correct(tiles[k], flat, dark, maxima=(m1, m2))

        list(pool.map(apply, idx))
        image = o_st.stitch(out, overlap)
        fg, bg = o_rois.chip_masks(x[:, 0], y[:, 0], fg_radius[:, 0], roi_length, chamber_radius, max_button_radius,
                                   image.shape[-1], image.shape[-2])
        roi = o_rois.gather_rois(image, x, y, roi_length)
        fgt = np.repeat(fg[:, None], t, 1)
        bgt = np.repeat(bg[:, None], t, 1)
        parts = np.array_split(np.arange(roi.shape[0]), max(1, threads))
        stats = list(pool.map(lambda s: o_red.masked_stats(roi[s], fgt[s], bgt[s]) if len(s) else None, parts))
    stats = np.concatenate([s for s in stats if s is not None], 0)
    return image, roi, fg, bg, stats


def cpu_sample_case(cfg, channels, seed=0):
    """Host-generated sample of the workload: one timepoint, `channels` channels."""
    rng = np.random.default_rng(seed)
    c, r, cc, h, w = channels, cfg["r"], cfg["cc"], cfg["h"], cfg["w"]
    tiles = np.clip(rng.normal(400, 20, (c, 1, r, cc, h, w)), 0, 65535).astype(np.uint16)
    from magnify_b200 import synth

    flat, dark = synth.smooth_flat_dark(h, w)
    rows, cols = cfg["rows"], cfg["cols"]
    kh, kw = h - cfg["overlap"], w - cfg["overlap"]
    y0 = (r * kh - (rows - 1) * cfg["row_dist"]) / 2
    x0 = (cc * kw - (cols - 1) * cfg["col_dist"]) / 2
    cy = y0 + np.arange(rows)[:, None] * cfg["row_dist"] + rng.uniform(-2, 2, (rows, cols))
    cx = x0 + np.arange(cols)[None, :] * cfg["col_dist"] + rng.uniform(-2, 2, (rows, cols))
    m = rows * cols
    rad = (10 + (np.add.outer(np.arange(rows), np.arange(cols)) % 6)).reshape(m, 1).astype(np.int32)
    return tiles, flat, dark, cx.reshape(m, 1), cy.reshape(m, 1), rad


def time_cpu(cfg, channels, steps, warmup, threads):
    tiles, flat, dark, x, y, rad = cpu_sample_case(cfg, channels)
    roi_px = x.shape[0] * channels * cfg["roi_length"] ** 2
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cpu_hot_path(tiles, flat, dark, cfg["overlap"], x, y, rad, cfg["roi_length"], cfg["chamber_radius"],
                     cfg["max_button_radius"], threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return roi_px * len(times) / total, total / len(times), roi_px


def run_reference_arm(args, cfg):
    """--impl reference: the reference's CPU arithmetic (oracle port; the reference itself is
    pure Python that cannot be imported here -- xarray/dask/zarr are absent) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    channels = cfg["c"]
    value, sec_per_step, roi_px = time_cpu(cfg, channels, args.steps, args.warmup, threads)
    sample = (f"1 timepoint x {channels} channels of the C3 stack per step ({cfg['r']}x{cfg['cc']} tiles of "
              f"{cfg['h']}x{cfg['w']}, {cfg['rows'] * cfg['cols']} ROIs of {cfg['roi_length']}^2), data in RAM")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(cfg, n_gpus):
    return {
        "workload": "C3 chip time series: flat-field + stitch + ROI gather + fg/bg masks + masked sums/means",
        "tiles_per_rank": [cfg["c"], cfg["t"], cfg["r"], cfg["cc"], cfg["h"], cfg["w"]],
        "overlap": cfg["overlap"], "markers": cfg["rows"] * cfg["cols"], "roi_length": cfg["roi_length"],
        "sharding": f"time x{n_gpus} (each rank its own {cfg['t']} timepoints)",
        "cache": "inputs (26.8 GB/rank at full size) far exceed the 126 MB L2; no explicit flush",
        "arithmetic": "uint16 pixels in and out; flat-field in float64 (exact reference rounding); integer sums",
    }


def c5_config(cfg, n_gpus):
    return {
        "workload": "C5 bead screen shard: flat-field + stitch + bead label raster/masks + ROI gather + masked sums/means",
        "tiles_per_rank": [cfg["c"], cfg["t"], cfg["r"], cfg["cc"], cfg["h"], cfg["w"]], "overlap": cfg["overlap"],
        "markers": cfg["n_beads"], "roi_length": cfg["roi_length"],
        "sharding": f"time x{n_gpus} (each rank its own {cfg['t']} timepoints)",
        "cache": "inputs (40 GB/rank) far exceed the 126 MB L2; no explicit flush",
    }


def run_b200_arm(args, cfg):
    import torch
    import torch.distributed as dist

    from magnify_b200 import _lib, pipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; magnify_b200 has no CPU path")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    from magnify_b200 import numa

    bound_cpus = numa.bind_to_gpu_numa(local_rank) if world > 1 else []
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lib = _lib.load()
    peak, peak_src = load_peaks()

    c, t = cfg["c"], cfg["t"]
    if args.config == "c5":
        case = synth.bead_case(**cfg, seed=rank, device=dev)
        plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark,
                                     device=dev, group=group)
        plan.set_bead_markers(case.beads)
        m = len(case.beads)
    else:
        gen_keys = ("c", "t", "r", "cc", "h", "w", "overlap", "rows", "cols", "row_dist", "col_dist", "roi_length",
                    "chamber_radius", "max_button_radius")
        case = synth.chip_case(**{k: cfg[k] for k in gen_keys}, seed=rank, device=dev)
        plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark,
                                     device=dev, group=group)
        plan.set_chip_markers(case.x, case.y, case.fg_radius, case.chamber_radius, case.max_button_radius)
        m = case.x.shape[0]
    length = case.roi_length
    from magnify_b200 import ops as _ops

    image_out = _ops.alloc_image(plan.image_shape, torch.uint16, dev)
    roi_out = torch.empty((m, c, t, length, length), dtype=torch.uint16, device=dev)
    stats_out = torch.empty((m, c, t, 8), dtype=torch.float64, device=dev)
    gathered = torch.empty((world,) + tuple(stats_out.shape), dtype=torch.float64, device=dev) if world > 1 else None
    symm = None
    if world > 1 and not args.no_fused_gather:
        try:  # summaries all-gathered by the gather kernel itself over NVLink peer memory
            from magnify_b200.dist import SymmetricSummaries

            symm = SymmetricSummaries(stats_out.shape, dev, group)
        except Exception as exc:
            print(f"[bench] symmetric memory unavailable ({exc!r}); using NCCL all_gather", file=sys.stderr)
            symm = None
    roi_px_rank = m * c * t * length * length
    tile_px_rank = case.tiles.numel()
    phi = (plan.image_shape[-1] * plan.image_shape[-2]) / (cfg["r"] * cfg["cc"] * cfg["h"] * cfg["w"])

    def step(record=None):
        if symm is not None:
            plan.run_device(case.tiles, want_roi=True, image_out=image_out, roi_out=roi_out, record=record,
                            peer_stats=symm.peer_blocks)
            symm.barrier()
            return
        plan.run_device(case.tiles, want_roi=True, image_out=image_out, roi_out=roi_out, stats_out=stats_out,
                        record=record)
        if world > 1:
            dist.all_gather_into_tensor(gathered, stats_out, group=group)

    def barrier():
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if symm is not None:   # the fused gather must equal the NCCL all-gather of the same summaries
        plan.run_device(case.tiles, want_roi=False, image_out=image_out, stats_out=stats_out)
        dist.all_gather_into_tensor(gathered, stats_out, group=group)
        torch.cuda.synchronize(dev)
        same = torch.equal(torch.nan_to_num(gathered, nan=-1.0), torch.nan_to_num(symm.gathered, nan=-1.0))
        if not same:
            raise RuntimeError("fused peer-store gather of the summaries differs from the NCCL all_gather")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = lib.mgb_launch_count()
    records = []
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        rec = []
        step(rec)
        records.append(rec)
    end.record()
    barrier()
    launches = lib.mgb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX, group=group)
    elapsed_s = float(elapsed_ms.item()) / 1e3
    value = roi_px_rank * world * args.steps / elapsed_s

    # per-stage device times (this rank), averaged over the timed steps
    stage_ms = {}
    for rec in records:
        for name, a, b in rec:
            stage_ms.setdefault(name, []).append(a.elapsed_time(b))
    stage_ms = {k: sum(v) / len(v) for k, v in stage_ms.items()}
    # algorithmic bytes per stage (SURVEY.md section 8d / DESIGN.md)
    stage_bytes = {
        "flatfield_max": 2.0 * tile_px_rank,
        "flatfield_stitch": (2.0 + 2.0 * phi) * tile_px_rank,
        "roi_gather_stats": 4.0 * roi_px_rank,
    }
    stages = {}
    for name, ms in stage_ms.items():
        entry = {"ms": ms}
        if name in stage_bytes and ms > 0:
            gbs = stage_bytes[name] / (ms * 1e-3) / 1e9
            entry.update({"algorithmic_GB": stage_bytes[name] / 1e9, "GB/s": gbs, "frac_of_peak": gbs / peak})
        stages[name] = entry
    dom = "flatfield_stitch"
    if args.config != "c3":
        stage_bytes["roi_gather_stats"] = 4.0 * roi_px_rank
    achieved = stages.get(dom, {}).get("GB/s")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from ncu --set full
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        key = f"stitch_u16_kernel<1> T={cfg['t']}"
        traffic = tj.get(key)
    roofline = {
        "bound": "hbm", "kernel": "stitch_u16_kernel<MODE=1> (flat-field apply fused with stitch; 1 launch per step, "
                                  "timed together with the 0.03 ms coefficient-table kernel)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
        "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": stage_bytes[dom],
        "avg_launch_ms": stage_ms.get(dom, 0.0),
    }

    # ---- end to end from pinned host buffers -------------------------------------------------
    e2e = None
    if args.config == "c5":
        e2e = {"value": None, "unit": UNIT, "note": "host leg is measured on the default workload (config 3) only"}
    else:
        try:
            e2e = measure_e2e(args, cfg, case, plan, dev, world, rank, group, barrier)
        except Exception as exc:  # keep the device numbers even if the host leg cannot allocate
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)[:300]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == "c3":
        threads = os.cpu_count() or 1
        v, sec, _ = time_cpu(cfg, cfg["c"], steps=4, warmup=1, threads=threads)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"1 timepoint x {cfg['c']} channels of the C3 stack (64 tiles of 2048^2, 1792 ROIs x 4), "
                                  f"4 timed passes of {sec:.2f} s (~{5 * sec * threads:.0f} core-seconds), NumPy oracle over a "
                                  f"{threads}-thread pool, data in RAM"}
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_s * 1e3 / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": dict(workload_config(cfg, world) if args.config == "c3" else c5_config(cfg, world),
                           summary_gather=("fused peer stores over NVLink (symmetric memory)" if symm is not None
                                           else ("nccl all_gather" if world > 1 else "single rank"))),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "stages": stages, "cpu_baseline": cpu_baseline,
        }
    if world > 1:
        dist.barrier(group=group)
        dist.destroy_process_group()
    return line


def measure_e2e(args, cfg, case, plan, dev, world, rank, group, barrier):
    import psutil
    import torch
    import torch.distributed as dist

    from magnify_b200 import pipeline

    c, t = cfg["c"], cfg["t"]
    m = case.x.shape[0]
    length = case.roi_length
    per_t_in = case.tiles[:, :1].numel() * 2
    per_t_out = (plan.image_shape[-1] * plan.image_shape[-2] * c + m * c * length * length) * 2 + m * c * 48
    avail = psutil.virtual_memory().available
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    t_e2e = int(min(t, max(1, (0.25 * avail / local_world) // (per_t_in + per_t_out))))
    if args.e2e_timepoints:
        t_e2e = min(t, args.e2e_timepoints)
    if world > 1:  # every rank must agree on the shape of the collective-free e2e problem
        tt = torch.tensor([t_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MIN, group=group)
        t_e2e = int(tt.item())
    shape = (c, t_e2e) + tuple(case.tiles.shape[2:])
    sub = pipeline.QuantifyPlan(shape, case.overlap, length, case.flat, case.dark, device=dev, group=group)
    sub.set_chip_markers(case.x[:, :t_e2e], case.y[:, :t_e2e], case.fg_radius, case.chamber_radius,
                         case.max_button_radius)
    # free the device-resident buffers of the first leg before allocating the runner's
    tiles_host = torch.empty(shape, dtype=torch.uint16, pin_memory=True)
    tiles_host.copy_(case.tiles[:, :t_e2e])
    torch.cuda.synchronize(dev)
    runner = pipeline.HostStagedRunner(sub, want_image=True, want_roi=True)
    image_h, roi_h, stats_h = runner.alloc_host_outputs()
    steps = 5 if not args.e2e_timepoints else max(1, min(args.steps, 6))   # assays pipelined back to back

    def step():
        runner.run(tiles_host, image_h, roi_h, stats_h)

    step()
    runner.synchronize()
    barrier()
    # Wall clock around fully synchronised ends: the region contains every H2D copy, kernel and D2H
    # copy of `steps` assays (the D2H of assay k overlaps the H2D of assay k+1 -- PCIe full duplex).
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    runner.synchronize()
    elapsed = time.perf_counter() - t0
    barrier()
    ms = torch.tensor([elapsed * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
    sec = float(ms.item()) / 1e3
    roi_px = m * c * t_e2e * length * length
    return {
        "value": roi_px * world * steps / sec, "unit": UNIT, "h2d_bytes_per_step": int(runner.h2d_bytes),
        "d2h_bytes_per_step": int(runner.d2h_bytes), "timepoints": t_e2e, "steps": steps,
        "ms_per_step": sec * 1e3 / steps,
        "h2d_plus_d2h_GBps_per_gpu": (runner.h2d_bytes + runner.d2h_bytes) * steps / sec / 1e9,
        "outputs_copied_back": "stitched image + roi + summaries",
        "numa_bound_cpus": len(os.sched_getaffinity(0)),
        "pipelining": "D2H of assay k overlaps H2D of assay k+1; timed with host clock around synchronised ends",
    }


class StdoutToStderr:
    """Everything written to fd 1 while active goes to stderr (NCCL prints its version banner to
    stdout); the JSON line is printed after restoring, so stdout carries exactly one line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--timepoints", type=int, default=None, help="timepoints per rank (default 50 = config 3)")
    ap.add_argument("--e2e-timepoints", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fused-gather", action="store_true",
                    help="N>1: gather the summaries with NCCL all_gather instead of peer stores from the kernel")
    ap.add_argument("--config", default="c3", choices=["c3", "c5"],
                    help="c3 = BASELINE config 3 (default, the metric's workload); c5 = one rank's shard of config 5")
    args = ap.parse_args()
    cfg = dict(C3 if args.config == "c3" else C5)
    if args.timepoints:
        cfg["t"] = args.timepoints
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        with StdoutToStderr():
            line = run_b200_arm(args, cfg)
        if line is not None:
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

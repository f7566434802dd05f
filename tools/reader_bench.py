"""TIFF staging throughput, run on the GPU box:
  1. native page reads (mgb_tiff_read_files) of 2048x2048 u16 single-page files from the page
     cache into PINNED memory, by thread count;
  2. the same files decoded by libtiff through cv2.imread (what a per-page Python reader costs);
  3. files -> pinned ring -> HBM -> flat-field max + fused flat-field/stitch + gather/stats
     (pipeline.ChunkStager fed by reader.TiffTiles.blocks), end to end.
Writes one JSON object to stdout."""
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from magnify_b200 import pipeline, reader, synth
from tiffgen import write_tiff

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
root = tempfile.mkdtemp(prefix="mgb_tiff_", dir=os.environ.get("MGB_TIFF_DIR", "/tmp"))
dev = torch.device("cuda:0")
out = {"tile": [2048, 2048], "dtype": "u16", "cores": os.cpu_count()}
try:
    case = synth.chip_case(c=2, t=T, r=4, cc=4, device=dev)        # C2 geometry, T timepoints
    tiles_np = case.tiles.cpu().numpy()
    c, t, r, cc, h, w = tiles_np.shape
    t0 = time.perf_counter()
    for idx in np.ndindex(c, t, r, cc):
        write_tiff(os.path.join(root, f"chip_ch{idx[0]}_202401{idx[1] + 1:02d}-000000_{idx[2]}_{idx[3]}.tif"),
                   [tiles_np[idx]], rows_per_strip=64)
    out["files"] = c * t * r * cc
    out["write_s"] = time.perf_counter() - t0
    (xp,) = list(reader.Reader(threads=16)(os.path.join(root, "chip_(channel)_(time)_(row)_(col).tif")))
    tiles = xp["tile"].data
    paths = tiles.filenames
    nbytes = tiles.nbytes
    pinned = torch.empty(tiles.shape, dtype=torch.uint16, pin_memory=True)
    flat = pinned.numpy().reshape((-1, h, w))
    reader.read_files(paths, flat, threads=16)                      # warm the page cache
    assert np.array_equal(pinned.numpy(), tiles_np)
    out["native_GBps_by_threads"] = {}
    for threads in (1, 2, 4, 8, 16, 32):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            reader.read_files(paths, flat, threads=threads)
            best = min(best, time.perf_counter() - t0)
        out["native_GBps_by_threads"][threads] = nbytes / best / 1e9
    import cv2

    n_cv = min(len(paths), 64)
    t0 = time.perf_counter()
    for p in paths[:n_cv]:
        img = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    out["cv2_libtiff_GBps_1thread"] = n_cv * h * w * 2 / (time.perf_counter() - t0) / 1e9
    assert np.array_equal(img, flat[n_cv - 1])

    # end to end: files -> pinned ring -> HBM -> kernels -> host outputs
    plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev)
    plan.set_chip_markers(case.x, case.y, case.fg_radius, case.chamber_radius, case.max_button_radius)
    ref = plan.run_device(case.tiles)
    stats_ref = ref.stats.cpu()
    runner = pipeline.HostStagedRunner(plan)
    image_h, roi_h, stats_h = runner.alloc_host_outputs()
    res = {}
    for depth, threads in ((2, 4), (4, 8), (4, 16), (6, 32)):
        tiles.threads = threads
        stager = pipeline.ChunkStager(runner, depth=depth, threads=threads)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            stager.feed(tiles.blocks())
            runner.finish(image_h, roi_h, stats_h)
            runner.synchronize()
            best = min(best, time.perf_counter() - t0)
        stager.close()
        assert torch.equal(stats_h, stats_ref)
        roi_px = roi_h.numel()
        res[f"depth{depth}_threads{threads}"] = {"s": best, "tile_GBps": nbytes / best / 1e9,
                                                 "roi_px_per_s": roi_px / best}
    out["files_to_results"] = res
    out["tile_bytes"] = nbytes
finally:
    shutil.rmtree(root, ignore_errors=True)
print(json.dumps(out))

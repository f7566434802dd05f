"""Flat-field + stitch throughput versus the overlap (alignment of the kept window): which part of
the gap between config 3 (overlap 102: clip 51, kept width 1946) and the contiguous case comes from
the unaligned crop."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from magnify_b200 import ops, synth

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
case = synth.chip_case(c=4, t=T, seed=0, device=dev)
px = case.tiles.numel()
for overlap in (102, 96, 104, 112, 128, 100, 0):
    plan = ops.FlatFieldPlan(case.tiles.shape, case.flat, case.dark, device=dev)
    maxima = ops.flatfield_maxima(case.tiles, plan).clone()
    image = ops.alloc_image(ops.stitched_shape(case.tiles.shape, overlap), torch.uint16, dev)
    fn = lambda: ops.flatfield_stitch(case.tiles, overlap=overlap, plan=plan, maxima=maxima, out=image)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    kept = 2048 - overlap
    phi = (kept / 2048) ** 2
    real = (2 * kept / 2048 + 2 * phi) * px          # rows outside the kept band are never read
    print(f"overlap {overlap:4d} kept {kept:5d} (w%8={kept % 8}, clip%8={(overlap // 2) % 8}): {ms:7.3f} ms  "
          f"algorithmic {(2 + 2 * phi) * px / ms / 1e6:6.0f} GB/s  touched-rows {real / ms / 1e6:6.0f} GB/s")

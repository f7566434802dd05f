"""Timing of the setup / optional kernels (run on the GPU box): medians, bead label raster, masks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magnify_b200 import ops

dev = torch.device("cuda:0")
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

# medians on a C3-like roi at T=8: (1792, 4, 8, 72, 72)
m, c, t, L = 1792, 4, 8, 72
roi = torch.randint(380, 460, (m, c, t, L, L), dtype=torch.int32, device=dev).to(torch.uint16)
yy, xx = torch.meshgrid(torch.arange(L, device=dev), torch.arange(L, device=dev), indexing="ij")
d2 = (yy - 36) ** 2 + (xx - 36) ** 2
fg = (d2 <= 14 ** 2).to(torch.uint8)[None, None].expand(m, 1, L, L).contiguous()
bg = ((d2 <= 30 ** 2) & (d2 > 15 ** 2)).to(torch.uint8)[None, None].expand(m, 1, L, L).contiguous()
for name, mask in (("fg", fg), ("bg", bg)):
    ms = timeit(lambda: ops.roi_median(roi, mask))
    print(f"roi_median {name}: {ms:.3f} ms for {m*c*t} ROIs of {L}^2 -> {roi.numel()/ms/1e6:.1f} Gpx/s ({2*roi.numel()/ms/1e6:.0f} GB/s read)")
ms = timeit(lambda: ops.roi_stats(roi, fg, bg))
print(f"roi_stats: {ms:.3f} ms -> {2*roi.numel()/ms/1e6:.0f} GB/s read")
# bead screen setup: 1e5 beads on 20480^2
rng = np.random.default_rng(0)
n, size, Lb = 100_000, 20480, 50
beads = torch.from_numpy(np.stack([rng.integers(0, size, n), rng.integers(0, size, n), rng.integers(4, 13, n)], 1).astype(np.int32)).to(dev)
ms = timeit(lambda: ops.bead_labels(beads, size, size), n=3)
print(f"bead_labels (memset + raster): {ms:.3f} ms for {n} beads on {size}^2 ({4*size*size/ms/1e6:.0f} GB/s of label image)")
labels = ops.bead_labels(beads, size, size)
x = beads[:, 1:2].to(torch.float64).contiguous(); y = beads[:, 0:1].to(torch.float64).contiguous()
ms = timeit(lambda: ops.bounding_boxes(x, y, Lb, size, size))
print(f"bounding_boxes: {ms:.4f} ms for {n}")
boxes = ops.bounding_boxes(x, y, Lb, size, size)[:, 0].contiguous()
ms = timeit(lambda: ops.bead_masks(labels, boxes, Lb))
print(f"bead_masks: {ms:.3f} ms ({n*Lb*Lb*6/ms/1e6:.0f} GB/s: 4 B label read + 2 B masks written per px)")
rel = torch.randint(20, 50, (1792, 2), dtype=torch.int32, device=dev); rad = torch.full((1792,), 14, dtype=torch.int32, device=dev)
ms = timeit(lambda: ops.chip_masks(rel, rad, 15, 30, 72))
print(f"chip_masks: {ms:.4f} ms for 1792 x 72^2")

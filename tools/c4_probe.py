"""Config 4 (stitch + flat-field sweep, 10x10 tiles of 2048^2, overlap 102) throughput probe."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from magnify_b200 import ops, synth
T = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
shape = (4, T, 10, 10, 2048, 2048)
tiles = torch.empty(shape, dtype=torch.uint16, device=dev)
g = torch.Generator(device=dev); g.manual_seed(0)
for c in range(4):
    for t in range(T):
        tiles[c, t] = torch.randint(0, 65536, (10, 10, 2048, 2048), dtype=torch.int32, device=dev, generator=g).to(torch.uint16)
flat, dark = synth.smooth_flat_dark(2048, 2048)
plan = ops.FlatFieldPlan(shape, flat, dark, device=dev)
image = ops.alloc_image(ops.stitched_shape(shape, 102), torch.uint16, dev)
print("image shape", tuple(image.shape), "pitch", ops.image_pitch(image), "contiguous", image.is_contiguous())
def timeit(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
px = tiles.numel(); phi = (19460 * 19460) / (100 * 2048 * 2048)
ms = timeit(lambda: ops.flatfield_maxima(tiles, plan))
print(f"max pass: {ms:.3f} ms {2 * px / ms / 1e6:.0f} GB/s")
maxima = ops.flatfield_maxima(tiles, plan).clone()
ms2 = timeit(lambda: ops.flatfield_stitch(tiles, overlap=102, plan=plan, maxima=maxima, out=image))
print(f"flat-field + stitch: {ms2:.3f} ms {(2 + 2 * phi) * px / ms2 / 1e6:.0f} GB/s")
ms3 = timeit(lambda: ops.stitch(tiles, 102, out=image))
print(f"plain stitch: {ms3:.3f} ms {4 * phi * px / ms3 / 1e6:.0f} GB/s")
print(f"C4 total (4 ch x 20 t) extrapolated: {(ms + ms2) * 20 / T:.1f} ms for 194.8 GB algorithmic -> {194.8 / ((ms + ms2) * 20 / T) * 1e3:.0f} GB/s, {px * 20 / T / ((ms + ms2) * 20 / T) / 1e6:.1f} Gtile-px/s")

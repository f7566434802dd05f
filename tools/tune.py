"""Kernel tuning probe (run on the GPU box): times the flat-field+stitch variants and the gather."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from magnify_b200 import _lib, ops, pipeline, synth

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
case = synth.chip_case(c=4, t=T, seed=0, device=dev)
plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev)
plan.set_chip_markers(case.x, case.y, case.fg_radius, case.chamber_radius, case.max_button_radius)
image = torch.empty(plan.image_shape, dtype=torch.uint16, device=dev)
maxima = ops.flatfield_maxima(case.tiles, plan.ff).clone()
lib = _lib.load()

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

px = case.tiles.numel()
phi = (plan.image_shape[-1] * plan.image_shape[-2]) / (4 * 4 * 2048 * 2048)
ref = None
for v in (0, 1, 2):
    lib.mgb_set_stitch_variant(v)
    ms = timeit(lambda: ops.flatfield_stitch(case.tiles, overlap=case.overlap, plan=plan.ff, maxima=maxima, out=image))
    chk = int(image.view(torch.int16).to(torch.int64).sum().item())
    ref = chk if ref is None else ref
    print(f"stitch variant {v}: {ms:.3f} ms  {(2 + 2 * phi) * px / ms / 1e6:.0f} GB/s  checksum_equal={chk == ref}")
for v in (0, 1, 2):
    lib.mgb_set_stitch_variant(v)
    ms = timeit(lambda: ops.stitch(case.tiles, case.overlap, out=image))
    print(f"plain stitch variant {v}: {ms:.3f} ms  {(4 * phi) * px / ms / 1e6:.0f} GB/s")
m, c, t, L = plan.boxes.shape[0], 4, T, case.roi_length
roi = torch.empty((m, c, t, L, L), dtype=torch.uint16, device=dev)
stats = torch.empty((m, c, t, 8), dtype=torch.float64, device=dev)
for gran in (0,):
  for name, tma, loader in (("staged cp.async", 1, 1), ("staged TMA", 1, 0), ("plain LSU", 0, 1)):
      lib.mgb_set_tma_enabled(tma)
      lib.mgb_set_gather_loader(loader)
      ms = timeit(lambda: ops.roi_gather_stats(image, plan.boxes, plan.fg, plan.bg, L, mask_t=plan.mask_t, out_roi=roi, out_stats=stats))
      chk = float(stats[..., 2:4].sum().item())
      print(f"gather+stats {name}: {ms:.3f} ms  {4 * roi.numel() / ms / 1e6:.0f} GB/s  (stats checksum {chk:.0f})")
      ms = timeit(lambda: ops.roi_gather(image, plan.boxes, L, out=roi))
      print(f"gather only  {name}: {ms:.3f} ms  {4 * roi.numel() / ms / 1e6:.0f} GB/s")
      ms = timeit(lambda: ops.roi_gather_stats(image, plan.boxes, plan.fg, plan.bg, L, mask_t=plan.mask_t, want_roi=False, out_stats=stats))
      print(f"stats only   {name}: {ms:.3f} ms  {2 * roi.numel() / ms / 1e6:.0f} GB/s (read only)")
lib.mgb_set_tma_enabled(1); lib.mgb_set_gather_loader(1)
ms = timeit(lambda: ops.flatfield_maxima(case.tiles, plan.ff))
print(f"flatfield max pass: {ms:.3f} ms {2 * px / ms / 1e6:.0f} GB/s")
# pure elementwise flat-field (every tile its own image, overlap 0, S = 0, contiguous)
lib.mgb_set_stitch_variant(0)
tiles_img = case.tiles.view(4, T * 16, 1, 1, 2048, 2048)
out_t = torch.empty_like(case.tiles).view(4, T * 16, 2048, 2048)
ms = timeit(lambda: ops.flatfield_stitch(tiles_img, overlap=0, plan=plan.ff, maxima=maxima, out=out_t))
print(f"flat-field only (contiguous): {ms:.3f} ms  {4 * px / ms / 1e6:.0f} GB/s")
ms = timeit(lambda: ops.stitch(tiles_img, 0, out=out_t))
print(f"plain copy through stitch kernel (contiguous): {ms:.3f} ms  {4 * px / ms / 1e6:.0f} GB/s")
a = case.tiles.view(torch.int16); b = torch.empty_like(a)
ms = timeit(lambda: b.copy_(a))
print(f"torch copy_: {ms:.3f} ms  {4 * px / ms / 1e6:.0f} GB/s")
# even overlap 96 (aligned phases, S = 0 for all) for comparison
for ov in (96, 104, 100):
    img2 = torch.empty(ops.stitched_shape(case.tiles.shape, ov), dtype=torch.uint16, device=dev)
    phi2 = (img2.shape[-1] * img2.shape[-2]) / (16 * 2048 * 2048)
    ms = timeit(lambda: ops.stitch(case.tiles, ov, out=img2))
    print(f"plain stitch overlap {ov}: {ms:.3f} ms  {(4 * phi2) * px / ms / 1e6:.0f} GB/s")

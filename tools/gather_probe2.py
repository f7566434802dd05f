"""Gather + reductions variants on a config-3 shaped stack (run on the GPU box):
    python tools/gather_probe2.py [T] [--c5] [--variants=cta:12,cta:6,...]
Times, with CUDA events on the launching stream, the fused gather with and without medians, in
both work layouts (CTA per marker / warp per marker), with and without the crops, against the
round-1 arrangement (dp2a sums in the gather + two stand-alone median passes).  One JSON line
per variant."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from magnify_b200 import _lib, ops, pipeline, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
c5 = "--c5" in sys.argv
dev = torch.device("cuda:0")
lib = _lib.load()
if c5:
    case = synth.bead_case(c=4, t=T, r=10, cc=10, h=2048, w=2048, overlap=0, n_beads=100000, min_radius=4,
                           max_radius=12, roi_length=50, seed=0, device=dev)
    plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev)
    plan.set_bead_markers(case.beads)
else:
    case = synth.chip_case(c=4, t=T, seed=0, device=dev)
    plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev)
    plan.set_chip_markers(case.x, case.y, case.fg_radius, case.chamber_radius, case.max_button_radius)
image = plan.stitched(case.tiles)
del case.tiles
m, c, t, L = plan.boxes.shape[0], 4, T, plan.roi_length
roi = torch.empty((m, c, t, L, L), dtype=torch.uint16, device=dev)
stats = torch.empty((m, c, t, ops.NSTATS), dtype=torch.float64, device=dev)
print(json.dumps({"markers": m, "L": L, "T": T, "mask_counts": plan.mask_counts}), flush=True)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def gather(medians=True, want_roi=True, counts=plan.mask_counts):
    return ops.roi_gather_stats(image, plan.boxes, plan.fg, plan.bg, L, mask_t=plan.mask_t, want_roi=want_roi,
                                out_roi=roi if want_roi else None, out_stats=stats, order=plan.order,
                                medians=medians, mask_counts=counts)


ref = None
alg = 4.0 * roi.numel()
variants = (("auto", "auto"), ("cta", "8"), ("cta", "12"), ("warp", "8"), ("warp", "12"))
for arg in sys.argv:
    if arg.startswith("--variants="):           # e.g. --variants=cta:12,cta:6
        variants = tuple(tuple(v.split(":")) for v in arg.split("=", 1)[1].split(","))
for layout, algo in variants:
    if layout == "auto":
        os.environ.pop("MGB_GATHER_LAYOUT", None)
    else:
        os.environ["MGB_GATHER_LAYOUT"] = "1" if layout == "warp" else "0"
    if algo == "auto":
        os.environ.pop("MGB_GATHER_WARPS", None)
    else:
        os.environ["MGB_GATHER_WARPS"] = algo
    for name, kw in (("lists+medians", dict(medians=True)), ("lists sums only", dict(medians=False)),
                     ("lists+medians, no crops", dict(medians=True, want_roi=False))):
        ms = timeit(lambda: gather(**kw))
        out = {"variant": name, "layout": layout, "warps": algo, "ms": round(ms, 4), "alg_GBps": round(alg / ms / 1e6, 1)}
        if kw.get("medians", True):
            chk = torch.nan_to_num(stats, nan=-1.0).sum().item()
            ref = chk if ref is None else ref
            out["checksum_equal"] = chk == ref
        print(json.dumps(out), flush=True)
os.environ.pop("MGB_GATHER_LAYOUT", None)
os.environ.pop("MGB_GATHER_WARPS", None)
# round-1 arrangement: dp2a sums in the gather (no lists: unknown mask counts), medians as two passes
ms_g = timeit(lambda: gather(medians=False, counts=(-1, -1)))
ms_m = timeit(lambda: (ops.roi_median(roi, plan.fg, mask_t=plan.mask_t), ops.roi_median(roi, plan.bg, mask_t=plan.mask_t)))
gather(medians=True, counts=(-1, -1))
chk = torch.nan_to_num(stats, nan=-1.0).sum().item()
print(json.dumps({"variant": "dp2a sums + 2 median passes", "gather_ms": round(ms_g, 4), "medians_ms": round(ms_m, 4),
                  "total_ms": round(ms_g + ms_m, 4), "checksum_equal": chk == ref}), flush=True)
ms = timeit(lambda: ops.roi_gather(image, plan.boxes, L, out=roi, order=plan.order))
print(json.dumps({"variant": "crops only", "ms": round(ms, 4), "alg_GBps": round(alg / ms / 1e6, 1)}), flush=True)

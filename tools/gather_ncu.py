"""ncu target: a few launches of the fused gather (lists + medians) on a config-3 shaped stack.
    python tools/gather_ncu.py [T] [--no-medians] [--c5]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from magnify_b200 import ops, pipeline, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
dev = torch.device("cuda:0")
if "--c5" in sys.argv:
    case = synth.bead_case(c=4, t=T, r=10, cc=10, h=2048, w=2048, overlap=0, n_beads=100000, min_radius=4,
                           max_radius=12, roi_length=50, seed=0, device=dev)
    plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev)
    plan.set_bead_markers(case.beads)
else:
    case = synth.chip_case(c=4, t=T, seed=0, device=dev)
    plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev)
    plan.set_chip_markers(case.x, case.y, case.fg_radius, case.chamber_radius, case.max_button_radius)
image = plan.stitched(case.tiles)
m, L = plan.boxes.shape[0], plan.roi_length
roi = torch.empty((m, 4, T, L, L), dtype=torch.uint16, device=dev)
stats = torch.empty((m, 4, T, ops.NSTATS), dtype=torch.float64, device=dev)
for _ in range(3):
    ops.roi_gather_stats(image, plan.boxes, plan.fg, plan.bg, L, mask_t=plan.mask_t, out_roi=roi, out_stats=stats,
                         order=plan.order, medians="--no-medians" not in sys.argv, mask_counts=plan.mask_counts)
torch.cuda.synchronize()
print("done")

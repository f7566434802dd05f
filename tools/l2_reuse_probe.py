"""VERDICT r1 item 3, measured: can the gather's window reads hit L2 instead of DRAM when it runs
right behind the stitch that wrote the image?  (run under ncu on the GPU box)

    python tools/l2_reuse_probe.py [T] [mode]
    mode = whole   stitch of the whole stack, then one gather (the shipped arrangement)
           plane   per (channel, time) image: stitch that plane, then gather its windows
           persist like `plane`, with the plane as a persisting-L2 access-policy window
           time    per timepoint (4 planes = 484 MB): stitch, then gather
           band    per tile row of a plane (30 MB of image): stitch that band, then gather the markers inside it

The Python loop's launch overhead makes the wall time of the per-plane modes meaningless; what
the experiment measures is `dram__bytes_read.sum` / `lts__t_sector_hit_rate.pct` of the gather
launches (ncu --metrics ..., summed per kernel by tools/ncu_sum.py)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from magnify_b200 import _lib, ops, pipeline, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4
mode = sys.argv[2] if len(sys.argv) > 2 else "whole"
dev = torch.device("cuda:0")
lib = _lib.load()
case = synth.chip_case(c=4, t=T, seed=0, device=dev)
plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev)
plan.set_chip_markers(case.x, case.y, case.fg_radius, case.chamber_radius, case.max_button_radius)
maxima = ops.flatfield_maxima(case.tiles, plan.ff).clone()
c, t = 4, T
m, L = plan.boxes.shape[0], plan.roi_length
image = ops.alloc_image(plan.image_shape, torch.uint16, dev)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def gather(img, boxes, mask_t):
    return ops.roi_gather_stats(img, boxes, plan.fg, plan.bg, L, mask_t=mask_t, order=plan.order, medians=True,
                                mask_counts=plan.mask_counts)


if mode == "whole":
    for _ in range(2):
        ops.flatfield_stitch(case.tiles, overlap=case.overlap, plan=plan.ff, maxima=maxima, out=image)
        gather(image, plan.boxes, plan.mask_t)
elif mode in ("plane", "persist"):
    boxes_t = [plan.boxes[:, ti:ti + 1].contiguous() for ti in range(t)]
    mask_1 = plan.mask_t[:1].contiguous()
    for _ in range(2):
        for ti in range(t):
            for ci in range(c):
                out = image[ci:ci + 1, ti:ti + 1]
                if mode == "persist":
                    _lib.call("mgb_l2_persist", stream, ctypes.c_void_p(out.data_ptr()), out.numel() * 2, ctypes.c_float(0.6))
                # per-channel flat-field tables: this case has one table for all channels (K = 1)
                ops.flatfield_stitch(case.tiles[ci:ci + 1, ti:ti + 1], overlap=case.overlap, plan=plan.ff, maxima=maxima, out=out)
                gather(out, boxes_t[ti], mask_1)
        if mode == "persist":
            _lib.call("mgb_l2_persist", stream, None, 0, ctypes.c_float(0.0))
elif mode == "time":
    boxes_t = [plan.boxes[:, ti:ti + 1].contiguous() for ti in range(t)]
    mask_1 = plan.mask_t[:1].contiguous()
    for _ in range(2):
        for ti in range(t):
            out = image[:, ti:ti + 1]
            tiles_t = case.tiles[:, ti:ti + 1].contiguous()
            img_t = ops.alloc_image((c, 1) + tuple(plan.image_shape[2:]), torch.uint16, dev)
            ops.flatfield_stitch(tiles_t, overlap=case.overlap, plan=plan.ff, maxima=maxima, out=img_t)
            gather(img_t, boxes_t[ti], mask_1)
elif mode == "band":
    r, kh = case.tiles.shape[2], case.tiles.shape[4] - case.overlap
    boxes_cpu = plan.boxes.cpu()
    mask_1 = plan.mask_t[:1].contiguous()
    groups = []
    for ri in range(r):
        inside = ((boxes_cpu[:, 0, 0] >= ri * kh) & (boxes_cpu[:, 0, 0] + L <= (ri + 1) * kh)).nonzero().flatten().to(dev)
        groups.append((inside, plan.fg[inside].contiguous(), plan.bg[inside].contiguous()))
    print("markers per band:", [len(g[0]) for g in groups])
    for _ in range(2):
        for ti in range(t):
            for ci in range(c):
                for ri in range(r):
                    out = image[ci:ci + 1, ti:ti + 1, ri * kh:(ri + 1) * kh]
                    ops.flatfield_stitch(case.tiles[ci:ci + 1, ti:ti + 1, ri:ri + 1], overlap=case.overlap, plan=plan.ff,
                                         maxima=maxima, out=out)
                    idx, fg, bg = groups[ri]
                    bx = plan.boxes[idx, ti:ti + 1].clone()
                    bx[..., 0] -= ri * kh
                    ops.roi_gather_stats(out, bx.contiguous(), fg, bg, L, mask_t=mask_1, medians=True,
                                         mask_counts=plan.mask_counts)
torch.cuda.synchronize()
print("done", mode)

"""ncu target: one gather launch per loader on a C3-shaped image (T timepoints)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from magnify_b200 import _lib, ops
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
lib = _lib.load()
him, wim = 7784, 7784
image = torch.randint(0, 65536, (4, T, him, wim), dtype=torch.int32, device=dev).to(torch.uint16)
case_x, case_y = None, None
import numpy as np
rng = np.random.default_rng(0)
rows, cols = 56, 32
cy = 400 + np.arange(rows)[:, None] * 126.1 + rng.uniform(-2, 2, (rows, cols))
cx = 280 + np.arange(cols)[None, :] * 232.9 + rng.uniform(-2, 2, (rows, cols))
x = torch.from_numpy(np.repeat(cx.reshape(-1, 1), T, 1)).to(dev).contiguous()
y = torch.from_numpy(np.repeat(cy.reshape(-1, 1), T, 1)).to(dev).contiguous()
boxes, rel = ops.bounding_boxes(x, y, 72, wim, him, want_rel=True)
rad = torch.full((rows * cols,), 14, dtype=torch.int32, device=dev)
fg, bg = ops.chip_masks(rel[:, 0].contiguous(), rad, 15, 30, 72)
fg, bg = fg[:, None].contiguous(), bg[:, None].contiguous()
roi = torch.empty((rows * cols, 4, T, 72, 72), dtype=torch.uint16, device=dev)
stats = torch.empty((rows * cols, 4, T, 8), dtype=torch.float64, device=dev)
for gran in (None,):
    for name, tma, loader in (("cpasync", 1, 1), ("tma", 1, 0), ("plain", 0, 1)):
        lib.mgb_set_tma_enabled(tma); lib.mgb_set_gather_loader(loader)
        for _ in range(2):
            ops.roi_gather_stats(image, boxes, fg, bg, 72, out_roi=roi, out_stats=stats)
        torch.cuda.synchronize()
print("done")

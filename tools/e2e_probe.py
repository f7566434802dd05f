"""Where the time of the component path goes (run on the GPU box): python tools/e2e_probe.py [T]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from magnify_b200 import components, devarray, synth  # noqa: E402
from magnify_b200.dataset import Dataset  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
case = synth.chip_case(c=4, t=T, seed=0, device=dev)
host = torch.empty(tuple(case.tiles.shape), dtype=torch.uint16, pin_memory=True)
host.copy_(case.tiles)
torch.cuda.synchronize()
tiles_np = host.numpy()
print("numpy view pinned:", components._is_pinned(tiles_np), "tensor pinned:", host.is_pinned())
rows, cols = case.grid
x = case.x.reshape(rows, cols, T)
y = case.y.reshape(rows, cols, T)
rad = case.fg_radius[:, 0].reshape(rows, cols)
coords = {"tag": (("mark_row", "mark_col"), np.full((rows, cols), "default", dtype="<U200")),
          "valid": (("mark_row", "mark_col", "time"), np.ones((rows, cols, T), dtype=bool))}
pipe = [("flatfield_correct", components.make_flatfield_correct(case.flat, case.dark)),
        ("stitch", components.make_stitch(102)),
        ("find_buttons", components.ButtonFinder(126.1, 232.9, 16, 30, 60, centers=lambda xp, ts: (x[..., ts], y[..., ts], rad))),
        ("quantify", components.make_quantify())]
import functools  # noqa: E402

from magnify_b200 import ops  # noqa: E402


def timed(mod, name):
    fn = getattr(mod, name)

    @functools.wraps(fn)
    def wrapper(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        print(f"      {name:24s} {1e3 * (time.perf_counter() - t0):8.1f} ms", flush=True)
        return out

    setattr(mod, name, wrapper)


for mod, name in ((components, "_stage_tiles"), (components, "_emit"), (ops, "flatfield_stitch"), (ops, "FlatFieldPlan"),
                  (ops, "roi_gather_stats"), (ops, "bounding_boxes"), (ops, "chip_masks"), (components, "_image_on_device"),
                  (components, "_emit_markers")):
    timed(mod, name)
_empty = torch.empty


def timed_empty(*a, **k):
    t0 = time.perf_counter()
    out = _empty(*a, **k)
    dt = time.perf_counter() - t0
    if dt > 5e-3:
        print(f"      torch.empty{tuple(out.shape)} pin={k.get('pin_memory', False)} dev={out.device} {1e3 * dt:8.1f} ms", flush=True)
    return out


torch.empty = timed_empty
for rep in range(3):
    assay = Dataset({"tile": (components.TILE_DIMS, tiles_np)}, coords=coords)
    t_all = time.perf_counter()
    for name, comp in pipe:
        t0 = time.perf_counter()
        assay = comp(assay)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"rep {rep} {name:18s} call {1e3 * (t1 - t0):8.1f} ms  +sync {1e3 * (t2 - t1):8.1f} ms", flush=True)
    for name in ("image", "roi", "fg_mean"):
        t0 = time.perf_counter()
        v = assay[name].values
        print(f"rep {rep} read {name:13s} {1e3 * (time.perf_counter() - t0):8.1f} ms  {v.nbytes / 1e9:.2f} GB", flush=True)
    print(f"rep {rep} total {1e3 * (time.perf_counter() - t_all):8.1f} ms", flush=True)
    del assay, v

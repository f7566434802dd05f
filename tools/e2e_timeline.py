"""Timeline of the component path with assays back to back (run on the GPU box):
    python tools/e2e_timeline.py [T]
For every assay: host time when run_pipe returns / when the previous result has been read, and device
times at which the upload stream, the compute stream and the download stream reached the end of what
that assay queued."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from magnify_b200 import components, devarray, synth  # noqa: E402
from magnify_b200.dataset import Dataset  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda:0")
case = synth.chip_case(c=4, t=T, seed=0, device=dev)
host = torch.empty(tuple(case.tiles.shape), dtype=torch.uint16, pin_memory=True)
host.copy_(case.tiles)
torch.cuda.synchronize()
tiles_np = host.numpy()
del case.tiles
rows, cols = case.grid
x, y = case.x.reshape(rows, cols, T), case.y.reshape(rows, cols, T)
rad = case.fg_radius[:, 0].reshape(rows, cols)
coords = {"tag": (("mark_row", "mark_col"), np.full((rows, cols), "default", dtype="<U200")),
          "valid": (("mark_row", "mark_col", "time"), np.ones((rows, cols, T), dtype=bool))}
pipe = [components.make_flatfield_correct(case.flat, case.dark), components.make_stitch(102),
        components.ButtonFinder(126.1, 232.9, 16, 30, 60, centers=lambda xp, ts: (x[..., ts], y[..., ts], rad)),
        components.make_quantify()]
streams = devarray.Streams.of(dev)
compute = torch.cuda.current_stream(dev)


def run_pipe():
    assay = Dataset({"tile": (components.TILE_DIMS, tiles_np)}, coords=coords)
    for comp in pipe:
        assay = comp(assay)
    return assay


def read_back(assay):
    for name in ("image", "roi", "fg_mean"):
        assay[name].values


prev = None
for _ in range(3):
    cur = run_pipe()
    if prev is not None:
        read_back(prev)
    prev = cur
read_back(prev)
del prev, cur
torch.cuda.synchronize()
start = torch.cuda.Event(enable_timing=True)
start.record(compute)
t0 = time.perf_counter()
marks = []
prev = None
for k in range(5):
    cur = run_pipe()
    t_ret = time.perf_counter() - t0
    evs = []
    for st in (streams.h2d, compute, streams.d2h):
        e = torch.cuda.Event(enable_timing=True)
        e.record(st)
        evs.append(e)
    if prev is not None:
        read_back(prev)
    t_read = time.perf_counter() - t0
    marks.append((t_ret, t_read, evs))
    prev = cur
read_back(prev)
torch.cuda.synchronize()
total = time.perf_counter() - t0
for k, (t_ret, t_read, evs) in enumerate(marks):
    print(f"assay {k}: run_pipe returned {1e3 * t_ret:7.1f} ms, previous read {1e3 * t_read:7.1f} ms | "
          f"h2d done {start.elapsed_time(evs[0]):7.1f}  compute done {start.elapsed_time(evs[1]):7.1f}  "
          f"d2h done {start.elapsed_time(evs[2]):7.1f} ms")
print(f"total {1e3 * total:.1f} ms for 5 assays of T={T}; h2d {tiles_np.nbytes / 1e9:.1f} GB each")

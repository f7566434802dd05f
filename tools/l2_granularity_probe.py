"""Does the L2 fetch granularity hint change what the hot kernels cost?  (run on the GPU box)
    python tools/l2_granularity_probe.py [T]
The ROI gather reads 144-byte row fragments at arbitrary alignment and DRAM delivers them in
128-byte lines (1.9x, profiles/r02_l2_reuse.md).  cudaLimitMaxL2FetchGranularity (0x05) is the
runtime's hint for that granularity (32 / 64 / 128 bytes).  For each setting: the limit read back,
and CUDA-event times of the flat-field max pass, flat-field + stitch and the fused gather on a
config-3 stack; one JSON line per setting.  Under `ncu --metrics dram__bytes_read.sum,...` the same
script gives the bytes."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from magnify_b200 import ops, pipeline, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda:0")
torch.cuda.init()
torch.zeros(1, device=dev)
rt = ctypes.CDLL("libcudart.so.12")
LIMIT = 0x05  # cudaLimitMaxL2FetchGranularity


def get_limit():
    v = ctypes.c_size_t(0)
    rc = rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT)
    return int(v.value) if rc == 0 else f"error {rc}"


def set_limit(n):
    return rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(n))


case = synth.chip_case(c=4, t=T, seed=0, device=dev)
plan = pipeline.QuantifyPlan(case.tiles.shape, case.overlap, case.roi_length, case.flat, case.dark, device=dev)
plan.set_chip_markers(case.x, case.y, case.fg_radius, case.chamber_radius, case.max_button_radius)
m, c, L = plan.boxes.shape[0], 4, plan.roi_length
image = ops.alloc_image(plan.image_shape, torch.uint16, dev)
roi = torch.empty((m, c, T, L, L), dtype=torch.uint16, device=dev)
stats = torch.empty((m, c, T, ops.NSTATS), dtype=torch.float64, device=dev)


def measure(n=5):
    for _ in range(2):
        plan.run_device(case.tiles, want_roi=True, image_out=image, roi_out=roi, stats_out=stats)
    torch.cuda.synchronize()
    acc = {}
    for _ in range(n):
        rec = []
        plan.run_device(case.tiles, want_roi=True, image_out=image, roi_out=roi, stats_out=stats, record=rec)
        torch.cuda.synchronize()
        for name, a, b in rec:
            acc.setdefault(name, []).append(a.elapsed_time(b))
    return {k: round(sum(v) / len(v), 4) for k, v in acc.items()}


ref = None
print(json.dumps({"T": T, "limit_at_start": get_limit()}), flush=True)
for gran in (None, 32, 64, 128, None):
    rc = None
    if gran is not None:
        rc = set_limit(gran)
    ms = measure()
    chk = float(torch.nan_to_num(stats, nan=-1.0).sum().item())
    ref = chk if ref is None else ref
    print(json.dumps({"requested": gran, "rc": rc, "limit_now": get_limit(), "ms": ms, "checksum_equal": chk == ref}), flush=True)

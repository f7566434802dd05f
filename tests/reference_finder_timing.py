"""Timing of the reference's OWN utils.find_circles (loaded in place from /root/reference) on the
workloads of tools/finder_bench.py and tools/finder_bench_c5.py.  Test infrastructure: it runs only
in the build container (the GPU box has no /root/reference) and is the CPU side of the numbers in
DESIGN.md section 3.5.  Usage: python tests/reference_finder_timing.py [beads|c5]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import circles as oc
from oracle._refload import load_reference_utils

utils = load_reference_utils()
if utils is None:
    sys.exit("/root/reference is not available here")
which = sys.argv[1] if len(sys.argv) > 1 else "beads"
out = {"cores": os.cpu_count()}
if which == "beads":
    import finder_bench as fb

    img = oc.to_uint8(fb.bead_image())
    utils.find_circles(img[:256, :256], **dict(fb.BEADS, num_iter=1000), gui=None)          # numba warm-up
    t0 = time.perf_counter()
    c, s = utils.find_circles(img, **fb.BEADS, gui=None)
    out["reference_beads_s"] = time.perf_counter() - t0
    out["reference_beads_found"] = len(c)
    rois = fb.roi_batch()
    t0 = time.perf_counter()
    hits = 0
    for r in rois[:256]:
        c, s = utils.find_circles(oc.to_uint8(r), **fb.ROIS, gui=None)
        hits += len(c) > 0
    out["reference_rois_s_extrapolated_1792"] = (time.perf_counter() - t0) * 1792 / 256
    out["reference_rois_hit_fraction"] = hits / 256
else:
    import finder_bench_c5 as fb5

    img = fb5.image()
    utils.find_circles(img[:512, :512], **dict(fb5.ARGS, num_iter=1000), gui=None)
    t0 = time.perf_counter()
    c, s = utils.find_circles(img, **fb5.ARGS, gui=None)
    out.update(side=fb5.SIDE, beads=fb5.N_BEADS, num_iter=fb5.NUM_ITER, reference_s=time.perf_counter() - t0,
               reference_found=len(c))
print(json.dumps(out))

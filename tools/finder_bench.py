"""Circle finder timing.  On the GPU box: stage times of the GPU
finder for (a) a 2048^2 bead image with the reference's default 5e6 draws and (b) the per-chamber
refinement batch of config 2 (1792 crops of 72^2, 5e6 // 1792 draws each).  The reference's own
utils.find_circles is timed on the same images by tests/reference_finder_timing.py (build container)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from test_circles_host import synthetic_discs


def bead_image():
    rng = np.random.default_rng(5)
    discs = [(int(rng.integers(30, 2018)), int(rng.integers(30, 2018)), int(rng.integers(8, 26))) for _ in range(300)]
    return synthetic_discs(2048, 2048, discs, seed=3, noise=4)


def roi_batch():
    rng = np.random.default_rng(11)
    return np.stack([synthetic_discs(72, 72, [(int(rng.integers(26, 46)), int(rng.integers(26, 46)), int(rng.integers(8, 15)))],
                                     seed=k, noise=4, level=1500) for k in range(1792)])


BEADS = dict(low_edge_quantile=0.1, high_edge_quantile=0.9, grid_length=20, num_iter=5_000_000, min_radius=8, max_radius=25,
             min_roundness=0.3, min_dist=8)
ROIS = dict(low_edge_quantile=0.1, high_edge_quantile=1 - np.pi * 8 / 72**2, grid_length=20, num_iter=5_000_000 // 1792,
            min_radius=8, max_radius=15, min_roundness=0.2, min_dist=0)



def main():
    out = {}
    import torch

    from magnify_b200 import circles as mc

    dev = torch.device("cuda:0")

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best, r

    raw = torch.from_numpy(bead_image()).to(dev)
    u8 = mc.to_uint8(raw)
    stages = {}
    stages["to_uint8"], u8 = timed(lambda: mc.to_uint8(raw))
    stages["edges"], (edges, dx, dy) = timed(lambda: mc.find_edges(u8, 0.1, 0.9))
    stages["cell_lists"], lists = timed(lambda: mc.EdgeLists(edges, 20))
    stages["sample_dedupe"], (_, circles) = timed(lambda: mc.sample_circles(lists, BEADS["num_iter"], 8, 25, seed=1))
    stages["angles"], angle = timed(lambda: mc.gradient_angles(dx, dy))
    stages["score"], scores = timed(lambda: mc.score_circles(circles, edges, angle, 8, 25))
    found, sc = circles.cpu().numpy(), scores.cpu().numpy()
    t0 = time.perf_counter()
    sel = mc.select_circles(found[:, 1:], sc, 0.3, 8)
    stages["host_select_nms"] = time.perf_counter() - t0
    out["gpu_beads_stages_s"] = stages
    out["gpu_beads_unique_candidates"] = int(circles.shape[0])
    out["gpu_beads_total_s"], res = timed(lambda: mc.find_circles(u8, seed=1, **BEADS))
    out["gpu_beads_found"] = len(res[0])
    rois = torch.from_numpy(roi_batch()).to(dev)
    out["gpu_rois_total_s"], res = timed(lambda: mc.find_circles(mc.to_uint8(rois, batched=True), seed=2, **ROIS))
    out["gpu_rois_hit_fraction"] = float(np.mean([len(r[0]) > 0 for r in res]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

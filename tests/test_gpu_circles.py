"""GPU front end of the circle finder (csrc/circles.cu) against the reference's own OpenCV / NumPy
calls (oracle/circles.py): bit-exact at every stage."""
import numpy as np
import pytest
import torch

from oracle import circles as oc
from test_circles_host import synthetic_discs

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32, np.float64])
def test_to_uint8(cuda_device, dtype):
    from magnify_b200 import circles as mc

    rng = np.random.default_rng(1)
    cases = [rng.integers(3, 250, (37, 91)).astype(dtype), (rng.random((64, 64)) * 5000 + 100).astype(dtype),
             np.full((5, 7), 9).astype(dtype), np.zeros((0, 4), dtype)]
    if np.dtype(dtype).kind == "f":
        cases.append((rng.standard_normal((33, 17)) * 1e3).astype(dtype))
    for arr in cases:
        got = mc.to_uint8(dev(arr, cuda_device)).cpu().numpy()
        np.testing.assert_array_equal(got, oc.to_uint8(arr))


SHAPES = [(1, 1), (1, 9), (2, 2), (3, 5), (4, 33), (31, 32), (33, 65), (72, 72), (100, 100), (257, 511), (700, 900)]


@pytest.mark.parametrize("shape", SHAPES)
def test_gradients_quantiles_canny_match_opencv(cuda_device, shape):
    from magnify_b200 import circles as mc

    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    discs = [(int(rng.integers(0, h)), int(rng.integers(0, w)), int(rng.integers(2, 30))) for _ in range(6)]
    img = oc.to_uint8(synthetic_discs(h, w, discs, seed=h + w)) if h * w > 1 else np.array([[7]], np.uint8)
    for low_q, high_q in ((0.1, 0.9), (0.5, 0.99), (0.0, 1.0), (0.9, 0.1)):
        want = oc.edge_stages(img, low_q, high_q)
        dx, dy = mc.edge_gradients(dev(img, cuda_device))
        np.testing.assert_array_equal(dx.cpu().numpy(), want["dx"])         # float32 holding exact integers
        np.testing.assert_array_equal(dy.cpu().numpy(), want["dy"])
        low, high = mc.gradient_quantiles(dx, dy, (low_q, high_q))
        assert low == want["low"] and high == want["high"], (low, want["low"], high, want["high"])
        assert low.dtype == want["low"].dtype
        edges, sweeps = mc.canny(dx, dy, low, high, return_sweeps=True)
        np.testing.assert_array_equal(edges.cpu().numpy(), want["edges"])
        assert sweeps >= 1
        e2, _, _ = mc.find_edges(dev(img, cuda_device), low_q, high_q)
        np.testing.assert_array_equal(e2.cpu().numpy(), want["edges"])


def test_canny_long_spiral_needs_many_sweeps(cuda_device):
    """Hysteresis must follow a weak edge across many tiles from a single strong seed."""
    from magnify_b200 import circles as mc

    h = w = 400
    img = np.zeros((h, w), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    rr = np.hypot(yy - 200, xx - 200)
    theta = np.arctan2(yy - 200, xx - 200)
    spiral = np.abs(rr - (20 + 25 * (theta + np.pi) / (2 * np.pi) + 25 * np.round((rr - 20) / 25 - (theta + np.pi) / (2 * np.pi)))) < 2.5
    img[spiral] = 60
    img[195:205, 225:235][spiral[195:205, 225:235]] = 255          # one bright stretch
    want = oc.edge_stages(img, 0.5, 0.999)
    dx, dy = mc.edge_gradients(dev(img, cuda_device))
    edges, sweeps = mc.canny(dx, dy, want["low"], want["high"], return_sweeps=True)
    np.testing.assert_array_equal(edges.cpu().numpy(), want["edges"])
    assert want["edges"].sum() > 1000 and sweeps > 2


def test_full_size_tile(cuda_device):
    from magnify_b200 import circles as mc

    rng = np.random.default_rng(5)
    discs = [(int(rng.integers(0, 2048)), int(rng.integers(0, 2048)), int(rng.integers(8, 26))) for _ in range(300)]
    raw = synthetic_discs(2048, 2048, discs, seed=3)
    u8 = mc.to_uint8(dev(raw, cuda_device))
    np.testing.assert_array_equal(u8.cpu().numpy(), oc.to_uint8(raw))
    want = oc.edge_stages(oc.to_uint8(raw), 0.1, 0.9)
    edges, dx, dy = mc.find_edges(u8, 0.1, 0.9)
    np.testing.assert_array_equal(edges.cpu().numpy(), want["edges"])


def test_batched_rois_match_per_image_opencv(cuda_device):
    """A batch of 72 x 72 crops (the per-ROI refinement of find.py:339-360): every image scaled,
    thresholded and edge-detected on its own, in one set of launches."""
    from magnify_b200 import circles as mc

    rng = np.random.default_rng(11)
    b, length = 37, 72
    raw = np.stack([synthetic_discs(length, length, [(int(rng.integers(20, 52)), int(rng.integers(20, 52)),
                                                       int(rng.integers(6, 16)))], seed=k, noise=8 + k % 5,
                                    level=500 + 300 * (k % 7)) for k in range(b)])
    raw[5] = 1234                                                     # a constant crop -> all zeros
    u8 = mc.to_uint8(dev(raw, cuda_device), batched=True)
    want_u8 = np.stack([oc.to_uint8(r) for r in raw])
    np.testing.assert_array_equal(u8.cpu().numpy(), want_u8)
    high_q = np.float64(1 - np.pi * 8 / length**2)                    # find.py:345-347
    edges, dx, dy = mc.find_edges(u8, 0.1, high_q)
    lows_highs = mc.gradient_quantiles(dx, dy, (0.1, high_q))
    for k in range(b):
        want = oc.edge_stages(want_u8[k], 0.1, high_q)
        np.testing.assert_array_equal(dx[k].cpu().numpy(), want["dx"])
        assert lows_highs[k][0] == want["low"] and lows_highs[k][1] == want["high"], k
        np.testing.assert_array_equal(edges[k].cpu().numpy(), want["edges"], err_msg=str(k))


# ---- candidates, scores, full finder ---------------------------------------------------------
def disc_image(seed=0):
    discs = [(40, 50, 14), (80, 100, 20), (30, 120, 9), (100, 30, 11)]
    return discs, oc.to_uint8(synthetic_discs(130, 150, discs, seed=seed, noise=5))


def test_cell_lists_and_draws_match_oracle_bit_for_bit(cuda_device):
    from magnify_b200 import circles as mc

    _, img = disc_image()
    other = oc.to_uint8(synthetic_discs(130, 150, [(60, 70, 25)], seed=4, noise=5))
    batch = np.stack([img, other, np.zeros_like(img)])               # third image: no edges at all
    edges, dx, dy = mc.find_edges(dev(batch, cuda_device), 0.3, 0.95)
    lists = mc.EdgeLists(edges, 20)
    want = [oc.grid_lists(oc.edge_stages(b, 0.3, 0.95)["edges"], 20) for b in batch[:2]]
    np.testing.assert_array_equal(lists.grid_coords(), np.concatenate([w[0] for w in want]))
    assert lists.total == sum(len(w[0]) for w in want)
    rng = np.random.default_rng(2)
    num_iter = 700
    randoms = rng.integers(0, 2**32, (3, num_iter, 3), dtype=np.uint64)
    raw, circles = mc.sample_circles(lists, num_iter, 6, 26, randoms=dev(randoms.astype(np.uint32).view(np.int32), cuda_device),
                                     want_raw=True)
    raw = raw.cpu().numpy()
    found = circles.cpu().numpy()
    for k in range(2):
        e = oc.edge_stages(batch[k], 0.3, 0.95)["edges"]
        want_raw = oc.sampled_circles(e, 20, randoms[k])
        np.testing.assert_array_equal(raw[k].view(np.uint32), want_raw.view(np.uint32))      # bits, NaNs included
        want_set = {tuple(c) for c in oc.filter_round(want_raw, 6, 26, e.shape)}
        got_set = {tuple(c[1:]) for c in found if c[0] == k}
        assert got_set == want_set and len(got_set) > 10
        assert sum(c[0] == k for c in found) == len(got_set)                                  # de-duplicated
    assert np.isnan(raw[2]).all() and not (found[:, 0] == 2).any()


def test_scores_match_oracle(cuda_device):
    from magnify_b200 import circles as mc

    _, img = disc_image(1)
    st = oc.edge_stages(img, 0.3, 0.95)
    edges, dx, dy = mc.find_edges(dev(img, cuda_device), 0.3, 0.95)
    lists = mc.EdgeLists(edges, 20)
    _, circles = mc.sample_circles(lists, 3000, 6, 26, seed=5)
    found = circles.cpu().numpy()
    assert len(found) > 100
    scores = mc.score_circles(circles, edges, mc.gradient_angles(dx, dy), 6, 26).cpu().numpy()
    want = oc.perimeter_scores(found[:, 1:], st["edges"], st["dx"], st["dy"], 26)
    # float32 arctan2 of NumPy vs float64 arctan2 rounded once: a few ulp of the angle per term
    np.testing.assert_allclose(scores, want, rtol=0, atol=2e-6)
    assert scores.max() > 0.5


def match(found, truth, tol=2):
    return all(any(abs(f[0] - t[0]) <= tol and abs(f[1] - t[1]) <= tol and abs(f[2] - t[2]) <= tol for f in found)
               for t in truth)


def test_find_circles_recovers_planted_discs(cuda_device):
    """The contract of the reference's own tests (tests/test_beads.py:62-66,93-96: centres and radii
    within tolerance): every planted disc is found, best first, deterministically for a seed."""
    from magnify_b200 import circles as mc

    discs, img = disc_image(2)
    args = dict(low_edge_quantile=0.1, high_edge_quantile=0.9, grid_length=20, num_iter=20000, min_radius=6,
                max_radius=26, min_roundness=0.3, min_dist=6)
    circles, scores = mc.find_circles(dev(img, cuda_device), seed=7, **args)
    assert circles.dtype == np.int32 and scores.dtype == np.float32
    assert match(circles, discs) and (np.diff(scores) <= 0).all()
    again, scores2 = mc.find_circles(dev(img, cuda_device), seed=7, **args)
    np.testing.assert_array_equal(circles, again)
    np.testing.assert_array_equal(scores, scores2)
    # survivors do not overlap (utils.py:252-285)
    for i in range(len(circles)):
        for j in range(i):
            assert np.hypot(*(circles[i, :2] - circles[j, :2])) > 6
    # batch: the same image twice plus a blank one
    batch = dev(np.stack([img, np.zeros_like(img), img]), cuda_device)
    res = mc.find_circles(batch, seed=7, **args)
    assert len(res) == 3 and len(res[1][0]) == 0 and match(res[0][0], discs) and match(res[2][0], discs)
    # min_dist = 0 (the chip refinement, find.py:341-356): best circle = one of the discs
    best, s = mc.find_circles(dev(img, cuda_device), seed=3, **dict(args, min_dist=0))
    assert len(best) and match(discs, [best[int(np.argmax(s))]], tol=2)


def test_find_circles_agrees_with_reference_run(cuda_device, golden):
    """tests/golden/circles.npz: the reference's own find_circles on a fixture where its random
    search converges (same circles on every run).  The GPU finder must report exactly those
    circles, with the reference's scores up to the float32 arctan2 difference."""
    from magnify_b200 import circles as mc

    g = golden("circles")
    kw = {k: g[k].item() for k in ("low_edge_quantile", "high_edge_quantile", "grid_length", "num_iter", "min_radius",
                                    "max_radius", "min_roundness", "min_dist")}
    image = dev(g["image"], cuda_device)
    edges, _, _ = mc.find_edges(image, kw["low_edge_quantile"], kw["high_edge_quantile"])
    np.testing.assert_array_equal(edges.cpu().numpy(), g["edges"])
    want = {tuple(c): s for c, s in zip(g["circles"], g["scores"])}
    for seed in (1, 2, 3):
        circles, scores = mc.find_circles(image, seed=seed, **kw)
        got = {tuple(c): s for c, s in zip(circles, scores)}
        assert set(got) == set(want), (seed, sorted(got), sorted(want))
        for c in want:
            assert abs(got[c] - want[c]) <= 2e-6, (c, got[c], want[c])


def test_device_neighbour_suppression_equals_sequential_pass(cuda_device):
    """mgb_filter_neighbors_device (rounds of the fixed-priority rule) == the sequential ring-claim
    pass of utils.py:252-285 (host version, itself equal to the reference's numba function)."""
    from magnify_b200 import circles as mc

    rng = np.random.default_rng(0)
    for trial in range(40):
        b = int(rng.integers(1, 4))
        h, w = int(rng.integers(40, 300)), int(rng.integers(40, 300))
        min_dist, max_radius = int(rng.integers(1, 12)), int(rng.integers(12, 25))
        lo = -min(min_dist + 1, max_radius)                       # stay out of the raster wrap-around regime
        per_image = []
        for k in range(b):
            n = int(rng.integers(0, 400))
            c = np.stack([np.full(n, k), rng.integers(lo, h + max_radius, n), rng.integers(lo, w + max_radius, n),
                          rng.integers(3, max_radius + 1, n)], 1).astype(np.int32)
            if trial % 4 == 0 and n:                              # dense clusters: long dependency chains
                c[:, 1] = rng.integers(10, 30, n)
                c[:, 2] = np.sort(rng.integers(0, w, n))
            per_image.append(c)
        circles = np.concatenate(per_image)
        keep, rounds = mc.filter_neighbors_device(dev(circles, cuda_device), b, h, w, max_radius, min_dist, return_rounds=True)
        want = np.concatenate([mc.filter_neighbors(c[:, 1:], min_dist) for c in per_image]) if len(circles) else np.zeros(0, bool)
        np.testing.assert_array_equal(keep.cpu().numpy(), want, err_msg=f"trial {trial}")
        assert rounds >= (1 if len(circles) else 0)
    far = dev(np.array([[0, -9, 5, 12], [0, 4, 4, 6]], np.int32), cuda_device)
    assert mc.needs_host_suppression(far, 3) and not mc.needs_host_suppression(far, 8)
    cm = mc.conflict_map(2)
    assert cm.shape == (9, 9) and cm[4, 4] == 1 and (cm == cm[::-1, ::-1]).all()


def test_bead_field_detections_agree_with_reference_run(cuda_device, golden):
    """2048^2 field with 300 beads, the reference's default 5e6 draws: the reference's own detections
    (tests/golden/bead_field_reference.npz, one unseeded run) and the GPU finder's agree bead for
    bead -- at least 98 % of either set has a counterpart within one pixel in row, col and radius."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import finder_bench as fb
    from magnify_b200 import circles as mc

    g = golden("bead_field_reference")
    kw = {k: g[k].item() for k in fb.BEADS}
    assert kw == fb.BEADS
    img = mc.to_uint8(dev(fb.bead_image(), cuda_device))
    circles, scores = mc.find_circles(img, seed=0, **kw)
    ref = g["circles"]
    assert abs(len(circles) - len(ref)) <= 6

    def covered(a, b):
        d = np.abs(a[:, None, :].astype(np.int64) - b[None, :, :]).max(axis=2)
        return (d.min(axis=1) <= 1).mean()

    assert covered(ref, circles) >= 0.98 and covered(circles, ref) >= 0.98


def test_device_suppression_chain_limit(cuda_device):
    """A line of circles where every one conflicts with its predecessor is the worst case for the
    round-based rule (one decision per round): short chains settle and equal the sequential pass,
    a chain beyond the round limit raises so that the finder takes the host pass."""
    from magnify_b200 import circles as mc

    def chain(n):                      # centres 3 px apart, best first: kept, rejected, kept, ...
        return np.stack([np.zeros(n), np.full(n, 20), 10 + 3 * np.arange(n), np.full(n, 8)], 1).astype(np.int32)

    short = chain(150)
    keep, rounds = mc.filter_neighbors_device(dev(short, cuda_device), 1, 64, 1000, 10, 4, return_rounds=True)
    np.testing.assert_array_equal(keep.cpu().numpy(), mc.filter_neighbors(short[:, 1:], 4))
    assert 10 < rounds <= 256
    with pytest.raises(mc.SuppressionNotSettled):
        mc.filter_neighbors_device(dev(chain(3000), cuda_device), 1, 64, 9100, 10, 4)

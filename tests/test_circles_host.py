"""Host formulas of the circle finder front end (magnify_b200/circles.py) against NumPy itself,
and the cv2/NumPy oracle against the reference's own find_circles run in place."""
import numpy as np
import pytest

from magnify_b200 import circles as mc


def synthetic_discs(h, w, discs, seed=0, noise=12, level=3000, dtype=np.uint16):
    rng = np.random.default_rng(seed)
    img = (rng.random((h, w)) * noise).astype(np.float64) * 20
    yy, xx = np.mgrid[0:h, 0:w]
    for cy, cx, r in discs:
        img[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] += level
    return img.astype(dtype)


def test_linear_quantile_is_numpy_quantile():
    rng = np.random.default_rng(0)
    for trial in range(1500):
        n = int(rng.integers(1, 5000)) if trial % 3 else int(rng.integers(1, 40))
        a = np.sqrt(rng.integers(0, 33_000_000, n).astype(np.float32)) if trial % 2 else \
            rng.integers(0, 50, n).astype(np.float32)
        q = float(rng.random()) if trial % 5 else [0.0, 1.0, 0.5, 0.1, 0.9][trial % 7 % 5]
        if trial % 11 == 0:
            q = np.float64(1 - np.pi * 8 / 72**2)          # the chip finder's high quantile, find.py:345-347
        want = np.quantile(a, q)
        srt = np.sort(a)
        _, prev, nxt = mc._quantile_indexes(n, q)
        lower = srt[prev]
        upper = srt[min(nxt, n - 1)] if prev != -1 else srt[-1]
        got = mc.linear_quantile(n, q, lower, upper, srt[-1])
        assert got == want and got.dtype == want.dtype, (n, q, got, want)
    with pytest.raises(ValueError):
        mc._quantile_indexes(10, 1.5)


def test_canny_thresholds():
    assert mc.canny_thresholds(3.5, 10.25) == (12, 105)
    assert mc.canny_thresholds(10.25, 3.5) == (12, 105)            # swapped when out of order
    assert mc.canny_thresholds(0.0, 40000.0) == (0, 32767 * 32767)
    assert mc.canny_thresholds(-2.0, 2.0) == (-2, 4)


def test_oracle_edges_are_the_reference_functions_edges():
    from oracle import circles as oc
    from oracle._refload import reference_find_circles_stages

    pytest.importorskip("cv2")
    img = oc.to_uint8(synthetic_discs(300, 260, [(60, 70, 15), (150, 120, 22), (220, 200, 12), (100, 200, 18)]))
    for low_q, high_q in ((0.1, 0.9), (0.5, 0.97)):
        ref = reference_find_circles_stages(img, low_q, high_q, 20, 2000, 8, 30, 0.2, 8)
        if ref is None:
            pytest.skip("/root/reference not available (GPU box)")
        stages = oc.edge_stages(img, low_q, high_q)
        np.testing.assert_array_equal(stages["edges"], ref[0])
    utils_to_uint8 = __import__("oracle._refload", fromlist=["x"]).load_reference_utils().to_uint8
    for arr in (synthetic_discs(20, 30, [(10, 10, 5)]), np.full((4, 4), 7, np.uint16), np.zeros((0, 3), np.float32),
                synthetic_discs(20, 30, [(10, 10, 5)], dtype=np.float32) - 50):
        np.testing.assert_array_equal(oc.to_uint8(arr), utils_to_uint8(arr))


def reference_utils():
    from oracle._refload import load_reference_utils

    utils = load_reference_utils()
    if utils is None:
        pytest.skip("/root/reference not available (GPU box)")
    return utils


def test_perimeter_and_suppression_match_reference_numba():
    utils = reference_utils()
    for r in list(range(1, 40)) + [60, 100, 255]:
        for four in (False, True):
            np.testing.assert_array_equal(mc.circle_perimeter(r, four), utils.circle_points(r, four))
    rng = np.random.default_rng(0)
    for trial in range(120):
        n, min_dist = int(rng.integers(1, 300)), int(rng.integers(1, 12))
        c = np.stack([rng.integers(-5, 120, n), rng.integers(-5, 150, n), rng.integers(3, 20, n)], 1).astype(np.int32)
        if trial % 3 == 0:
            c[:, :2] = np.abs(c[:, :2])
        np.testing.assert_array_equal(mc.filter_neighbors(c, min_dist), utils.filter_neighbors(c, min_dist))
    assert mc.filter_neighbors(np.empty((0, 3), np.int32), 3).shape == (0,)


def test_oracle_circumcircle_is_numbas_arithmetic():
    """Three edge pixels in one grid cell: the reference's candidate_circles can only produce the
    27 ordered draws of them, so the SET of its float32 outputs pins the arithmetic bit for bit."""
    from oracle import circles as oc

    utils = reference_utils()
    rng = np.random.default_rng(0)
    for trial in range(25):
        pts = set()
        while len(pts) < 3:
            pts.add((int(rng.integers(20, 40)), int(rng.integers(40, 60))))
        pts = sorted(pts)
        edges = np.zeros((100, 100), np.uint8)
        for r, c in pts:
            edges[r, c] = 1
        got = {tuple(row.view(np.uint32).tolist()) for row in utils.candidate_circles(edges, 20, 3000)}
        want = {tuple(oc.circumcircle(a, b, c).view(np.uint32).tolist()) for a in pts for b in pts for c in pts}
        assert got == want, trial


def test_oracle_grid_lists_and_scores_match_reference_numba():
    from oracle import circles as oc

    utils = reference_utils()
    img = oc.to_uint8(synthetic_discs(120, 150, [(40, 50, 14), (80, 100, 20), (30, 120, 9)], noise=6))
    st = oc.edge_stages(img, 0.3, 0.95)
    coords, starts, counts = oc.grid_lists(st["edges"], 20)
    rc, rs, rn = utils.grid_array(st["edges"], 20)
    np.testing.assert_array_equal(coords, rc)
    np.testing.assert_array_equal(starts, rs)
    np.testing.assert_array_equal(counts, rn)
    # scores: the reference's own mean_grad on the same circles
    rng = np.random.default_rng(1)
    raw = oc.sampled_circles(st["edges"], 20, rng.integers(0, 2**32, (400, 3), dtype=np.uint64))
    circles = oc.filter_round(raw, 6, 24, img.shape)
    assert len(circles) > 20
    mine = oc.perimeter_scores(circles, st["edges"], st["dx"], st["dy"], 24)
    pad = 2 * 24
    angles, padded = np.pad(np.arctan2(st["dy"], st["dx"]), pad), np.pad(st["edges"], pad)
    for k, (row, col, radius) in enumerate(circles):
        pts = utils.circle_points(int(radius))
        want = utils.mean_grad(angles, padded, np.array([[row + pad, col + pad]], np.int32), pts)[0] / len(pts)
        assert mine[k] == np.float32(want), (k, mine[k], want)

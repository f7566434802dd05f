"""magnify_b200/gridfit.py (all comb offsets evaluated at once, closed-form line fits) side by side
with the reference's own helpers (find.py:630-757) loaded in place, and with the tail of
ButtonFinder.find_centers (find.py:233-306): identical labels, fits equal to rounding."""
import numpy as np
import pytest

from magnify_b200 import gridfit


def reference_find():
    from oracle._refload import load_reference_find

    mod = load_reference_find()
    if mod is None:
        pytest.skip("/root/reference not available (GPU box)")
    return mod


def jittered_grid(rng, rows, cols, row_dist, col_dist, top, left, drop=0.1, extra=5, shear=0.01):
    yy, xx = np.mgrid[0:rows, 0:cols].astype(float)
    y = top + yy * row_dist + shear * xx * col_dist + rng.normal(0, 1.5, yy.shape)
    x = left + xx * col_dist - shear * yy * row_dist + rng.normal(0, 1.5, yy.shape)
    pts = np.stack([y.ravel(), x.ravel()], 1)
    pts = pts[rng.random(len(pts)) > drop]
    noise = np.stack([rng.uniform(0, top + rows * row_dist + 40, extra), rng.uniform(0, left + cols * col_dist + 40, extra)], 1)
    return np.concatenate([pts, noise])


def test_helpers_match_reference_functions():
    ref = reference_find()
    rng = np.random.default_rng(0)
    for trial in range(12):
        rows, cols = int(rng.integers(2, 9)), int(rng.integers(2, 7))
        row_dist, col_dist = float(rng.uniform(30, 60)), float(rng.uniform(50, 90))
        top, left = float(rng.uniform(10, 60)), float(rng.uniform(10, 60))
        pts = jittered_grid(rng, rows, cols, row_dist, col_dist, top, left)
        shape = (int(top + rows * row_dist + 80), int(left + cols * col_dist + 80))
        ideal_r, ideal_c = np.full(rows, cols), np.full(cols, rows)
        a = gridfit.comb_labels(pts[:, 0], shape[0], rows, row_dist, ideal_r, 10)
        np.testing.assert_array_equal(a, ref.cluster_1d(pts[:, 0], shape[0], rows, row_dist, ideal_r, 10))
        b = gridfit.spaced_labels(pts[:, 1], left - 20, cols, 40, col_dist - 40)
        np.testing.assert_array_equal(b, ref.label_clusters(pts[:, 1], left - 20, cols, 40, col_dist - 40))
        keep = (a >= 0) & (b >= 0)
        got = gridfit.fit_lines(pts[keep, 1], pts[keep, 0], a[keep], rows, ideal_r)
        want = ref.regress_clusters(pts[keep, 1], pts[keep, 0], a[keep], rows, ideal_r)
        np.testing.assert_allclose(got[0], want[0], rtol=1e-10)
        np.testing.assert_allclose(got[1], want[1], rtol=1e-10, atol=1e-9)
    one = gridfit.fit_lines(np.array([1.0, 2.0, 4.0]), np.array([2.0, 4.1, 8.2]), np.zeros(3, int), 1, np.array([3]))
    ref_one = ref.regress_clusters(np.array([1.0, 2.0, 4.0]), np.array([2.0, 4.1, 8.2]), np.zeros(3, int), 1, np.array([3]))
    np.testing.assert_allclose(one, ref_one, rtol=1e-12)


def test_grid_centers_matches_reference_find_centers_tail(monkeypatch):
    """ButtonFinder.find_centers with its circle finder pinned to given points
    (find.py:205-306 executed in place) == gridfit.merge_channel_points + grid_centers."""
    ref = reference_find()
    from oracle._refload import LabelledArray

    rng = np.random.default_rng(3)
    rows, cols, row_dist, col_dist = 6, 4, 50.0, 80.0
    tag = np.full((rows, cols), "a", dtype="<U8")
    tag[2, 1] = ""
    per_channel = [jittered_grid(rng, rows, cols, row_dist, col_dist, 40, 50), jittered_grid(rng, rows, cols, row_dist, col_dist, 40, 50)]
    shape = (400, 420)
    class Counts:
        def __init__(self, mask):
            self.mask = mask

        def sum(self, dim):
            return LabelledArray(self.mask.sum(axis=1 if dim == "mark_col" else 0), ("k",), {})

    class Tag:
        def __ne__(self, other):
            return Counts(tag != other)

    assay = type("A", (), {"tag": Tag(), "sizes": {"mark_row": rows, "mark_col": cols}})()
    images = [LabelledArray(np.zeros(shape, np.uint8), ("im_y", "im_x"), {}) for _ in per_channel]
    for top, left in ((None, None), (15, 20)):
        finder = ref.ButtonFinder(row_dist=row_dist, col_dist=col_dist, min_button_diameter=10, max_button_diameter=24,
                                  chamber_diameter=40, top_chamber=top, left_chamber=left, low_edge_quantile=0.1,
                                  high_edge_quantile=0.9, num_iter=100, min_roundness=0.2, cluster_penalty=10,
                                  roi_length=None, progress_bar=False, search_timestep=0, search_channel=None,
                                  interactive=False)
        calls = iter(per_channel)

        def pinned_circles(img, **kw):
            pts = next(calls)
            return np.column_stack([pts, np.full(len(pts), 9.0)]), None

        monkeypatch.setattr(ref.utils, "find_circles", pinned_circles)
        want_x, want_y = finder.find_centers(images, assay)
        pts = np.empty((0, 2))
        for new in per_channel:
            pts = gridfit.merge_channel_points(pts, new, finder.chamber_radius)
        got_x, got_y = gridfit.grid_centers(pts, tag, shape, row_dist, col_dist, finder.chamber_radius, top, left, 10)
        np.testing.assert_allclose(got_x, want_x, rtol=1e-10, atol=1e-8)
        np.testing.assert_allclose(got_y, want_y, rtol=1e-10, atol=1e-8)

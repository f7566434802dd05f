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

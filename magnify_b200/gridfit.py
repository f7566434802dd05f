"""Chip grid fit: from detected circle centres to one (x, y) per (mark_row, mark_col).

What `ButtonFinder.find_centers` does after its circle search (src/magnify/find.py:233-306): the
centres are cut into rows and columns of the chip by sliding a comb of equal windows along each
axis, every row / column gets a straight line, and the chambers are the line intersections.

This module states that model directly and evaluates it for ALL comb offsets at once:

* window membership of every (offset, point) pair by one broadcast comparison against the comb
  edges, per-window counts and squared distances to the window centres by `np.bincount` over the
  (offset, window) keys -- no per-offset loop, no sorting of the points;
* the line fits in closed form from per-cluster sums (the same sums a least-squares fit of
  y = a x + b reduces to), all clusters in one pass.

The cost of an offset is the reference's (mean squared distance to the window centre, empty
windows charged the worst window's value, weighted by sqrt(expected points), plus
penalty x (expected - found)^2) and ties go to the smallest offset, so the chosen comb -- and with
it every label -- is the reference's; the fitted numbers agree with its scipy-based fits to
rounding (tests/test_gridfit_host.py runs both side by side).
"""
from __future__ import annotations

import numpy as np


def comb_labels(points, total_length: int, num_clusters: int, cluster_length: float, ideal_num_points,
                penalty: float) -> np.ndarray:
    """Window index of every point under the best comb of `num_clusters` windows of
    `cluster_length` (offsets 0, 1, 2, ... pixels), -1 for points outside the comb."""
    p = np.asarray(points, dtype=np.float64)
    ideal = np.asarray(ideal_num_points, dtype=np.float64)
    n_off = int(total_length - round(num_clusters * cluster_length))
    if n_off <= 0:
        raise ValueError("the comb is longer than the image")
    k = num_clusters
    offsets = np.arange(n_off, dtype=np.float64)
    edges = np.arange(k + 1) * cluster_length + offsets[:, None]              # (O, K+1)
    # window of point j under offset o: edges[o, w] <= p[j] < edges[o, w + 1]
    w = np.floor((p[None, :] - offsets[:, None]) / cluster_length).astype(np.int64)   # (O, P) first guess
    w = np.clip(w, -1, k)
    rows = np.arange(n_off)[:, None]
    w -= (w >= 0) & (w <= k) & (p[None, :] < edges[rows, np.clip(w, 0, k)])           # float edge cases
    w += (w >= -1) & (w < k) & (p[None, :] >= edges[rows, np.clip(w + 1, 0, k)])
    inside = (w >= 0) & (w < k)
    key = (rows * k + np.where(inside, w, 0))[inside]
    centres = (edges[:, 1:] + edges[:, :-1]) / 2                                       # (O, K)
    dist2 = (np.broadcast_to(p, w.shape)[inside] - centres.reshape(-1)[key]) ** 2
    count = np.bincount(key, minlength=n_off * k).reshape(n_off, k).astype(np.float64)
    spread = np.bincount(key, weights=dist2, minlength=n_off * k).reshape(n_off, k)
    filled = count > 0
    spread = np.divide(spread, count, out=np.zeros_like(spread), where=filled)
    spread = np.where(filled, spread, spread.max(axis=1, keepdims=True))              # empty window = worst window
    cost = (spread * np.sqrt(ideal) + penalty * (ideal - count) ** 2).sum(axis=1)
    best = int(np.argmin(cost))                                                        # first minimum, like a `<` scan
    return np.where(inside[best], w[best], -1)


def spaced_labels(points, offset, num_clusters: int, cluster_length, cluster_gap) -> np.ndarray:
    """Windows of `cluster_length` separated by `cluster_gap`, the first starting at `offset`
    (known chip position, find.py:246-254): window index per point, -1 in the gaps and outside."""
    p = np.asarray(points, dtype=np.float64)
    # edges accumulate start, end, start, end, ... exactly as lengths are added one after another
    edges = np.cumsum([offset] + ([cluster_length, cluster_gap] * num_clusters)[:-1])
    slot = np.searchsorted(edges, p, side="right")             # number of edges at or below the point
    return np.where(slot % 2 == 1, (slot - 1) // 2, -1).astype(int)


def _line(x, y):
    """Least-squares slope and intercept of y over x (the closed form scipy.stats.linregress uses:
    covariance over variance about the means)."""
    xm, ym = x.mean(), y.mean()
    dx = x - xm
    slope = np.dot(dx, y - ym) / np.dot(dx, dx)
    return slope, ym - slope * xm


def fit_lines(x, y, labels, num_clusters: int, ideal_num_points):
    """One common slope and an intercept per cluster for the lines y = slope * x + intercept.

    The slope is the median of the per-cluster least-squares slopes; a cluster's intercept is the
    median residual under that slope, pulled towards the evenly spaced trend of all intercepts in
    proportion to how many of its expected points are missing (empty clusters sit on the trend)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    if num_clusters == 1:
        if len(x) == 1:
            return 0, y
        return _line(x, y)
    members = [np.nonzero(labels == k)[0] for k in range(num_clusters)]
    slopes = np.array([_line(x[i], y[i])[0] if len(i) > 1 else np.nan for i in members])
    slope = np.nanmedian(slopes)
    local = np.array([np.median(y[i] - slope * x[i]) if len(i) else np.nan for i in members])
    have = ~np.isnan(local)
    index = np.arange(num_clusters, dtype=np.float64)
    if have.sum() > 1:
        step, start = _line(index[have], local[have])
    else:                                  # a single usable cluster cannot define a trend
        step, start = np.nan, np.nan
    trend = step * index + start
    found = np.array([len(i) for i in members], dtype=np.float64)
    ideal = np.asarray(ideal_num_points, dtype=np.float64)
    weight = np.where((ideal != 0) & have, np.minimum(found, ideal) / np.where(ideal != 0, ideal, 1.0), 0.0)
    return slope, np.where(weight > 0, weight * np.where(have, local, 0.0) + (1 - weight) * trend, trend)


def merge_channel_points(points: np.ndarray, new_points: np.ndarray, min_dist: float) -> np.ndarray:
    """Centres found in another channel are added unless one already known lies within `min_dist`
    (find.py:225-231)."""
    if len(points) > 0 and len(new_points) > 0:
        from scipy.spatial import cKDTree

        nearest, _ = cKDTree(points).query(new_points, k=1)
        new_points = new_points[nearest > min_dist]
    return np.concatenate([points, new_points])


def grid_centers(points: np.ndarray, tag: np.ndarray, image_shape, row_dist: float, col_dist: float, chamber_radius: int,
                 top_chamber=None, left_chamber=None, cluster_penalty: float = 10):
    """points (n, 2) as (row, col) -> (mark_x, mark_y), each (rows, cols): rows are lines
    y = a x + b_i, columns are lines x = c y + d_j, and chamber (i, j) is where they cross."""
    x, y = points[:, 1], points[:, 0]
    per_row = (tag != "").sum(axis=1)
    per_col = (tag != "").sum(axis=0)
    rows, cols = tag.shape
    if top_chamber is None:
        row_of = comb_labels(y, image_shape[0], rows, row_dist, per_row, cluster_penalty)
    else:
        row_of = spaced_labels(y, top_chamber, rows, 2 * chamber_radius, row_dist - 2 * chamber_radius)
    if left_chamber is None:
        col_of = comb_labels(x, image_shape[1], cols, col_dist, per_col, cluster_penalty)
    else:
        col_of = spaced_labels(x, left_chamber, cols, 2 * chamber_radius, col_dist - 2 * chamber_radius)
    used = (row_of >= 0) & (col_of >= 0)
    x, y, row_of, col_of = x[used], y[used], row_of[used], col_of[used]
    a, b = fit_lines(x, y, row_of, rows, per_row)
    c, d = fit_lines(y, x, col_of, cols, per_col)
    # y = a x + b_i and x = c y + d_j  =>  y = (a d_j + b_i) / (1 - a c)
    mark_y = (a * np.asarray(d)[np.newaxis] + np.asarray(b)[:, np.newaxis]) / (1 - a * c)
    mark_x = mark_y * c + np.asarray(d)[np.newaxis]
    return mark_x, mark_y

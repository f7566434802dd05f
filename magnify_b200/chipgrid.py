"""Chip grid fitting: from detected circle centres to one (x, y) per (mark_row, mark_col).

Host-side NumPy restatement of the second half of `ButtonFinder.find_centers`
(src/magnify/find.py:230-306) and its helpers `cluster_1d` (:630-677), `label_clusters`
(:680-697) and `regress_clusters` (:700-757).  A few thousand points and a few hundred candidate
offsets: no GPU work here, but the arithmetic (cumulative sums, their differences, the order of
the cost terms) follows the reference so that the chosen clustering and the fitted lines are the
same numbers (checked against the reference's functions in tests/test_chipgrid_host.py).
"""
from __future__ import annotations

import numpy as np


def cluster_1d(points, total_length: int, num_clusters: int, cluster_length: float, ideal_num_points, penalty: float):
    """Slide `num_clusters` equal windows of `cluster_length` over the sorted points and keep the
    offset with the lowest cost (within-window variance weighted by sqrt(ideal count), plus
    `penalty` x squared count error).  Returns a window label per point, -1 outside all windows."""
    points = np.asarray(points)
    ideal = np.asarray(ideal_num_points)
    order = np.argsort(points)
    srt = points[order]
    steps = np.arange(num_clusters + 1) * cluster_length
    best_cost, best_spans = np.inf, None
    for offset in range(total_length - round(num_clusters * cluster_length)):
        edges = steps + offset
        mids = (edges[1:] + edges[:-1]) / 2
        spans = np.searchsorted(srt, edges)
        per_window = spans[1:] - spans[:-1]
        sq = (srt[spans[0]:spans[-1]] - np.repeat(mids, per_window)) ** 2
        running = np.insert(np.cumsum(sq), 0, 0)
        cost = np.diff(running[spans - spans[0]])
        filled = per_window > 0
        cost[filled] /= per_window[filled]
        cost[~filled] = np.max(cost)
        cost *= np.sqrt(ideal)
        cost = cost + penalty * (ideal - per_window) ** 2
        total = cost.sum()
        if total < best_cost:
            best_cost, best_spans = total, spans
    labels = -np.ones_like(srt, dtype=int)
    labels[best_spans[0]:best_spans[-1]] = np.repeat(np.arange(num_clusters), best_spans[1:] - best_spans[:-1])
    return labels[np.argsort(order)]


def label_clusters(points, offset, num_clusters: int, cluster_length, cluster_gap):
    """Windows of `cluster_length` separated by `cluster_gap`, the first starting at `offset`."""
    points = np.asarray(points)
    order = np.argsort(points)
    srt = points[order]
    widths = [offset] + ([cluster_length, cluster_gap] * num_clusters)[:-1]
    spans = np.searchsorted(srt, np.cumsum(widths))
    labels = -np.ones_like(srt, dtype=int)
    for k in range(num_clusters):
        labels[spans[2 * k]:spans[2 * k + 1]] = k
    return labels[np.argsort(order)]


def regress_clusters(x, y, labels, num_clusters: int, ideal_num_points):
    """One common slope (median of the per-cluster regressions) and an intercept per cluster,
    blended with the evenly spaced global estimate according to how full the cluster is."""
    from scipy.stats import linregress

    x, y = np.asarray(x), np.asarray(y)
    if num_clusters == 1:
        if len(x) == 1:
            return 0, y
        return linregress(x, y)[:2]
    groups = [(x[labels == k], y[labels == k]) for k in range(num_clusters)]
    slopes = np.full(num_clusters, np.nan)
    intercepts = np.full(num_clusters, np.nan)
    for k, (gx, gy) in enumerate(groups):
        if len(gx) > 1:
            fit = linregress(gx, gy)
            slopes[k], intercepts[k] = fit[0], fit[1]
    slope = np.nanmedian(slopes)
    for k, (gx, gy) in enumerate(groups):
        if len(gx) > 0:
            intercepts[k] = np.median(gy - slope * gx)
    known = ~np.isnan(intercepts)
    index = np.arange(num_clusters)
    trend = linregress(index[known], intercepts[known])
    for k, (gx, _) in enumerate(groups):
        spaced = trend[0] * k + trend[1]
        if ideal_num_points[k] != 0 and known[k]:
            weight = min(len(gx), ideal_num_points[k]) / ideal_num_points[k]
            intercepts[k] = weight * intercepts[k] + (1 - weight) * spaced
        else:
            intercepts[k] = spaced
    return slope, intercepts


def merge_channel_points(points: np.ndarray, new_points: np.ndarray, min_dist: float) -> np.ndarray:
    """find.py:225-231: append the centres found in another channel unless they lie within
    `min_dist` of a centre already known."""
    if len(points) > 0:
        gaps = np.linalg.norm(points[np.newaxis] - new_points[:, np.newaxis], axis=2)
        new_points = new_points[np.min(gaps, axis=1) > min_dist]
    return np.concatenate([points, new_points])


def grid_centers(points: np.ndarray, tag: np.ndarray, image_shape, row_dist: float, col_dist: float, chamber_radius: int,
                 top_chamber=None, left_chamber=None, cluster_penalty: float = 10):
    """find.py:233-306: points (n, 2) as (row, col) -> (mark_x, mark_y), each (rows, cols)."""
    x, y = points[:, 1], points[:, 0]
    per_row = (tag != "").sum(axis=1)
    per_col = (tag != "").sum(axis=0)
    rows, cols = tag.shape
    if top_chamber is None:
        row_labels = cluster_1d(y, image_shape[0], rows, row_dist, per_row, cluster_penalty)
    else:
        row_labels = label_clusters(y, top_chamber, rows, 2 * chamber_radius, row_dist - 2 * chamber_radius)
    if left_chamber is None:
        col_labels = cluster_1d(x, image_shape[1], cols, col_dist, per_col, cluster_penalty)
    else:
        col_labels = label_clusters(x, left_chamber, cols, 2 * chamber_radius, col_dist - 2 * chamber_radius)
    inside = (row_labels >= 0) & (col_labels >= 0)
    x, y, row_labels, col_labels = x[inside], y[inside], row_labels[inside], col_labels[inside]
    row_slope, row_icpt = regress_clusters(x, y, row_labels, rows, per_row)
    col_slope, col_icpt = regress_clusters(y, x, col_labels, cols, per_col)      # x as a function of y
    mark_y = (row_slope * col_icpt[np.newaxis] + row_icpt[:, np.newaxis]) / (1 - row_slope * col_slope)
    mark_x = mark_y * col_slope + col_icpt[np.newaxis]
    return mark_x, mark_y

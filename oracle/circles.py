"""Circle finder front end exactly as the reference computes it -- TEST INFRASTRUCTURE.

`utils.find_circles` (src/magnify/utils.py:100-139) does its edge detection with OpenCV and
NumPy calls; both libraries are in the image (cv2 4.13.0 == the reference's uv.lock pin
4.13.0.90, numpy 2.3.5 == the lock), so the oracle for these steps is the very same calls in the
very same order -- not a restatement.  `oracle/_refload.py::reference_find_circles_stages` runs
the reference's own function and captures its internal `edges` to confirm that.
"""
from __future__ import annotations

import numpy as np


def to_uint8(arr: np.ndarray) -> np.ndarray:
    """utils.py:20-27."""
    if arr.size == 0:
        return arr.astype(np.uint8)
    arr = arr.astype(float)
    arr = arr - np.min(arr)
    if np.max(arr) > 0:
        arr = 255 * arr / np.max(arr)
    return arr.astype(np.uint8)


def edge_stages(img: np.ndarray, low_edge_quantile: float, high_edge_quantile: float) -> dict:
    """utils.py:113-139 with every intermediate kept."""
    import cv2 as cv

    blurred = cv.GaussianBlur(img, (5, 5), 0)                               # :114
    dx = cv.Scharr(blurred, ddepth=cv.CV_32F, dx=1, dy=0)                   # :117
    dy = cv.Scharr(blurred, ddepth=cv.CV_32F, dx=0, dy=1)                   # :118
    grad = np.sqrt(dx**2 + dy**2)                                           # :119
    low = np.quantile(grad, low_edge_quantile)                              # :125
    high = np.quantile(grad, high_edge_quantile)                            # :126
    edges = cv.Canny(dx.astype(np.int16), dy.astype(np.int16), threshold1=low, threshold2=high,
                     L2gradient=True)                                       # :127-133
    edges[edges != 0] = 1                                                   # :139
    return {"blurred": blurred, "dx": dx, "dy": dy, "grad": grad, "low": low, "high": high, "edges": edges}


# ---------------------------------------------------------------------------------------------
# candidates and scores
# ---------------------------------------------------------------------------------------------
def grid_lists(edges: np.ndarray, grid_length: int):
    """utils.py:347-377 `grid_array`: (grid_coords (n,2) int32, grid_starts, grid_counts)."""
    rows, cols = -(-edges.shape[0] // grid_length), -(-edges.shape[1] // grid_length)
    counts = np.zeros((rows, cols), dtype=np.int64)
    starts = np.zeros((rows, cols), dtype=np.int64)
    coords = []
    n = 0
    for i in range(rows):
        for j in range(cols):
            r, c = np.where(edges[i * grid_length:(i + 1) * grid_length, j * grid_length:(j + 1) * grid_length])
            starts[i, j], counts[i, j] = n, len(r)
            coords.append(np.stack([r + i * grid_length, c + j * grid_length], axis=1))
            n += len(r)
    coords = np.concatenate(coords).astype(np.int32) if coords else np.empty((0, 2), np.int32)
    return coords, starts, counts


def circumcircle(p0, p1, p2) -> np.ndarray:
    """utils.py:317-342 for three pixels (row, col) with numba's typing of those lines: float64
    throughout, results stored to float32, the radius computed from the stored float32 centre in
    float32.  Pinned against the reference's own candidate_circles on three-pixel edge maps
    (tests/test_circles_host.py)."""
    p0 = np.asarray(p0, dtype=np.int64)
    q1, q2 = np.asarray(p1, dtype=np.int64) - p0, np.asarray(p2, dtype=np.int64) - p0
    eps = float(np.float32(1e-20))
    with np.errstate(all="ignore"):
        mid1, mid2 = 0.5 * q1.astype(np.float64), 0.5 * q2.astype(np.float64)
        m1 = np.float64(-q1[1]) / (np.float64(q1[0]) + eps)
        m2 = np.float64(-q2[1]) / (np.float64(q2[0]) + eps)
        b1 = mid1[0] - m1 * mid1[1]
        b2 = mid2[0] - m2 * mid2[1]
        c1 = np.float32((b1 - b2) / (m2 - m1 + eps))
        c0 = np.float32(m1 * np.float64(c1) + b1)
        radius = np.sqrt(np.float32(c0 * c0) + np.float32(c1 * c1))
        return np.array([np.float32(np.float64(c0) + p0[0]), np.float32(np.float64(c1) + p0[1]), radius], np.float32)


def sampled_circles(edges: np.ndarray, grid_length: int, randoms: np.ndarray) -> np.ndarray:
    """The draws of utils.py:304-342 for given uniform 32-bit numbers (n, 3): index = u * count >> 32,
    p0 from the cell-major list (uniform over the edge pixels, like the reference's row-major
    `coords`), p1 / p2 from p0's grid cell."""
    coords, starts, counts = grid_lists(edges, grid_length)
    out = np.full((len(randoms), 3), np.nan, dtype=np.float32)
    if len(coords) == 0:
        return out
    for i, (u0, u1, u2) in enumerate(np.asarray(randoms, dtype=np.uint64)):
        p0 = coords[int((u0 * np.uint64(len(coords))) >> np.uint64(32))]
        cell = (p0[0] // grid_length, p0[1] // grid_length)
        st, cnt = int(starts[cell]), np.uint64(counts[cell])
        p1 = coords[st + int((u1 * cnt) >> np.uint64(32))]
        p2 = coords[st + int((u2 * cnt) >> np.uint64(32))]
        out[i] = circumcircle(p0, p1, p2)
    return out


def filter_round(circles: np.ndarray, min_radius: int, max_radius: int, shape) -> np.ndarray:
    """utils.py:157-165."""
    with np.errstate(invalid="ignore"):
        circles = circles[(circles[:, 2] >= min_radius) & (circles[:, 2] <= max_radius)]
    circles = np.round(circles).astype(np.int32)
    return circles[(circles[:, 0] + circles[:, 2] >= 0) & (circles[:, 1] + circles[:, 2] >= 0)
                   & (circles[:, 0] - circles[:, 2] < shape[0]) & (circles[:, 1] - circles[:, 2] < shape[1])]


def perimeter_scores(circles: np.ndarray, edges: np.ndarray, dx: np.ndarray, dy: np.ndarray, max_radius: int):
    """utils.py:169-183 + mean_grad (:221-249) for int32 circles (n, 3): float32 score per circle
    (alignment sum in float64 in perimeter order -> float32 -> / perimeter length)."""
    from . import geometry as g

    grad_angles = np.arctan2(dy, dx)                                        # float32, :169
    pad = 2 * max_radius
    grad_angles = np.pad(grad_angles, pad)
    padded_edges = np.pad(edges, pad)
    scores = np.empty(len(circles), dtype=np.float32)
    for i, (row, col, radius) in enumerate(circles):
        pts = g.circle_points(int(radius))
        expected = np.arctan2(pts[:, 0], pts[:, 1])
        total = 0.0
        for (dr, dc), want in zip(pts, expected):
            y, x = row + pad + dr, col + pad + dc
            if padded_edges[y, x] > 0:
                diff = np.abs(np.float64(grad_angles[y, x]) - want)
                if diff > np.pi:
                    diff = diff - np.pi
                total += 4 * np.abs(diff - np.pi / 2) / np.pi - 1
        scores[i] = np.float32(total) / len(pts)
    return scores

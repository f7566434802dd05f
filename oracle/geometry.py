"""Integer ROI / mask geometry of the reference, restated in plain NumPy / Python.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the reference lines it
follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np


def ceildiv(a: int, b: int) -> int:
    """src/magnify/utils.py:55-57."""
    return -(a // -b)


def bounding_box(x: int, y: int, box_length: int, image_width: int, image_height: int):
    """src/magnify/utils.py:60-80.  Returns (top, bottom, left, right).

    The box is `box_length` wide, centred so that the extra pixel of an odd length goes to the
    bottom/right, then slid back inside the image when it pokes out.
    """
    half_lo = box_length // 2
    half_hi = ceildiv(box_length, 2)
    top, bottom = y - half_lo, y + half_hi
    if top < 0:
        bottom, top = bottom - top, 0
    if bottom > image_height:
        top, bottom = top - (bottom - image_height), image_height
    left, right = x - half_lo, x + half_hi
    if left < 0:
        right, left = right - left, 0
    if right > image_width:
        left, right = left - (right - image_width), image_width
    return top, bottom, left, right


def round_half_even(v) -> np.ndarray:
    """Python's `round()` on float64 (find.py:163-164,328-329,371-372,574-575) is
    round-half-to-even, which is what np.rint does."""
    return np.rint(np.asarray(v, dtype=np.float64)).astype(np.int64)


def boxes_from_centres(x, y, box_length: int, image_width: int, image_height: int) -> np.ndarray:
    """Vectorised `bounding_box(round(x), round(y), ...)` -> int64 (..., 2) of (top, left)."""
    xs = round_half_even(x)
    ys = round_half_even(y)
    out = np.empty(xs.shape + (2,), dtype=np.int64)
    flat = out.reshape(-1, 2)
    for i, (xi, yi) in enumerate(zip(xs.reshape(-1), ys.reshape(-1))):
        t, _, l, _ = bounding_box(int(xi), int(yi), box_length, image_width, image_height)
        flat[i, 0] = t
        flat[i, 1] = l
    return out


def circle_points(r: int, four_connected: bool = False) -> np.ndarray:
    """Perimeter offsets of the reference's Bresenham-style circle, src/magnify/utils.py:433-465.

    Walk one octant from (0, -r): after emitting the 8 mirror images of (a, b) step `a`
    rightwards while still inside the circle, otherwise step `b` towards the centre (and,
    unless four_connected, `a` as well).  The four axis points come first and the four
    diagonal points close the octant when the walk lands on the diagonal.
    """
    if r < 1:
        raise ValueError("radius must be >= 1")
    pts = [(0, -r), (-r, 0), (0, r), (r, 0)]
    a, b = 1, -r
    while a < -b:
        pts += [(a, b), (b, a), (-a, b), (-b, a), (a, -b), (b, -a), (-a, -b), (-b, -a)]
        if a * a + b * b - r * r <= 0:
            a += 1
        else:
            b += 1
            if not four_connected:
                a += 1
    if b == -a:
        pts += [(a, b), (-a, -b), (-a, b), (a, -b)]
    return np.asarray(pts, dtype=np.int32)


def disc_halfwidths(r: int) -> np.ndarray:
    """hw[|drow|] for the filled disc of `filled_circle_points` (src/magnify/utils.py:398-430).

    The reference marks the perimeter in a (2r+1)^2 raster and fills each row between its
    left and right perimeter runs, so a row `drow` is the single span |dcol| <= hw[|drow|] with
    hw = the largest |dcol| the perimeter reaches on that row.
    """
    per = circle_points(r)
    hw = np.zeros(r + 1, dtype=np.int32)
    np.maximum.at(hw, np.abs(per[:, 0]), np.abs(per[:, 1]))
    return hw


def filled_circle_points(r: int) -> np.ndarray:
    """Offsets (drow, dcol) of the reference's filled disc, src/magnify/utils.py:398-430.

    Row-major order (the reference's order differs; only the set matters to its callers).
    """
    hw = disc_halfwidths(r)
    rows = []
    for d in range(-r, r + 1):
        w = int(hw[abs(d)])
        cols = np.arange(-w, w + 1, dtype=np.int32)
        rows.append(np.stack([np.full_like(cols, d), cols], axis=1))
    return np.concatenate(rows, axis=0)


def circle_labels(circles: np.ndarray, num_rows: int, num_cols: int) -> np.ndarray:
    """Label raster of src/magnify/utils.py:380-395.

    circles: integer (M, 3) rows of (row, col, radius).  -1 = no disc, i = only disc i,
    -2 = two or more discs.  Off-image points are skipped (utils.py:389).
    """
    circles = np.asarray(circles)
    owner = np.full((num_rows, num_cols), -1, dtype=np.int32)
    hits = np.zeros((num_rows, num_cols), dtype=np.int32)
    for i in range(len(circles)):
        cy, cx, r = (int(v) for v in circles[i, :3])
        hw = disc_halfwidths(r)
        for d in range(-r, r + 1):
            yy = cy + d
            if yy < 0 or yy >= num_rows:
                continue
            w = int(hw[abs(d)])
            lo, hi = max(cx - w, 0), min(cx + w, num_cols - 1)
            if lo > hi:
                continue
            hits[yy, lo : hi + 1] += 1
            owner[yy, lo : hi + 1] = i
    labels = np.where(hits == 1, owner, np.where(hits == 0, -1, -2)).astype(np.int32)
    return labels


def circle(image_shape, center, radius: int) -> np.ndarray:
    """Filled disc of `utils.circle` (src/magnify/utils.py:30-40) as a bool array.

    The reference rasterises with cv.circle(thickness=-1) (opencv-python-headless 4.13.0.90,
    pinned in uv.lock).  For integer centre/radius that raster equals the integer disc
    dx^2 + dy^2 <= r^2 clipped to the canvas; tests/test_oracle_geometry.py checks this
    against the cv2 in this image for r = 0..256 including centres outside the canvas.
    center = (row, col).
    """
    h, w = image_shape
    yy = np.arange(h, dtype=np.int64)[:, None] - int(center[0])
    xx = np.arange(w, dtype=np.int64)[None, :] - int(center[1])
    if radius < 0:
        return np.zeros((h, w), dtype=bool)
    return (yy * yy + xx * xx) <= int(radius) * int(radius)


def annulus(image_shape, center, outer_radius: int, inner_radius: int) -> np.ndarray:
    """`utils.annulus` (src/magnify/utils.py:43-52): outer disc AND NOT inner disc."""
    return circle(image_shape, center, outer_radius) & ~circle(image_shape, center, inner_radius)

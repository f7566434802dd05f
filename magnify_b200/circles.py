"""Circle finder of the reference (src/magnify/utils.py:100-344) on the GPU -- SURVEY.md
section 8f rows N1 (per-ROI button refinement, find.py:339-360) and N4 (bead detection,
find.py:476-491).

The edge-detection front end (to_uint8 -> GaussianBlur -> Scharr -> gradient quantiles -> Canny,
utils.py:20-27, 113-139) is integer / IEEE-exact and bit-identical to the reference's NumPy +
OpenCV calls (kernels in csrc/circles.cu).  The host only restates two scalar formulas: NumPy's
linear quantile interpolation between two order statistics and OpenCV's threshold squaring.
"""
from __future__ import annotations

import ctypes
import math
from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .ops import _DTYPE_CODE, _check, _ptr, _stream


def _batched(t: torch.Tensor, name: str, dtype) -> Tuple[torch.Tensor, bool]:
    """View a (H, W) or (B, H, W) tensor as (B, H, W); second value: whether a batch dim was added."""
    if t.ndim not in (2, 3):
        raise ValueError(f"{name} must be (H, W) or (B, H, W), got {tuple(t.shape)}")
    _check(t, name, dtype, t.ndim)
    return (t.unsqueeze(0), True) if t.ndim == 2 else (t, False)


def to_uint8(x: torch.Tensor, batched: bool = False) -> torch.Tensor:
    """utils.py:20-27: uint8(255 * (x - min) / (max - min)) in float64, truncated; zeros when the
    array is constant; an empty array stays empty.  With batched=True the leading dim indexes
    independent images, each scaled by its own min / max (find.py:343 per ROI and channel)."""
    if not x.is_cuda or not x.is_contiguous():
        raise ValueError("to_uint8 needs a contiguous CUDA tensor")
    if x.dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported dtype {x.dtype}")
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    if x.numel() == 0:
        return out
    b = x.shape[0] if batched else 1
    minmax = torch.empty((b, 2), dtype=torch.float64, device=x.device)
    _lib.call("mgb_to_uint8", _ptr(x), _DTYPE_CODE[x.dtype], b, x.numel() // b, _ptr(out), _ptr(minmax), _stream())
    return out


def edge_gradients(image: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils.py:114-118 on (H, W) or (B, H, W) uint8 images: Scharr dx, dy of the 5x5-Gaussian-
    blurred image as int16 (the float32 arrays of the reference hold exactly these integers)."""
    img, squeeze = _batched(image, "image", torch.uint8)
    b, h, w = img.shape
    blurred = torch.empty_like(img)
    dx = torch.empty((b, h, w), dtype=torch.int16, device=image.device)
    dy = torch.empty_like(dx)
    _lib.call("mgb_edge_gradients_u8", _ptr(img), b, h, w, _ptr(blurred), _ptr(dx), _ptr(dy), _stream())
    return (dx[0], dy[0]) if squeeze else (dx, dy)


def linear_quantile(n: int, q, lower: np.float32, upper: np.float32, last: np.float32) -> np.float32:
    """np.quantile(a, q) (method "linear") of a float32 array of n values, given its order
    statistics a_sorted[floor(v)], a_sorted[floor(v)+1] and a_sorted[-1] for v = (n-1)*q.

    Follows NumPy 2.x step by step, dtypes included: a Python-float q is cast to the array's
    float32, so the virtual index (n-1)*q is a float32 product; gamma = v - floor(v) is formed in
    float64 and cast to float32; the interpolation a + (b-a)*g (or b - (b-a)*(1-g) for g >= 0.5)
    runs in float32."""
    v, prev, _ = _quantile_indexes(n, q)
    if prev == -1:
        lower = upper = last
    gamma = np.asanyarray(np.asanyarray(v - np.asanyarray(np.intp(prev))), dtype=v.dtype)
    a, b = np.asanyarray(lower, dtype=np.float32), np.asanyarray(upper, dtype=np.float32)   # scalars or (B,) arrays
    diff = np.subtract(b, a)
    out = np.asanyarray(np.add(a, diff * gamma))
    np.subtract(b, diff * (1 - gamma), out=out, where=gamma >= 0.5, casting="unsafe", dtype=type(out.dtype))
    return out[()]


def _quantile_indexes(n: int, q):
    """(virtual index as float32 0-d array, previous index, next index) of np.quantile; -1 means
    "the maximum" (virtual index at or beyond n-1)."""
    q32 = np.asanyarray(q, dtype=np.float32)
    if not (0 <= q32 <= 1):
        raise ValueError("Quantiles must be in the range [0, 1]")
    v = np.asanyarray((n - 1) * q32)
    prev = int(np.floor(v))
    nxt = prev + 1
    if v >= n - 1:
        prev = nxt = -1
    return v, prev, nxt


def gradient_quantiles(dx: torch.Tensor, dy: torch.Tensor, quantiles: Sequence[float]):
    """np.quantile(sqrt(dx**2 + dy**2), q) for every q (utils.py:119, 125-126), float32 results
    identical to NumPy's: a list over q for one image, a list over images of such lists for a
    batch.  Needs |dx|, |dy| <= 4095 (true for Scharr of an 8-bit image) so that
    float32(dx^2 + dy^2) equals the reference's float32 sum of squares."""
    gx, squeeze = _batched(dx, "dx", torch.int16)
    gy, _ = _batched(dy, "dy", torch.int16)
    b, h, w = gx.shape
    n = h * w
    wanted = []          # ranks whose order statistic is needed, per quantile
    for q in quantiles:
        _, prev, nxt = _quantile_indexes(n, q)
        wanted.append((n - 1, n - 1) if prev == -1 else (prev, min(nxt, n - 1)))
    ranks = sorted({r for pair in wanted for r in pair})          # at most 4 for two quantiles: one call
    values = {}                                                   # rank -> (b,) float32 order statistics
    scratch = torch.empty(b * 4 * (1 + 2048), dtype=torch.int32, device=dx.device)
    for i in range(0, len(ranks), 4):
        chunk = ranks[i: i + 4]
        arr = (ctypes.c_int64 * len(chunk))(*chunk)
        out = (ctypes.c_int64 * (b * len(chunk)))()
        _lib.call("mgb_gradient_order_stats", _ptr(gx), _ptr(gy), b, n, arr, len(chunk), out, _ptr(scratch), _stream())
        m = np.sqrt(np.ctypeslib.as_array(out).reshape(b, len(chunk)).astype(np.float32))   # sqrt(float32 sum of squares)
        for j, r in enumerate(chunk):
            values[r] = m[:, j].copy()
    # a virtual index at or beyond n - 1 asked for rank n - 1 twice above, so `hi` is the maximum then;
    # linear_quantile is elementwise, so whole batches go through it at once
    per_q = [np.atleast_1d(linear_quantile(n, q, values[lo], values[hi], values[hi])) for q, (lo, hi) in zip(quantiles, wanted)]
    result = [[col[k] for col in per_q] for k in range(b)]
    return result[0] if squeeze else result


def canny_thresholds(threshold1: float, threshold2: float) -> Tuple[int, int]:
    """The integer (low, high) cv::Canny compares squared magnitudes with when L2gradient=True:
    swap if out of order, clamp to 32767, square when positive, floor."""
    lo, hi = float(threshold1), float(threshold2)
    if lo > hi:
        lo, hi = hi, lo
    lo, hi = min(32767.0, lo), min(32767.0, hi)
    if lo > 0:
        lo *= lo
    if hi > 0:
        hi *= hi
    return math.floor(lo), math.floor(hi)


def canny(dx: torch.Tensor, dy: torch.Tensor, threshold1, threshold2, return_sweeps: bool = False):
    """cv.Canny(dx, dy, threshold1, threshold2, L2gradient=True) != 0 as uint8 0/1 maps
    (utils.py:127-139).  For a batch, threshold1 / threshold2 are sequences (one pair per image)."""
    gx, squeeze = _batched(dx, "dx", torch.int16)
    gy, _ = _batched(dy, "dy", torch.int16)
    b, h, w = gx.shape
    t1 = [threshold1] * b if np.ndim(threshold1) == 0 else list(threshold1)
    t2 = [threshold2] * b if np.ndim(threshold2) == 0 else list(threshold2)
    if len(t1) != b or len(t2) != b:
        raise ValueError("one threshold pair per image")
    thr = np.array([canny_thresholds(a, c) for a, c in zip(t1, t2)], dtype=np.int32).reshape(b, 2)
    thresholds = torch.from_numpy(thr).to(dx.device)
    edges = torch.empty((b, h, w), dtype=torch.uint8, device=dx.device)
    changed = torch.empty(1, dtype=torch.int32, device=dx.device)
    sweeps = ctypes.c_int()
    _lib.call("mgb_canny", _ptr(gx), _ptr(gy), b, h, w, _ptr(thresholds), _ptr(edges), _ptr(changed),
              ctypes.byref(sweeps), _stream())
    edges = edges[0] if squeeze else edges
    return (edges, sweeps.value) if return_sweeps else edges


def find_edges(image: torch.Tensor, low_edge_quantile: float, high_edge_quantile: float):
    """Steps 1-2 of find_circles (utils.py:113-139) for (H, W) or (B, H, W) uint8 images:
    (edges 0/1, dx, dy)."""
    dx, dy = edge_gradients(image)
    q = gradient_quantiles(dx, dy, (low_edge_quantile, high_edge_quantile))
    if image.ndim == 2:
        return canny(dx, dy, q[0], q[1]), dx, dy
    return canny(dx, dy, [v[0] for v in q], [v[1] for v in q]), dx, dy


# ---------------------------------------------------------------------------------------------
# candidates, scores, suppression  (utils.py:141-218, 221-377)
# ---------------------------------------------------------------------------------------------
def circle_perimeter(r: int, four_connected: bool = False) -> np.ndarray:
    """utils.py:433-465 `circle_points`: (n, 2) int32 (drow, dcol) in the reference's order."""
    pts = (ctypes.c_int32 * (40 * r))()
    n = ctypes.c_int()
    _lib.call("mgb_circle_perimeter", int(r), int(bool(four_connected)), pts, 20 * r, ctypes.byref(n))
    return np.ctypeslib.as_array(pts)[: 2 * n.value].reshape(-1, 2).copy()


def perimeter_tables(rmin: int, rmax: int):
    """Concatenated perimeters of radius rmin..rmax for the scoring kernel: offsets (rmax-rmin+2)
    int32, points (P, 2) int16, expected angles (P) float64 = arctan2(drow, dcol) (utils.py:230)."""
    pts = [circle_perimeter(r) for r in range(rmin, rmax + 1)]
    offsets = np.zeros(len(pts) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum([len(p) for p in pts])
    points = np.concatenate(pts).astype(np.int32)
    expected = np.arctan2(points[:, 0], points[:, 1])
    return offsets, points.astype(np.int16), expected.astype(np.float64)


def filter_neighbors(circles: np.ndarray, min_dist: int) -> np.ndarray:
    """utils.py:252-285 on (n, 3) int32 circles sorted best first: boolean keep mask."""
    circles = np.ascontiguousarray(circles, dtype=np.int32)
    valid = np.ones(len(circles), dtype=np.uint8)
    if len(circles):
        _lib.call("mgb_filter_neighbors", circles.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(circles),
                  int(min_dist), valid.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return valid.astype(bool)


class SuppressionNotSettled(RuntimeError):
    """The device suppression met a dependency chain longer than its round limit; use the host pass."""


def conflict_map(min_dist: int) -> np.ndarray:
    """(4 min_dist + 1)^2 uint8: entry (drow + 2 min_dist, dcol + 2 min_dist) is 1 when the rings
    of radius min_dist (utils.py:262, 4-connected raster) around two centres (drow, dcol) apart
    share a pixel -- the pairs utils.py:252-285 treats as neighbours."""
    ring = circle_perimeter(min_dist, four_connected=True)
    reach = 2 * min_dist
    out = np.zeros((2 * reach + 1, 2 * reach + 1), dtype=np.uint8)
    diff = (ring[:, None, :] - ring[None, :, :]).reshape(-1, 2)
    out[diff[:, 0] + reach, diff[:, 1] + reach] = 1
    return out


def filter_neighbors_device(circles: torch.Tensor, batch: int, height: int, width: int, max_radius: int,
                            min_dist: int, return_rounds: bool = False):
    """utils.py:252-285 on the device: boolean keep mask for (N, 4) int32 circles (image, row, col,
    radius) listed image by image, best first.  Only valid when no centre lies more than
    min_dist + 1 pixels above / left of the image (`needs_host_suppression`)."""
    _check(circles, "circles", torch.int32, 2)
    n = circles.shape[0]
    state = torch.empty(n, dtype=torch.uint8, device=circles.device)
    conflict = torch.from_numpy(conflict_map(min_dist)).to(circles.device)
    rounds = ctypes.c_int()
    rc = _lib.try_call("mgb_filter_neighbors_device", _ptr(circles), n, int(batch), int(height), int(width),
                       int(max_radius), int(min_dist), _ptr(conflict), _ptr(state), ctypes.byref(rounds), _stream())
    if rc == _lib.MGB_EUNSUPPORTED:
        raise SuppressionNotSettled(f"{rounds.value} rounds did not settle every circle")
    if rc != 0:
        raise _lib.MagnifyB200Error("mgb_filter_neighbors_device", rc, _lib.error_string(rc))
    keep = state == 1
    return (keep, rounds.value) if return_rounds else keep


def needs_host_suppression(circles: torch.Tensor, min_dist: int) -> bool:
    """True when some centre is so far outside the image that the reference's claim raster is
    indexed below zero (utils.py:270-271 then wrap around like NumPy's negative indices); only the
    host implementation reproduces that."""
    if circles.shape[0] == 0:
        return False
    return bool((circles[:, 1:3] < -(min_dist + 1)).any().item())


class EdgeLists:
    """Edge pixels of a batch of edge maps grouped by grid cell (utils.py:347-377)."""

    def __init__(self, edges: torch.Tensor, grid_length: int):
        e, _ = _batched(edges, "edges", torch.uint8)
        self.b, self.h, self.w = e.shape
        self.grid_length = int(grid_length)
        cells = self.b * math.ceil(self.h / grid_length) * math.ceil(self.w / grid_length)
        self.counts = torch.empty(cells + 1, dtype=torch.int64, device=e.device)
        self.starts = torch.empty(cells + 1, dtype=torch.int64, device=e.device)
        total = ctypes.c_int64()
        args = (_ptr(e), self.b, self.h, self.w, self.grid_length, _ptr(self.counts), _ptr(self.starts))
        _lib.call("mgb_edge_cell_lists", *args, None, 0, ctypes.byref(total), _stream())
        self.total = int(total.value)
        self.coords = torch.empty(max(self.total, 1), dtype=torch.int32, device=e.device)
        if self.total:
            _lib.call("mgb_edge_cell_lists", *args, _ptr(self.coords), self.total, ctypes.byref(total), _stream())

    def grid_coords(self) -> np.ndarray:
        """(total, 2) int32 (row, col) -- `grid_coords` of utils.py:362-375, images concatenated."""
        packed = self.coords[: self.total].cpu().numpy().view(np.uint32)
        return np.stack([packed >> 16, packed & 0xFFFF], axis=1).astype(np.int32)


def sample_circles(lists: EdgeLists, num_iter: int, min_radius=None, max_radius=None, seed: int = 0,
                   randoms: torch.Tensor = None, want_raw: bool = False):
    """utils.py:288-344 (+ :155-165 when a radius window is given).  Returns (raw, circles):
    raw (B, num_iter, 3) float32 draws (None unless want_raw), circles (U, 4) int32 unique rounded
    (image, row, col, radius) (None without a radius window)."""
    dev = lists.coords.device
    total = lists.b * int(num_iter)
    raw = torch.empty((lists.b, num_iter, 3), dtype=torch.float32, device=dev) if want_raw else None
    if randoms is not None:
        if randoms.dtype != torch.int32 or tuple(randoms.shape) != (lists.b, num_iter, 3) or not randoms.is_contiguous():
            raise ValueError("randoms must be a contiguous (B, num_iter, 3) int32 tensor holding uint32 bit patterns")
    window = min_radius is not None
    table = circles = counter = None
    n_unique = ctypes.c_int64()
    cap = 0
    if window and total:
        cap = 1 << max(1, (2 * total - 1).bit_length())
        table = torch.empty(cap, dtype=torch.int64, device=dev)
        circles = torch.empty((total, 4), dtype=torch.int32, device=dev)
        counter = torch.empty(1, dtype=torch.int64, device=dev)
    _lib.call("mgb_sample_circles", _ptr(lists.coords), _ptr(lists.starts), _ptr(lists.counts), lists.b, lists.h, lists.w,
              lists.grid_length, int(num_iter), float(min_radius if window else 0.0), float(max_radius if window else 0.0),
              int(seed) & (2**64 - 1), _ptr(randoms), _ptr(raw), _ptr(table), cap, _ptr(circles),
              ctypes.byref(n_unique), _ptr(counter), _stream())
    if window:
        circles = circles[: n_unique.value] if circles is not None else torch.empty((0, 4), dtype=torch.int32, device=dev)
    return raw, circles


def gradient_angles(dx: torch.Tensor, dy: torch.Tensor, edges: torch.Tensor = None) -> torch.Tensor:
    """float32 arctan2(dy, dx) per pixel (utils.py:169), computed in float64 and rounded once; with
    an edge map only at its edge pixels (0 elsewhere), which is all the score reads."""
    angle = torch.empty(dx.shape, dtype=torch.float32, device=dx.device)
    if edges is not None and (edges.dtype != torch.uint8 or edges.numel() != dx.numel() or not edges.is_contiguous()):
        raise ValueError("edges must be a contiguous uint8 map of the gradients' shape")
    _lib.call("mgb_gradient_angles", _ptr(dx), _ptr(dy), _ptr(edges), dx.numel(), _ptr(angle), _stream())
    return angle


def score_circles(circles: torch.Tensor, edges: torch.Tensor, angle: torch.Tensor, min_radius: int,
                  max_radius: int) -> torch.Tensor:
    """utils.py:181-183 + mean_grad (:221-249): alignment score / perimeter length, float32, for
    (N, 4) int32 circles (image, row, col, radius)."""
    e, _ = _batched(edges, "edges", torch.uint8)
    _check(circles, "circles", torch.int32, 2)
    b, h, w = e.shape
    offsets, points, expected = perimeter_tables(int(min_radius), int(max_radius))
    dev = e.device
    scores = torch.empty(circles.shape[0], dtype=torch.float32, device=dev)
    if circles.shape[0]:
        # keep the uploaded tables referenced until the launch is queued (the caching allocator
        # would hand a released block to the next upload)
        tables = [torch.from_numpy(a).to(dev) for a in (offsets, points, expected)]
        _lib.call("mgb_score_circles", _ptr(circles), circles.shape[0], h, w, _ptr(e), _ptr(angle), int(min_radius),
                  int(max_radius), _ptr(tables[0]), _ptr(tables[1]), _ptr(tables[2]), _ptr(scores), _stream())
    return scores


def select_circles(circles: np.ndarray, scores: np.ndarray, min_roundness: float, min_dist: int):
    """utils.py:187-197 on one image's unique circles (n, 3) int32 and float32 scores: keep
    score >= min_roundness, order best first, suppress neighbours.  Equal scores are ordered by
    (row, col, radius) -- the reference leaves that order to an unstable sort."""
    keep = scores >= min_roundness                     # float32 array vs Python float: NumPy casts the scalar
    circles, scores = circles[keep], scores[keep]
    order = np.lexsort((circles[:, 2], circles[:, 1], circles[:, 0], -scores))
    circles, scores = circles[order], scores[order]
    if min_dist > 0 and len(circles):
        valid = filter_neighbors(circles, min_dist)
        circles, scores = circles[valid], scores[valid]
    return circles, scores


def order_circles(circles: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
    """Device permutation: image by image, best score first, ties in input order (the input of
    `sample_circles` is sorted by (image, row, col, radius))."""
    order = torch.empty(circles.shape[0], dtype=torch.int32, device=circles.device)
    _lib.call("mgb_order_circles", _ptr(circles), _ptr(scores), circles.shape[0], _ptr(order), _stream())
    return order


def find_circles(image: torch.Tensor, low_edge_quantile: float, high_edge_quantile: float, grid_length: int,
                 num_iter: int, min_radius: int, max_radius: int, min_roundness: float, min_dist: int, seed: int = 0):
    """`utils.find_circles` (utils.py:100-218) for a (H, W) uint8 image -> (circles (n,3) int32
    [row, col, radius], scores (n,) float32), best first; for a (B, H, W) batch a list of such
    pairs.  Deterministic for a given seed.  Differences from the reference, both outside what it
    can reproduce itself: the random draws (see csrc/circles_sample.cu) and duplicates -- the
    reference returns a rounded circle once per draw that produced it when min_dist == 0, here
    every circle appears once."""
    edges, dx, dy = find_edges(image, low_edge_quantile, high_edge_quantile)
    e, squeeze = _batched(edges, "edges", torch.uint8)
    b = e.shape[0]
    empty = (np.empty((0, 3), np.int32), np.empty(0, np.float32))
    lists = EdgeLists(e, grid_length)
    results = [empty] * b
    if lists.total and num_iter > 0:
        _, circles = sample_circles(lists, num_iter, min_radius, max_radius, seed)
        if circles.shape[0]:
            angle = gradient_angles(dx, dy, edges)
            scores = score_circles(circles, e, angle, min_radius, max_radius)
            # utils.py:187-188 on the device, so that only survivors travel to the host; a float32
            # tensor compared with a Python float casts the scalar to float32, like NumPy does
            keep = scores >= min_roundness
            circles, scores = circles[keep].contiguous(), scores[keep].contiguous()
            if circles.shape[0]:
                order = order_circles(circles, scores).long()                 # utils.py:192-193
                circles, scores = circles[order].contiguous(), scores[order]
                on_device = min_dist > 0 and not needs_host_suppression(circles, min_dist)
                if on_device:                                                 # utils.py:194-196
                    try:
                        valid = filter_neighbors_device(circles, b, e.shape[1], e.shape[2], max_radius, min_dist)
                        circles, scores = circles[valid], scores[valid]
                    except SuppressionNotSettled:
                        on_device = False                                     # very long chains: sequential host pass
                found, scores = circles.cpu().numpy(), scores.cpu().numpy()
                bounds = np.searchsorted(found[:, 0], np.arange(b + 1))
                for k in range(b):
                    lo, hi = bounds[k], bounds[k + 1]
                    if hi > lo:
                        mine, sc = found[lo:hi, 1:], scores[lo:hi]
                        if min_dist > 0 and not on_device:                    # raster wrap-around case: host
                            valid = filter_neighbors(mine, min_dist)
                            mine, sc = mine[valid], sc[valid]
                        results[k] = (np.ascontiguousarray(mine), sc)
    return results[0] if squeeze else results

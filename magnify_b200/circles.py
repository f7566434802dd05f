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


def to_uint8(x: torch.Tensor) -> torch.Tensor:
    """utils.py:20-27: uint8(255 * (x - min) / (max - min)) in float64, truncated; zeros when the
    array is constant; an empty array stays empty."""
    if not x.is_cuda or not x.is_contiguous():
        raise ValueError("to_uint8 needs a contiguous CUDA tensor")
    if x.dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported dtype {x.dtype}")
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    if x.numel() == 0:
        return out
    minmax = torch.empty(2, dtype=torch.float64, device=x.device)
    _lib.call("mgb_to_uint8", _ptr(x), _DTYPE_CODE[x.dtype], x.numel(), _ptr(out), _ptr(minmax), _stream())
    return out


def edge_gradients(image: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils.py:114-118 on a (H, W) uint8 image: Scharr dx, dy of the 5x5-Gaussian-blurred image as
    int16 (the float32 arrays of the reference hold exactly these integers)."""
    _check(image, "image", torch.uint8, 2)
    h, w = image.shape
    blurred = torch.empty_like(image)
    dx = torch.empty((h, w), dtype=torch.int16, device=image.device)
    dy = torch.empty_like(dx)
    _lib.call("mgb_edge_gradients_u8", _ptr(image), h, w, _ptr(blurred), _ptr(dx), _ptr(dy), _stream())
    return dx, dy


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
    a, b = np.asanyarray(lower, dtype=np.float32), np.asanyarray(upper, dtype=np.float32)
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


def gradient_quantiles(dx: torch.Tensor, dy: torch.Tensor, quantiles: Sequence[float]) -> list:
    """np.quantile(sqrt(dx**2 + dy**2), q) for every q (utils.py:119, 125-126), float32 results
    identical to NumPy's.  Needs |dx|, |dy| <= 4095 (true for Scharr of an 8-bit image) so that
    float32(dx^2 + dy^2) equals the reference's float32 sum of squares."""
    _check(dx, "dx", torch.int16, 2)
    _check(dy, "dy", torch.int16, 2)
    n = dx.numel()
    wanted = []          # ranks whose order statistic is needed, per quantile
    for q in quantiles:
        _, prev, nxt = _quantile_indexes(n, q)
        wanted.append((n - 1, n - 1) if prev == -1 else (prev, min(nxt, n - 1)))
    ranks = sorted({r for pair in wanted for r in pair} | {n - 1})
    values = {}
    scratch = torch.empty(4 * 2048, dtype=torch.int32, device=dx.device)
    for i in range(0, len(ranks), 4):
        chunk = ranks[i: i + 4]
        arr = (ctypes.c_int64 * len(chunk))(*chunk)
        out = (ctypes.c_int64 * len(chunk))()
        _lib.call("mgb_gradient_order_stats", _ptr(dx), _ptr(dy), n, arr, len(chunk), out, _ptr(scratch), _stream())
        for r, m in zip(chunk, out):
            values[r] = np.sqrt(np.float32(m))      # sqrt(float32 sum of exact squares)
    return [linear_quantile(n, q, values[lo], values[hi], values[n - 1]) for q, (lo, hi) in zip(quantiles, wanted)]


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


def canny(dx: torch.Tensor, dy: torch.Tensor, threshold1: float, threshold2: float, return_sweeps: bool = False):
    """cv.Canny(dx, dy, threshold1, threshold2, L2gradient=True) != 0 as a (H, W) uint8 0/1 map
    (utils.py:127-139)."""
    _check(dx, "dx", torch.int16, 2)
    _check(dy, "dy", torch.int16, 2)
    h, w = dx.shape
    low, high = canny_thresholds(threshold1, threshold2)
    work = torch.empty((h, w), dtype=torch.uint8, device=dx.device)
    edges = torch.empty_like(work)
    changed = torch.empty(1, dtype=torch.int32, device=dx.device)
    sweeps = ctypes.c_int()
    _lib.call("mgb_canny", _ptr(dx), _ptr(dy), h, w, low, high, _ptr(work), _ptr(edges), _ptr(changed),
              ctypes.byref(sweeps), _stream())
    return (edges, sweeps.value) if return_sweeps else edges


def find_edges(image: torch.Tensor, low_edge_quantile: float, high_edge_quantile: float):
    """Steps 1-2 of find_circles (utils.py:113-139) for a (H, W) uint8 image: (edges 0/1, dx, dy)."""
    dx, dy = edge_gradients(image)
    low, high = gradient_quantiles(dx, dy, (low_edge_quantile, high_edge_quantile))
    return canny(dx, dy, low, high), dx, dy

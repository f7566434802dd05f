"""Masked per-marker reductions the reference's consumers run on roi/fg/bg.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED by the reference's suite (no
test asserts an intensity).  The reference expressions are xarray's
`roi.where(fg).mean(dim=["roi_x","roi_y"])` / `.median(...)` / `fg.sum(...)`
(src/magnify/identify.py:76-80, filter.py:21-22,51,74,82, README.md:21-22): `where` fills
masked-out pixels with NaN and the reductions skip NaN (nanmean / nanmedian), so an empty
mask yields NaN.  Restated here in float64.
"""
from __future__ import annotations

import warnings

import numpy as np

STATS = ("n_fg", "n_bg", "sum_fg", "sum_bg", "mean_fg", "mean_bg")


def masked_stats(roi: np.ndarray, fg: np.ndarray, bg: np.ndarray) -> np.ndarray:
    """roi (M, C, T, L, L), fg/bg (M, T, L, L) bool -> (M, C, T, 6) float64 in STATS order."""
    m, c, t = roi.shape[:3]
    out = np.empty((m, c, t, len(STATS)), dtype=np.float64)
    r = roi.astype(np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for k, mask in enumerate((fg, bg)):
            mk = np.broadcast_to(mask[:, None], r.shape)
            vals = np.where(mk, r, np.nan)
            out[..., 0 + k] = np.broadcast_to(mask.sum(axis=(-2, -1))[:, None], (m, c, t))
            out[..., 2 + k] = np.nansum(vals, axis=(-2, -1))
            out[..., 4 + k] = np.nanmean(vals, axis=(-2, -1))
    return out


def masked_median(roi: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """`roi.where(mask).median(dim=["roi_x","roi_y"])` -> (M, C, T) float64 (NaN when empty)."""
    r = roi.astype(np.float64)
    vals = np.where(np.broadcast_to(mask[:, None], r.shape), r, np.nan)
    flat = vals.reshape(vals.shape[:3] + (-1,))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmedian(flat, axis=-1)

"""Flat-field correction of the reference, restated in NumPy.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Line-by-line restatement of
src/magnify/preprocess.py:83-87, pinned by tests/golden/flatfield.npz, which holds the outputs of
the reference's own `flatfield_correct` source executed in place on ndarray operands
(oracle/_refload.py::reference_flatfield_correct; the reference's suite has no flat-field test).
"""
from __future__ import annotations

import numpy as np


def flatfield_maxima(tiles: np.ndarray, flatfield=1.0, darkfield=0.0):
    """The two global maxima of preprocess.py:84,86 -> (M, M2) as float64 scalars."""
    t = np.clip(tiles.astype(float) - darkfield, 0, None)  # :83
    m = t.max()  # :84  (global: every channel, time and tile, margins included)
    t = t / flatfield  # :85
    return float(m), float(t.max())  # :86 evaluates tiles.max() on the divided array


def flatfield_correct(tiles: np.ndarray, flatfield=1.0, darkfield=0.0, maxima=None) -> np.ndarray:
    """src/magnify/preprocess.py:83-87 on a (channel, time, tile_row, tile_col, tile_y, tile_x)
    array.  `flatfield` / `darkfield` are scalars or arrays broadcast on the trailing dims
    exactly as NumPy/xarray do for a bare ndarray operand (:75-81 read a 2-D TIFF).

    maxima: optional (M, M2) override -- the multi-GPU path all-reduces them, and chunked
    callers (bench CPU baseline) compute them in a first sweep.  The arithmetic order is the
    reference's: ((t / flat) * M) / M2, all in float64, then a C cast to the tile dtype.
    """
    t = np.clip(tiles.astype(float) - darkfield, 0, None)  # :83
    max_val = t.max() if maxima is None else maxima[0]  # :84
    t = t / flatfield  # :85
    second = t.max() if maxima is None else maxima[1]
    t = t * max_val / second  # :86  == (t * M) / M2, left to right
    return t.astype(tiles.dtype)  # :87  truncation toward zero for integer dtypes

"""Tile stitching of the reference, restated in NumPy.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows src/magnify/stitch.py:7-46; pinned by
the exact-slice assertions of the reference's tests/test_stitch.py (ported in
tests/test_oracle_paths.py) and by tests/golden/stitch.npz, the outputs of the reference's own
Stitcher source executed in place (oracle/_refload.py::reference_stitch).
"""
from __future__ import annotations

import numpy as np


def check_overlap(overlap: int, tile_y: int | None = None, tile_x: int | None = None) -> None:
    """Error behaviour of stitch.py:8-9 (constructor) and :16-20 (call)."""
    if overlap < 0:
        raise ValueError("Overlap must be non-negative.")
    if tile_y is not None and (overlap >= tile_y or overlap >= tile_x):
        raise ValueError(
            f"Overlap ({overlap}) must be smaller than tile size ({tile_y}x{tile_x})."
        )


def stitch(tiles: np.ndarray, overlap: int = 102) -> np.ndarray:
    """(C, T, R, Cc, H, W) -> (C, T, R*(H-ov), Cc*(W-ov)), stitch.py:22-39.

    Each tile keeps [ov//2 : size - ov//2 - ov%2] on both axes (:23-30); tile rows are joined
    along y and tile columns along x (:33-35).
    """
    c, t, rows, cols, h, w = tiles.shape
    check_overlap(overlap, h, w)
    clip, rem = overlap // 2, overlap % 2
    kept = tiles[..., clip : h - clip - rem, clip : w - clip - rem]
    kh, kw = kept.shape[-2:]
    return np.ascontiguousarray(kept.transpose(0, 1, 2, 4, 3, 5)).reshape(c, t, rows * kh, cols * kw)

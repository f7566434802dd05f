"""ROI gather and fg/bg masks of the reference's finders, restated in NumPy.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Marker centres are INPUTS: centre finding
(find.py:476-501, 205-306, 339-360) is stochastic, unseeded CPU code and out of scope.
"""
from __future__ import annotations

import numpy as np

from . import geometry as g


def gather_rois(image: np.ndarray, x: np.ndarray, y: np.ndarray, roi_length: int) -> np.ndarray:
    """roi[m, c, t] = image[c, t, top:bottom, left:right] with the box of
    bounding_box(round(x[m, t]), round(y[m, t]), L, im_x, im_y).

    One statement of the three crop loops: beads find.py:589-602 (x, y constant over time),
    chip search timestep find.py:324-334,370-377 and chip copy-forward find.py:160-169.
    image: (C, T, H, W); x, y: (M, T) float64 -> (M, C, T, L, L).
    """
    c, t, h, w = image.shape
    m = x.shape[0]
    roi = np.empty((m, c, t, roi_length, roi_length), dtype=image.dtype)
    for ti in range(t):
        for mi in range(m):
            top, bottom, left, right = g.bounding_box(
                round(float(x[mi, ti])), round(float(y[mi, ti])), roi_length, w, h
            )
            roi[mi, :, ti] = image[:, ti, top:bottom, left:right]
    return roi


def bead_masks(beads: np.ndarray, im_y: int, im_x: int, roi_length: int):
    """fg/bg of BeadFinder, find.py:561-584.

    beads: (M, 3) rows (row, col, radius) as returned by find_circles (integer valued).
    Returns labels (im_y, im_x) int32, fg, bg (M, L, L) bool; the reference then broadcasts
    them over time (find.py:585-586).
    """
    beads = np.asarray(beads)
    labels = g.circle_labels(beads.astype(int), im_y, im_x)  # :561
    m = len(beads)
    fg = np.empty((m, roi_length, roi_length), dtype=bool)
    bg = np.empty_like(fg)
    for i in range(m):
        # x = beads[:, 1], y = beads[:, 0]  (find.py:543-550)
        top, bottom, left, right = g.bounding_box(
            round(float(beads[i, 1])), round(float(beads[i, 0])), roi_length, im_x, im_y
        )
        sub = labels[top:bottom, left:right]
        fg[i] = sub == i  # :582
        bg[i] = sub == -1  # :584
    return labels, fg, bg


def chip_masks(x, y, fg_radius, roi_length: int, chamber_radius: int, max_button_radius: int,
               im_x: int, im_y: int):
    """fg/bg of ButtonFinder.find_rois at a search timestep, find.py:380-400.

    x, y: (M,) final (possibly refined) centres in image coordinates, float64;
    fg_radius: (M,) int -- the refined radius, or max_button_radius when refinement found
    nothing (find.py:363,378).  fg = disc(fg_radius); bg = annulus(chamber_radius,
    max_button_radius), both centred on (round(y) - top, round(x) - left).
    """
    m = len(x)
    fg = np.empty((m, roi_length, roi_length), dtype=bool)
    bg = np.empty_like(fg)
    for i in range(m):
        xr, yr = round(float(x[i])), round(float(y[i]))
        top, _, left, _ = g.bounding_box(xr, yr, roi_length, im_x, im_y)
        centre = (yr - top, xr - left)
        bg[i] = g.annulus((roi_length, roi_length), centre, chamber_radius, max_button_radius)
        fg[i] = g.circle((roi_length, roi_length), centre, int(fg_radius[i]))
    return fg, bg


def chip_copy_forward(num_times: int, search_timesteps) -> np.ndarray:
    """Source timestep of every timestep's centres/masks, find.py:143-151.

    Search timesteps map to themselves; any other t copies from search_timesteps[0] when it
    precedes the first search, else from t-1 (which already holds its own source), so the
    result is the latest search timestep <= t, or the first one.
    """
    search = sorted(int(s) for s in np.atleast_1d(search_timesteps))
    src = np.empty(num_times, dtype=np.int64)
    for t in range(num_times):
        if t in search:
            src[t] = t
        elif t < search[0]:
            src[t] = search[0]
        else:
            src[t] = src[t - 1]
    return src

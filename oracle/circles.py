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

"""TEST INFRASTRUCTURE: an oracle-backed stand-in for `magnify_b200.ops` on CPU torch tensors.

The build container has the reference but no GPU; the GPU box has a GPU but no reference.  To
run the component layer (dataset schema, lazy hand-off, registration, the reference's own
`Pipeline.__call__` and `*_pipe` builders) against the reference HERE, the array kernels are
replaced by the NumPy oracle for the duration of a test (`patch_components`).  The kernels
themselves are proven equal to the same oracle by the `-m gpu` parity tests, and the component
layer runs with the real kernels in tests/test_gpu_dropin.py against goldens made here.  Nothing
in the product imports this module.
"""
from __future__ import annotations

import types

import numpy as np
import torch

from oracle import flatfield as o_ff
from oracle import geometry as o_geo
from oracle import reduce as o_red
from oracle import stitch as o_st

NSTATS = 8
STATS = o_red.STATS


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _t(a, like=None):
    return torch.from_numpy(np.ascontiguousarray(a))


check_overlap = o_st.check_overlap


def stitched_shape(tile_shape, overlap):
    c, t, r, cc, h, w = tile_shape
    return (c, t, r * (h - overlap), cc * (w - overlap))


def alloc_image(shape, dtype, device):
    return torch.empty(tuple(int(s) for s in shape), dtype=dtype)


def image_pitch(image):
    if image.dim() != 4:
        raise ValueError("image must be 4-d")
    return int(image.shape[-1])


def to_host_dense(image, out=None, non_blocking=True):
    if out is None:
        return image.clone()
    out.copy_(image)
    return out


class FlatFieldPlan:
    def __init__(self, tile_shape, flatfield=1.0, darkfield=0.0, device="cpu"):
        self.tile_shape = tuple(tile_shape)
        self.flat, self.dark = flatfield, darkfield
        self.identity = np.isscalar(flatfield) and np.isscalar(darkfield) and float(flatfield) == 1.0 and float(darkfield) == 0.0
        self.maxima = torch.zeros(2, dtype=torch.float64)


def flatfield_maxima(tiles, plan):
    m = o_ff.flatfield_maxima(_np(tiles), plan.flat, plan.dark)
    plan.maxima[:] = torch.tensor([float(m[0]), float(m[1])], dtype=torch.float64)
    return plan.maxima


def flatfield_maxima_accumulate(block, plan, channel):
    flat, dark = plan.flat, plan.dark
    if not np.isscalar(flat) and np.asarray(flat).ndim == 6:
        flat = np.asarray(flat)[channel]
    if not np.isscalar(dark) and np.asarray(dark).ndim == 6:
        dark = np.asarray(dark)[channel]
    m = o_ff.flatfield_maxima(_np(block), flat, dark)
    plan.maxima[0] = max(float(plan.maxima[0]), float(m[0]))
    plan.maxima[1] = max(float(plan.maxima[1]), float(m[1]))


def stitch(tiles, overlap=102, out=None):
    return _t(o_st.stitch(_np(tiles), overlap))


def flatfield_correct(tiles, flatfield=1.0, darkfield=0.0, plan=None, maxima=None, group=None):
    if plan is not None:
        flatfield, darkfield = plan.flat, plan.dark
    mx = None if maxima is None else (float(maxima[0]), float(maxima[1]))
    return _t(o_ff.flatfield_correct(_np(tiles), flatfield, darkfield, maxima=mx))


def flatfield_stitch(tiles, flatfield=1.0, darkfield=0.0, overlap=102, plan=None, maxima=None, group=None, out=None):
    return stitch(flatfield_correct(tiles, flatfield, darkfield, plan=plan, maxima=maxima), overlap)


def bounding_boxes(x, y, roi_length, im_x, im_y, want_rel=False):
    xn, yn = _np(x), _np(y)
    if im_x < roi_length or im_y < roi_length:
        raise ValueError("image smaller than roi_length")
    boxes = o_geo.boxes_from_centres(xn, yn, roi_length, im_x, im_y).astype(np.int32)
    if not want_rel:
        return _t(boxes)
    rel = np.stack([o_geo.round_half_even(yn) - boxes[..., 0], o_geo.round_half_even(xn) - boxes[..., 1]], -1)
    return _t(boxes), _t(rel.astype(np.int32))


def chip_masks(rel, fg_radius, inner_radius, outer_radius, roi_length, want_counts=False):
    rel, rad = _np(rel), _np(fg_radius)
    m = rel.shape[0]
    fg = np.zeros((m, roi_length, roi_length), np.uint8)
    bg = np.zeros_like(fg)
    for i in range(m):
        centre = (int(rel[i, 0]), int(rel[i, 1]))
        fg[i] = o_geo.circle((roi_length, roi_length), centre, int(rad[i]))
        bg[i] = o_geo.annulus((roi_length, roi_length), centre, int(outer_radius), int(inner_radius))
    return _t(fg), _t(bg)


def disc_halfwidth_table(rmax):
    table = np.zeros((int(rmax) + 1, int(rmax) + 1), dtype=np.int32)
    for r in range(1, int(rmax) + 1):
        table[r, : r + 1] = o_geo.disc_halfwidths(r)
    return table


def bead_labels(beads, im_y, im_x, device=None):
    return _t(o_geo.circle_labels(_np(beads).astype(np.int64), im_y, im_x).astype(np.int32))


def bead_masks(labels, boxes, roi_length, want_counts=False):
    lab, bx = _np(labels), _np(boxes)
    m = bx.shape[0]
    fg = np.zeros((m, roi_length, roi_length), np.uint8)
    bg = np.zeros_like(fg)
    for i in range(m):
        sub = lab[bx[i, 0]:bx[i, 0] + roi_length, bx[i, 1]:bx[i, 1] + roi_length]
        fg[i] = sub == i
        bg[i] = sub == -1
    return _t(fg), _t(bg)


def roi_gather(image, boxes, roi_length, out=None, order=None):
    img, bx = _np(image), _np(boxes)
    c, t = img.shape[:2]
    m = bx.shape[0]
    roi = np.empty((m, c, t, roi_length, roi_length), dtype=img.dtype)
    for i in range(m):
        for ti in range(t):
            top, left = bx[i, ti]
            roi[i, :, ti] = img[:, ti, top:top + roi_length, left:left + roi_length]
    return _t(roi)


def roi_stats(roi, fg, bg, mask_t=None, medians=True):
    r, f, b = _np(roi), _np(fg).astype(bool), _np(bg).astype(bool)
    t = r.shape[2]
    if mask_t is None:
        mt = np.zeros(t, dtype=np.int64) if f.shape[1] == 1 else np.arange(t)
    else:
        mt = _np(mask_t).astype(np.int64)
    stats = o_red.masked_stats(r, f[:, mt], b[:, mt])
    if not medians:
        stats[..., 6:] = np.nan
    return _t(stats)


def roi_gather_stats(image, boxes, fg, bg, roi_length, mask_t=None, want_roi=True, out_roi=None, out_stats=None,
                     order=None, peer_stats=None, medians=True, mask_counts=None):
    roi = roi_gather(image, boxes, roi_length)
    return (roi if want_roi else None), roi_stats(roi, fg, bg, mask_t, medians)


def roi_median(roi, mask, mask_t=None):
    r, mk = _np(roi), _np(mask).astype(bool)
    t = r.shape[2]
    mt = (np.zeros(t, dtype=np.int64) if mk.shape[1] == 1 else np.arange(t)) if mask_t is None else _np(mask_t).astype(np.int64)
    return _t(o_red.masked_median(r, mk[:, mt]))


def mask_perimeters(masks):
    import cv2 as cv

    out = []
    for mk in _np(masks):
        contours, _ = cv.findContours((mk != 0).astype(np.uint8) * 255, cv.RETR_EXTERNAL, cv.CHAIN_APPROX_SIMPLE)
        out.append(sum(cv.arcLength(c, True) for c in contours))
    return _t(np.asarray(out, dtype=np.float64))


def _stage_tiles(src, dev, ff):
    """CPU version of components._stage_tiles: no streams, the whole stack at once."""
    from magnify_b200.devarray import DeviceArray

    if isinstance(src, DeviceArray):
        tiles = src.tensor
    else:
        tiles = torch.from_numpy(np.ascontiguousarray(np.asarray(src)))
    if ff is not None and not ff.identity:
        flatfield_maxima(tiles, ff)
    return tiles


def patch_components(monkeypatch):
    """Route the component layer to this module (CPU tensors, oracle arithmetic) for one test."""
    from magnify_b200 import components

    this = types.SimpleNamespace(**{k: v for k, v in globals().items() if not k.startswith("__")})
    monkeypatch.setattr(components, "ops", this)
    monkeypatch.setattr(components, "_device", lambda device: torch.device("cpu"))
    monkeypatch.setattr(components, "_stage_tiles", _stage_tiles)
    monkeypatch.setattr(components, "_upload", lambda values, dev: torch.from_numpy(np.ascontiguousarray(values)))

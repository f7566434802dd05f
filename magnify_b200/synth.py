"""Deterministic synthetic inputs shaped like BASELINE.json's configs.

There is no network and no dataset: tiles are background noise (N(400, 20)) plus filled discs on
the marker grid, generated on the device with torch (plumbing, not the measured path).  Marker
centres are OUTPUTS of the generator and INPUTS of the hot path -- centre finding is out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class ChipCase:
    tiles: torch.Tensor            # (C,T,R,Cc,H,W) uint16 on device
    overlap: int
    roi_length: int
    chamber_radius: int
    max_button_radius: int
    x: np.ndarray                  # (M,T) float64 centres in stitched-image coordinates
    y: np.ndarray
    fg_radius: np.ndarray          # (M,1) int32
    flat: np.ndarray               # (H,W) float64, strictly positive
    dark: np.ndarray               # (H,W) float64
    grid: tuple


def smooth_flat_dark(h: int, w: int):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    flat = 1.0 + 0.4 * np.cos((yy / h - 0.45) * 2.2) * np.cos((xx / w - 0.55) * 2.0) - 0.2
    flat = np.clip(flat, 0.6, 1.4)
    dark = 100.0 + 5.0 * np.sin(yy * 0.013 + 0.5) * np.cos(xx * 0.017)
    return flat, dark


def chip_case(c=2, t=1, r=4, cc=4, h=2048, w=2048, overlap=102, rows=56, cols=32, row_dist=126.1,
              col_dist=232.9, roi_length=72, chamber_radius=30, max_button_radius=15, seed=0,
              device="cuda") -> ChipCase:
    """BASELINE configs 2/3: R x Cc tiles per (channel, time), a rows x cols button grid at the
    'pc' chip spacing (registry.py:234-235), discs of radius 10-15 on noise."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + seed)
    clip = overlap // 2
    kh, kw = h - overlap, w - overlap
    him, wim = r * kh, cc * kw
    rng = np.random.default_rng(seed)
    y0 = max(roi_length, (him - (rows - 1) * row_dist) / 2)
    x0 = max(roi_length, (wim - (cols - 1) * col_dist) / 2)
    jit_y = rng.uniform(-2, 2, (rows, cols))
    jit_x = rng.uniform(-2, 2, (rows, cols))
    cy = y0 + np.arange(rows)[:, None] * row_dist + jit_y
    cx = x0 + np.arange(cols)[None, :] * col_dist + jit_x
    radius = 10 + (np.add.outer(np.arange(rows), np.arange(cols)) % 6)
    cy_d = torch.from_numpy(cy).to(dev, torch.float32)
    cx_d = torch.from_numpy(cx).to(dev, torch.float32)
    r2_d = torch.from_numpy((radius ** 2).astype(np.float32)).to(dev)
    tiles = torch.empty((c, t, r, cc, h, w), dtype=torch.uint16, device=dev)
    ty = torch.arange(h, device=dev, dtype=torch.float32)
    tx = torch.arange(w, device=dev, dtype=torch.float32)
    for ri in range(r):
        gy = ri * kh + (ty - clip)
        iy = torch.clamp(torch.round((gy - y0) / row_dist), 0, rows - 1).long()
        for ci in range(cc):
            gx = ci * kw + (tx - clip)
            ix = torch.clamp(torch.round((gx - x0) / col_dist), 0, cols - 1).long()
            yy = iy[:, None].expand(h, w)
            xx = ix[None, :].expand(h, w)
            d2 = (gy[:, None] - cy_d[yy, xx]) ** 2 + (gx[None, :] - cx_d[yy, xx]) ** 2
            disc = (d2 <= r2_d[yy, xx]).to(torch.float32)
            for ch in range(c):
                for ti in range(t):
                    noise = torch.randn((h, w), generator=gen, device=dev, dtype=torch.float32) * 20.0 + 400.0
                    amp = 3000.0 * (1 + ch) * (1.0 + 0.01 * ti)
                    tiles[ch, ti, ri, ci] = torch.clamp(noise + amp * disc, 0, 65535).to(torch.int32).to(torch.uint16)
    m = rows * cols
    x = np.repeat(cx.reshape(m, 1), t, axis=1)
    y = np.repeat(cy.reshape(m, 1), t, axis=1)
    flat, dark = smooth_flat_dark(h, w)
    return ChipCase(tiles, overlap, roi_length, chamber_radius, max_button_radius, x, y,
                    radius.reshape(m, 1).astype(np.int32), flat, dark, (rows, cols))


@dataclass
class BeadCase:
    tiles: torch.Tensor            # (C,T,R,Cc,H,W) uint16 on device
    overlap: int
    roi_length: int
    beads: np.ndarray              # (M,3) float64 rows (row, col, radius), integer valued
    flat: np.ndarray
    dark: np.ndarray


def bead_case(c=4, t=1, r=1, cc=1, h=2048, w=2048, overlap=0, n_beads=300, min_radius=10, max_radius=25,
              roi_length=100, seed=0, device="cuda") -> BeadCase:
    """BASELINE configs 1/5: beads scattered over the stitched image, 5% deliberately overlapping
    so that the label raster's -2 (shared) case is exercised; noise everywhere (the masks and
    crops do not depend on the pixel values)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4321 + seed)
    kh, kw = h - overlap, w - overlap
    him, wim = r * kh, cc * kw
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, him, n_beads)
    cols = rng.integers(0, wim, n_beads)
    rad = rng.integers(min_radius, max_radius + 1, n_beads)
    n_pair = n_beads // 20
    if n_pair:
        rows[-n_pair:] = np.clip(rows[:n_pair] + rng.integers(-8, 9, n_pair), 0, him - 1)
        cols[-n_pair:] = np.clip(cols[:n_pair] + rad[:n_pair], 0, wim - 1)
    beads = np.stack([rows, cols, rad], 1).astype(np.float64)
    tiles = torch.empty((c, t, r, cc, h, w), dtype=torch.uint16, device=dev)
    flat_view = tiles.view(-1, h, w)
    for i in range(flat_view.shape[0]):
        noise = torch.randn((h, w), generator=gen, device=dev, dtype=torch.float32) * 20.0 + 400.0
        flat_view[i] = torch.clamp(noise, 0, 65535).to(torch.int32).to(torch.uint16)
    flat, dark = smooth_flat_dark(h, w)
    return BeadCase(tiles, overlap, roi_length, beads, flat, dark)

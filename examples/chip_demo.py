"""End-to-end demo on synthetic data: writes a small chip experiment as TIFF tiles
(`chip_<channel>_<time>_<row>_<col>.tif`), then chains the GPU components by hand in the order
`mg.microfluidic_chip_pipe` uses (registry.py:243-269) -- read, stitch, button finding, crops /
masks / summaries, expression filter -- and prints what it found and how long each part took.
With magnify installed the same components run inside its own pipeline after
`magnify_b200.install()`.

    python examples/chip_demo.py [workdir]
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from magnify_b200 import reader
from magnify_b200.components import ButtonFinder, Stitcher, filter_expression, quantify
from magnify_b200.reader import write_tiff


def synthetic_chip(rows=12, cols=8, row_dist=126, col_dist=233, channels=("bf", "egfp"), times=3, seed=0):
    """(C, T, H, W) uint16: discs of radius 10-15 on a flat background, every 7th chamber empty."""
    rng = np.random.default_rng(seed)
    h, w = (rows + 1) * row_dist, (cols + 1) * col_dist
    image = np.full((len(channels), times, h, w), 400, dtype=np.uint16)
    yy, xx = np.mgrid[-16:17, -16:17]
    blanks = []
    for i in range(rows):
        for j in range(cols):
            if (i * cols + j) % 7 == 3:
                blanks.append((i, j))
                continue
            cy, cx, r = (i + 1) * row_dist + rng.integers(-2, 3), (j + 1) * col_dist + rng.integers(-2, 3), 10 + (i + j) % 6
            disc = yy * yy + xx * xx <= r * r
            for c in range(len(channels)):
                for t in range(times):
                    image[c, t, cy - 16:cy + 17, cx - 16:cx + 17][disc] = 3000 * (c + 1) + 200 * t + 11 * (i * cols + j)
    return image, blanks


def split_into_tiles(image, rows, cols, overlap):
    h, w = image.shape[0] // rows, image.shape[1] // cols
    clip, rem = overlap // 2, overlap % 2
    padded = np.pad(image, ((clip, clip + rem), (clip, clip + rem)), mode="reflect")
    return [[padded[i * h:(i + 1) * h + overlap, j * w:(j + 1) * w + overlap] for j in range(cols)] for i in range(rows)]


def main():
    workdir = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="mgb_demo_")
    rows, cols, overlap = 12, 8, 20
    image, blanks = synthetic_chip(rows, cols)
    channels = ("bf", "egfp")
    t0 = time.perf_counter()
    for c, name in enumerate(channels):
        for t in range(image.shape[1]):
            tiles = split_into_tiles(image[c, t, : image.shape[2] // 2 * 2, : image.shape[3] // 2 * 2], 2, 2, overlap)
            for i in range(2):
                for j in range(2):
                    write_tiff(os.path.join(workdir, f"chip_{name}_2024010{t + 1}-120000_{i}_{j}.tif"), tiles[i][j])
    print(f"wrote {len(channels) * image.shape[1] * 4} TIFF tiles to {workdir} in {time.perf_counter() - t0:.2f} s")

    tags = np.full((rows, cols), "sample", dtype="<U16")
    for b in blanks:
        tags[b] = ""
    t0 = time.perf_counter()
    (xp,) = [reader.standardize_format(x) for x in
             reader.Reader()(os.path.join(workdir, "chip_(channel)_(time)_(row)_(col).tif"))]
    xp = xp.assign_coords(tag=(("mark_row", "mark_col"), tags),                      # identify_buttons, identify.py:36-45
                          valid=(("mark_row", "mark_col", "time"), np.ones(tags.shape + (xp.sizes["time"],), bool)))
    xp = Stitcher(overlap=overlap)(xp)
    xp = ButtonFinder(row_dist=126, col_dist=233, min_button_diameter=16, max_button_diameter=34, chamber_diameter=60,
                      search_channel="egfp", num_iter=200000)(xp)
    t1 = time.perf_counter()
    print(f"read + stitch + find buttons + crops/masks: {t1 - t0:.2f} s; roi {xp.roi.shape} {xp.roi.dtype}")
    x0, y0 = xp.x.values[:, 0].reshape(rows, cols), xp.y.values[:, 0].reshape(rows, cols)
    err = max(np.abs(x0 - (np.arange(cols)[None, :] + 1) * 233).max(), np.abs(y0 - (np.arange(rows)[:, None] + 1) * 126).max())
    print(f"largest centre offset from the nominal grid: {err:.1f} px (the discs are jittered by +-2 px)")

    # summaries (already computed by the gather) and the expression filter
    xp = filter_expression(quantify(xp), search_channel="egfp")
    expressed = xp.valid.values[:, 0].reshape(rows, cols)
    print(f"expressed chambers: {int(expressed.sum())} of {rows * cols} ({len(blanks)} were left empty)")
    print("mean fg intensity (egfp, t=0) of the first row:", np.round(xp.fg_mean.values[:cols, 1, 0]).astype(int).tolist())


if __name__ == "__main__":
    main()

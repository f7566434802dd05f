"""The chip pipeline end to end at the component level, the way `mg.microfluidic_chip` chains it
(registry.py:243-269): TIFF tiles on disk -> read -> standardize_format -> (tags) -> flat-field +
stitch -> find_buttons (centres found on the GPU) -> quantify -> filter_expression."""
import os

import numpy as np
import pytest

from test_gpu_finders import draw_chip
from tiffgen import write_tiff

pytestmark = pytest.mark.gpu


def split_into_tiles(image, rows, cols, overlap):
    """Inverse of stitch.py:22-39: tiles of (h + overlap) x (w + overlap) whose kept centres tile the image."""
    h, w = image.shape[0] // rows, image.shape[1] // cols
    clip, rem = overlap // 2, overlap % 2
    padded = np.pad(image, ((clip, clip + rem), (clip, clip + rem)), mode="reflect")
    return np.stack([np.stack([padded[i * h: (i + 1) * h + overlap, j * w: (j + 1) * w + overlap] for j in range(cols)])
                     for i in range(rows)])


def test_chip_pipeline_from_tiff_tiles(cuda_device, tmp_path):
    from magnify_b200 import reader
    from magnify_b200.components import ButtonFinder, FlatfieldStitcher, filter_expression, quantify

    shape, overlap, t = (6, 4), 10, 2
    chip = draw_chip(shape, 20, row_dist=100, col_dist=100, value=1000, blanks=((1, 2),))     # (700, 500)
    chip = chip + 100
    rng = np.random.default_rng(0)
    frames = [chip + rng.integers(0, 20, chip.shape).astype(np.uint16) for _ in range(t)]
    for ti, frame in enumerate(frames):
        tiles = split_into_tiles(frame, 2, 2, overlap)
        for i in range(2):
            for j in range(2):
                write_tiff(os.path.join(tmp_path, f"chip_egfp_2024010{ti + 1}-120000_{i}_{j}.tif"), [tiles[i, j]],
                           rows_per_strip=50)
    (xp,) = list(reader.Reader()(os.path.join(tmp_path, "chip_(channel)_(time)_(row)_(col).tif")))
    xp = reader.standardize_format(xp)
    assert xp["tile"].shape == (1, t, 2, 2, 350 + overlap, 250 + overlap)
    tag = np.full(shape, "default", dtype="<U200")                                          # identify.py:30-32
    tag[1, 2] = ""
    xp = xp.assign_coords(tag=(("mark_row", "mark_col"), tag),
                          valid=(("mark_row", "mark_col", "time"), np.ones(shape + (t,), bool)))
    xp = FlatfieldStitcher(1.0, 0.0, overlap, device=cuda_device)(xp)
    np.testing.assert_array_equal(xp.image.values[0], np.stack(frames))                      # identity flat-field
    xp = ButtonFinder(row_dist=100, col_dist=100, min_button_diameter=16, max_button_diameter=32, chamber_diameter=60,
                      num_iter=20000, min_roundness=0.2, cluster_penalty=50, device=cuda_device)(xp)
    assert xp.sizes["mark"] == 24
    x, y = xp.x.values.reshape(6, 4, t), xp.y.values.reshape(6, 4, t)
    for i in range(6):
        for j in range(4):
            assert abs(x[i, j, 0] - (j + 1) * 100) <= 3 and abs(y[i, j, 0] - (i + 1) * 100) <= 3
    xp = quantify(xp, device=cuda_device)
    filled = (tag != "").reshape(-1)
    fg_mean = xp.fg_mean.values[:, 0, :]
    assert np.all(fg_mean[filled] > 1050) and np.all(xp.bg_mean.values[filled, 0, :] < 130)
    assert np.all(fg_mean[~filled] < 130)                                                    # the blank chamber
    xp = filter_expression(xp, device=cuda_device)
    valid = xp.valid.values
    assert valid[filled].all() and not valid[~filled].any()

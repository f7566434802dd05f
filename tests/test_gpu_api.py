"""`magnify_b200.api.beads` / `microfluidic_chip`: the reference's top-level calls (registry.py)
with the reference's own test scenarios (tests/test_beads.py, tests/test_chip.py) and their
output-structure checks (dims, coords, squeezed axes)."""
import os

import numpy as np
import pytest

from test_gpu_finders import draw_beads, draw_chip
from tiffgen import write_tiff

pytestmark = pytest.mark.gpu


def test_beads_call_and_structure(cuda_device):                                   # tests/test_beads.py:50-66, 250-274
    from magnify_b200 import api

    xp = api.beads(data=draw_beads((1024, 1024), [512, 512]), dims=("y", "x"), min_bead_diameter=16,
                   max_bead_diameter=24, overlap=0, num_iter=100)
    assert xp.roi.sizes["mark"] == 1
    assert 0.95 * 10 < np.sqrt(xp.fg.values.sum() / np.pi) < 1.05 * 10
    assert 0.95 * 512 < xp.x.values.item() < 1.05 * 512 and 0.95 * 512 < xp.y.values.item() < 1.05 * 512
    for name in ("x", "y", "fg", "bg"):
        assert name in xp.coords
    assert "roi" in xp.data_vars and "tile" not in xp and xp.roi.dims == ("mark", "roi_y", "roi_x")
    assert xp.image.dims == ("im_y", "im_x") and xp.fg.dims == ("mark", "roi_y", "roi_x")


def test_beads_multichannel_array(cuda_device):                                   # tests/test_beads.py:282-330
    from magnify_b200 import api

    a = draw_beads((512, 512), [[100, 100], [300, 300]])
    b = draw_beads((512, 512), [[100, 100], [400, 150]])
    xp = api.beads(data=np.stack([a, b]), dims=("channel", "y", "x"), coords={"channel": ["bf", "gfp"]},
                   min_bead_diameter=16, max_bead_diameter=24, overlap=0, num_iter=10000, search_channel="gfp")
    assert xp.roi.sizes["mark"] == 2 and xp.roi.dims == ("mark", "channel", "roi_y", "roi_x")
    assert list(xp.channel.values) == ["bf", "gfp"]


def test_chip_call_and_structure(cuda_device):                                    # tests/test_chip.py:99-127, 319-370
    from magnify_b200 import api

    xp = api.microfluidic_chip(data=draw_chip((10, 10), 20), dims=("y", "x"), shape=(10, 10), min_button_diameter=16,
                               max_button_diameter=32, overlap=0, row_dist=100, col_dist=100, num_iter=10000)
    assert xp.roi.sizes["mark_row"] == 10 and xp.roi.sizes["mark_col"] == 10
    radii = np.sqrt(xp.fg.values.sum(axis=(-1, -2)) / np.pi)
    assert 9 < radii.min() and radii.max() < 11
    assert 95 < xp.x.values[0, 0] < 105 and 395 < xp.x.values[4, 3] < 405 and 495 < xp.y.values[4, 3] < 505
    assert xp.roi.dims == ("mark_row", "mark_col", "roi_y", "roi_x") and xp.tag.dims == ("mark_row", "mark_col")
    assert xp.valid.dims == ("mark_row", "mark_col") and xp.valid.values.all()
    with pytest.raises(ValueError):
        api.microfluidic_chip(data=draw_chip((2, 2), 20), dims=("y", "x"), shape=(2, 2), chip_type="nope")


def test_chip_time_series_from_tiff_pattern(cuda_device, tmp_path):               # tests/test_chip.py:375-424 via files
    from magnify_b200 import api

    chip = draw_chip((3, 3), 20)
    for t in range(3):
        write_tiff(os.path.join(tmp_path, f"chip_2024010{t + 1}-120000.tif"), [chip])
    xp = api.microfluidic_chip(data=os.path.join(tmp_path, "chip_(time).tif"), shape=(3, 3), min_button_diameter=16,
                               max_button_diameter=32, overlap=0, row_dist=100, col_dist=100, num_iter=5000)
    assert xp.roi.dims == ("mark_row", "mark_col", "time", "roi_y", "roi_x") and xp.roi.sizes["time"] == 3
    for t in range(3):
        for r in range(3):
            for c in range(3):
                assert abs(xp.x.values[r, c, t] - (c + 1) * 100) < 5 and abs(xp.y.values[r, c, t] - (r + 1) * 100) < 5

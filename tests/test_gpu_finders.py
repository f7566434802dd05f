"""End-to-end finder tests ported from the reference's own suite (tests/test_beads.py,
tests/test_chip.py) to the component level: synthetic beads / chips drawn with the reference's
disc raster, centres found on the GPU, and the reference's tolerance assertions (positions within
5 %, sqrt(area / pi) within 5-10 % of the true radius, grid positions, copy-forward)."""
import numpy as np
import pytest

from oracle import geometry as g

pytestmark = pytest.mark.gpu


def draw_beads(shape, positions, diameters=20, value=1000, dtype=np.uint16):       # tests/test_beads.py:9-36
    positions = np.atleast_2d(np.array(positions))
    diameters = np.full(len(positions), diameters) if np.isscalar(diameters) else np.array(diameters)
    values = np.full(len(positions), value) if np.isscalar(value) else np.array(value)
    img = np.zeros(shape, dtype=dtype)
    for pos, d, v in zip(positions, diameters, values):
        pts = g.filled_circle_points(int(d) // 2) + pos
        ok = (pts[:, 0] >= 0) & (pts[:, 0] < shape[0]) & (pts[:, 1] >= 0) & (pts[:, 1] < shape[1])
        img[pts[ok, 0], pts[ok, 1]] = v
    return img


def draw_chip(shape, button_diameter=20, row_dist=100, col_dist=100, value=1000, blanks=()):    # tests/test_chip.py:9-34
    chip = np.zeros(((shape[0] + 1) * row_dist, (shape[1] + 1) * col_dist), dtype=np.uint16)
    disc = g.filled_circle_points(button_diameter // 2)
    for i in range(shape[0]):
        for j in range(shape[1]):
            if (i, j) not in set(blanks):
                chip[disc[:, 0] + (i + 1) * row_dist, disc[:, 1] + (j + 1) * col_dist] = value
    return chip


def bead_assay(image, channels=None):
    from magnify_b200.dataset import Assay

    image = image[None] if image.ndim == 2 else image
    names = channels or [f"c{k}" for k in range(image.shape[0])]
    return Assay({"image": (("channel", "time", "im_y", "im_x"), image[:, None])},
                 coords={"channel": (("channel",), np.array(names))})


def radii_from_fg(out):
    return np.sqrt(out.fg.values[:, 0].sum(axis=(1, 2)) / np.pi)


def test_bead_single(cuda_device):                                                  # tests/test_beads.py:50-66
    from magnify_b200.components import BeadFinder

    out = BeadFinder(min_bead_diameter=16, max_bead_diameter=24, num_iter=100)(bead_assay(draw_beads((1024, 1024), [512, 512])))
    assert out.sizes["mark"] == 1
    assert 0.95 * 10 < radii_from_fg(out)[0] < 1.05 * 10
    assert 0.95 * 512 < out.x.values.item() < 1.05 * 512 and 0.95 * 512 < out.y.values.item() < 1.05 * 512


def test_beads_multiple_edges_sizes_and_float(cuda_device):                         # :69-160, :235-247
    from magnify_b200.components import BeadFinder

    positions = [[200, 200], [200, 800], [512, 512], [800, 200], [800, 800]]
    out = BeadFinder(16, 24, num_iter=10000)(bead_assay(draw_beads((1024, 1024), positions)))
    assert out.sizes["mark"] == 5
    assert np.all(radii_from_fg(out) > 9) and np.all(radii_from_fg(out) < 11)
    found = {(int(round(y / 10)), int(round(x / 10))) for x, y in zip(out.x.values[:, 0], out.y.values[:, 0])}
    assert found == {(p[0] // 10, p[1] // 10) for p in positions}
    # beads near the image boundary (:98-118)
    edge = [[20, 512], [512, 20], [1003, 512], [512, 1003]]
    out = BeadFinder(16, 24, num_iter=10000)(bead_assay(draw_beads((1024, 1024), edge)))
    assert out.sizes["mark"] == 4 and out.roi.shape[-2:] == (48, 48)
    # float32 input (:235-247)
    out = BeadFinder(16, 24, num_iter=2000)(bead_assay(draw_beads((512, 512), [256, 256], dtype=np.float32)))
    assert out.sizes["mark"] == 1 and out.roi.dtype == np.float32
    # empty image -> zero marks (:219-232)
    out = BeadFinder(16, 24, num_iter=1000)(bead_assay(np.zeros((256, 256), np.uint16)))
    assert out.sizes["mark"] == 0


def test_beads_second_channel_adds_only_new_beads(cuda_device):                     # :282-430 (search semantics)
    from magnify_b200.components import BeadFinder

    a = draw_beads((512, 512), [[100, 100], [300, 300]])
    b = draw_beads((512, 512), [[100, 100], [400, 150]])
    out = BeadFinder(16, 24, num_iter=10000)(bead_assay(np.stack([a, b]), ["a", "b"]))
    assert out.sizes["mark"] == 3
    out = BeadFinder(16, 24, num_iter=10000, search_channel="b")(bead_assay(np.stack([a, b]), ["a", "b"]))
    assert out.sizes["mark"] == 2


def chip_assay(image, shape, t=1, blanks=()):
    from magnify_b200.dataset import Assay

    tag = np.full(shape, "default", dtype="<U200")
    for i, j in blanks:
        tag[i, j] = ""
    image = np.broadcast_to(image, (1, t) + image.shape).copy()
    return Assay({"image": (("channel", "time", "im_y", "im_x"), image)},
                 coords={"channel": (("channel",), np.array(["c0"])), "tag": (("mark_row", "mark_col"), tag),
                         "valid": (("mark_row", "mark_col", "time"), np.ones(shape + (t,), bool))})


CHIP = dict(row_dist=100, col_dist=100, min_button_diameter=16, max_button_diameter=32, chamber_diameter=60,
            min_roundness=0.2, cluster_penalty=50)


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_one_by_one_chip(cuda_device, dtype):                                       # tests/test_chip.py:53-96
    from magnify_b200.components import ButtonFinder

    out = ButtonFinder(num_iter=100, **CHIP)(chip_assay(draw_chip((1, 1), 20).astype(dtype), (1, 1)))
    assert out.sizes["mark"] == 1
    assert 0.9 * 10 < np.sqrt(out.fg.values.sum() / np.pi) < 1.1 * 10
    assert 0.95 * 100 < out.x.values.item() < 1.05 * 100


def test_ten_by_ten_chip_with_blanks_and_copy_forward(cuda_device):                 # :99-127, :449-456
    from magnify_b200.components import ButtonFinder

    blanks = ((2, 3), (7, 7))
    out = ButtonFinder(num_iter=10000, **CHIP)(chip_assay(draw_chip((10, 10), 20, blanks=blanks), (10, 10), t=3, blanks=blanks))
    assert out.sizes["mark"] == 100
    x, y = out.x.values.reshape(10, 10, 3), out.y.values.reshape(10, 10, 3)
    filled = np.ones((10, 10), bool)
    for b in blanks:
        filled[b] = False
    radii = np.sqrt(out.fg.values[:, 0].sum(axis=(1, 2)) / np.pi).reshape(10, 10)
    assert 9 < radii[filled].min() and radii[filled].max() < 11
    assert 95 < x[0, 0, 0] < 105 and 95 < y[0, 0, 0] < 105
    assert 395 < x[4, 3, 0] < 405 and 495 < y[4, 3, 0] < 505
    for i in range(10):
        for j in range(10):
            assert abs(x[i, j, 0] - (j + 1) * 100) <= 3 and abs(y[i, j, 0] - (i + 1) * 100) <= 3
    for ti in (1, 2):                                                                # copy-forward
        np.testing.assert_array_equal(x[..., ti], x[..., 0])
        np.testing.assert_array_equal(out.fg.values[:, ti], out.fg.values[:, 0])


def test_full_size_chip_all_buttons_found(cuda_device):
    """Config-2 geometry (7784^2 image, 56 x 32 buttons at the 'pc' spacing, radii 10-15) with the
    reference's `microfluidic_chip` defaults (registry.py:205-235): every button is located exactly
    and its refined radius reproduces the disc area.  Background constant, like the reference's own
    fixtures (on a noisy background the reference's finder itself reports mostly noise circles)."""
    from magnify_b200.components import ButtonFinder

    rows, cols, side = 56, 32, 7784
    rng = np.random.default_rng(0)
    row_dist, col_dist = 406 / 3.22, 750 / 3.22
    cy = np.round((side - (rows - 1) * row_dist) / 2 + np.arange(rows)[:, None] * row_dist + rng.uniform(-2, 2, (rows, cols)))
    cx = np.round((side - (cols - 1) * col_dist) / 2 + np.arange(cols)[None, :] * col_dist + rng.uniform(-2, 2, (rows, cols)))
    radius = 10 + (np.add.outer(np.arange(rows), np.arange(cols)) % 6)
    image = np.full((side, side), 400, dtype=np.uint16)
    yy, xx = np.mgrid[-16:17, -16:17]
    for i in range(rows):
        for j in range(cols):
            y, x = int(cy[i, j]), int(cx[i, j])
            image[y - 16:y + 17, x - 16:x + 17][yy * yy + xx * xx <= radius[i, j] ** 2] = 3000 + 37 * ((i * cols + j) % 50)
    out = ButtonFinder(row_dist=row_dist, col_dist=col_dist, min_button_diameter=8, max_button_diameter=30,
                       chamber_diameter=60, num_iter=5_000_000, min_roundness=0.2, cluster_penalty=50,
                       device=cuda_device)(chip_assay(image, (rows, cols)))
    assert out.sizes["mark"] == rows * cols
    np.testing.assert_array_equal(out.x.values[:, 0], cx.reshape(-1))
    np.testing.assert_array_equal(out.y.values[:, 0], cy.reshape(-1))
    found_radius = np.sqrt(out.fg.values[:, 0].sum(axis=(1, 2)) / np.pi)
    assert np.abs(found_radius - radius.reshape(-1)).max() < 0.5

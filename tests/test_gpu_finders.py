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


# ---- more of tests/test_chip.py ---------------------------------------------------------------
@pytest.mark.parametrize("shape,button,row_dist,col_dist,kw,num_iter", [
    ((3, 5), 20, 100, 100, {}, 5000),                                                     # :135-159
    ((5, 3), 20, 100, 100, {}, 5000),                                                     # :162-186
    ((4, 4), 40, 150, 150, dict(min_button_diameter=30, max_button_diameter=50, chamber_diameter=100), 5000),   # :194-221
    ((4, 4), 20, 80, 120, {}, 5000),                                                      # :224-252
    ((2, 2), 20, 100, 100, {}, 1000),                                                     # :260-283
])
def test_chip_geometries(cuda_device, shape, button, row_dist, col_dist, kw, num_iter):
    from magnify_b200.components import ButtonFinder

    params = dict(CHIP, row_dist=row_dist, col_dist=col_dist)
    params.update(kw)
    out = ButtonFinder(num_iter=num_iter, **params)(chip_assay(draw_chip(shape, button, row_dist, col_dist), shape))
    assert out.sizes["mark"] == shape[0] * shape[1]
    x, y = out.x.values.reshape(shape), out.y.values.reshape(shape)
    for i in range(shape[0]):
        for j in range(shape[1]):
            assert abs(x[i, j] - (j + 1) * col_dist) < 5 and abs(y[i, j] - (i + 1) * row_dist) < 5
    radii = np.sqrt(out.fg.values[:, 0].sum(axis=(1, 2)) / np.pi)
    assert 0.85 * button / 2 < radii.min() and radii.max() < 1.15 * button / 2


def test_chip_blank_positions_with_default_tags(cuda_device):                          # :286-311
    from magnify_b200.components import ButtonFinder

    blanks = [(0, 0), (1, 2), (2, 1), (3, 3)]
    out = ButtonFinder(num_iter=5000, **CHIP)(chip_assay(draw_chip((4, 4), 20, blanks=blanks), (4, 4)))
    assert out.sizes["mark"] == 16
    assert (out.fg.values[:, 0].sum(axis=(1, 2)) > 100).sum() >= 12


def test_chip_refinding_follows_shifted_buttons(cuda_device):                          # :487-546, :549-597
    from magnify_b200.components import ButtonFinder

    t0 = draw_chip((2, 2), 20)
    t1 = np.zeros_like(t0)
    t1[10:, 10:] = t0[:-10, :-10]
    assay = chip_assay(t0, (2, 2), t=2)
    assay["image"] = (("channel", "time", "im_y", "im_x"), np.stack([t0, t1])[None])
    out = ButtonFinder(num_iter=5000, search_timestep=[0, 1], **CHIP)(assay)
    x, y = out.x.values.reshape(2, 2, 2), out.y.values.reshape(2, 2, 2)
    for i in range(2):
        for j in range(2):
            assert abs(x[i, j, 0] - (j + 1) * 100) < 5 and abs(x[i, j, 1] - ((j + 1) * 100 + 10)) < 5
            assert abs(y[i, j, 0] - (i + 1) * 100) < 5 and abs(y[i, j, 1] - ((i + 1) * 100 + 10)) < 5
    # searched only at t = 0: t = 1 copies the t = 0 positions although the buttons moved
    out = ButtonFinder(num_iter=5000, search_timestep=0, **CHIP)(assay.copy())
    np.testing.assert_array_equal(out.x.values[:, 0], out.x.values[:, 1])
    assert np.abs(out.x.values[:, 0].reshape(2, 2)[0] - np.array([100, 200])).max() < 5


def test_chip_search_channel(cuda_device):                                             # :655-697
    from magnify_b200.components import ButtonFinder
    from magnify_b200.dataset import Assay

    chip = draw_chip((3, 3), 20)
    image = np.stack([chip, np.zeros_like(chip)])[:, None]
    tag = np.full((3, 3), "default", dtype="<U200")
    assay = Assay({"image": (("channel", "time", "im_y", "im_x"), image)},
                  coords={"channel": (("channel",), np.array(["bf", "gfp"])), "tag": (("mark_row", "mark_col"), tag),
                          "valid": (("mark_row", "mark_col", "time"), np.ones((3, 3, 1), bool))})
    out = ButtonFinder(num_iter=5000, search_channel="bf", **CHIP)(assay)
    x, y = out.x.values.reshape(3, 3), out.y.values.reshape(3, 3)
    for i in range(3):
        for j in range(3):
            assert abs(x[i, j] - (j + 1) * 100) < 5 and abs(y[i, j] - (i + 1) * 100) < 5
    radii = np.sqrt(out.fg.values[:, 0].sum(axis=(1, 2)) / np.pi)
    assert 8 < radii.min() and radii.max() < 12 and out.roi.shape[1] == 2


# ---- more of tests/test_beads.py ---------------------------------------------------------------
def test_beads_sizes_spacing_intensity(cuda_device):                                   # :121-216
    from magnify_b200.components import BeadFinder

    out = BeadFinder(14, 32, num_iter=10000)(bead_assay(draw_beads((1024, 1024), [[300, 300], [300, 700], [700, 300], [700, 700]],
                                                                   diameters=[16, 20, 24, 28])))
    areas = out.fg.values[:, 0].sum(axis=(1, 2))
    assert out.sizes["mark"] == 4 and areas.max() / areas.min() > 1.5
    out = BeadFinder(16, 24, num_iter=10000)(bead_assay(draw_beads((1024, 1024), [[500, 500], [500, 540], [540, 500]])))
    assert out.sizes["mark"] == 3
    pos = np.stack([out.x.values[:, 0], out.y.values[:, 0]], 1)
    assert min(np.linalg.norm(pos[i] - pos[j]) for i in range(3) for j in range(i)) > 20
    out = BeadFinder(16, 24, num_iter=10000)(bead_assay(draw_beads((1024, 1024), [[300, 500], [500, 500], [700, 500]],
                                                                   value=[500, 1000, 2000])))
    assert out.sizes["mark"] == 3 and (radii_from_fg(out) > 8.5).all()

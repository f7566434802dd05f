"""Generate the golden fixtures in tests/golden/ from the REAL reference helpers.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
It loads the reference's own src/magnify/utils.py in place (oracle/_refload.py; nothing is
copied) and replays the reference's call sites for the hot path on small deterministic
inputs, storing inputs and outputs as .npz.  The call sites replayed are quoted by file:line.
The fixtures pin the oracle (tests/test_oracle_*.py, CPU) and the CUDA path (tests/test_gpu_*.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle._refload import load_reference_utils  # noqa: E402

utils = load_reference_utils()
if utils is None:
    raise SystemExit("reference utils.py not available; run inside the build container")


def pattern_image(c, t, h, w, salt=0):
    """Deterministic uint16 image stack with every pixel distinct in its neighbourhood."""
    cc, tt, yy, xx = np.meshgrid(np.arange(c), np.arange(t), np.arange(h), np.arange(w), indexing="ij")
    v = cc * 7919 + tt * 10473 + yy * 131 + xx * 31 + (yy * xx) % 977 + salt * 2221
    return (v % 65536).astype(np.uint16)


def golden_geometry():
    rng = np.random.default_rng(20261018)
    # bounding_box known answers (utils.py:60-80), including boxes sliding off every edge.
    n = 4000
    w = rng.integers(1, 400, n)
    h = rng.integers(1, 400, n)
    length = rng.integers(1, 130, n)
    x = rng.integers(-60, 460, n)
    y = rng.integers(-60, 460, n)
    out = np.array(
        [utils.bounding_box(int(x[i]), int(y[i]), int(length[i]), int(w[i]), int(h[i])) for i in range(n)],
        dtype=np.int64,
    )
    # filled_circle_points (utils.py:398-430): per-radius row half-widths and areas.
    rmax = 160
    hw = np.full((rmax + 1, rmax + 1), -1, dtype=np.int32)
    area = np.zeros(rmax + 1, dtype=np.int64)
    perim = np.zeros(rmax + 1, dtype=np.int64)
    for r in range(1, rmax + 1):
        pts = utils.filled_circle_points(r)
        area[r] = len(pts)
        perim[r] = len(utils.circle_points(r))
        assert len(set(map(tuple, pts.tolist()))) == len(pts)
        for d in range(0, r + 1):
            cols = pts[pts[:, 0] == d, 1]
            cols_neg = pts[pts[:, 0] == -d, 1]
            assert cols.min() == -cols.max() and np.array_equal(np.sort(cols), np.sort(cols_neg))
            assert len(cols) == 2 * cols.max() + 1  # one contiguous, centred span
            hw[r, d] = cols.max()
    np.savez_compressed(
        os.path.join(HERE, "geometry.npz"),
        bb_args=np.stack([x, y, length, w, h], 1), bb_out=out,
        disc_halfwidth=hw, disc_area=area, perimeter_len=perim,
    )


def golden_beads():
    """Replay BeadFinder's ROI/mask half, find.py:561-602, with the real utils."""
    rng = np.random.default_rng(7)
    c, t, h, w, length = 2, 2, 200, 232, 50
    image = pattern_image(c, t, h, w, salt=1)
    # (row, col, radius): isolated, overlapping pairs, a triple, border-straddling, corner.
    beads = np.array(
        [[60, 60, 10], [60, 75, 10], [120, 40, 12], [125, 52, 8], [118, 55, 6], [3, 100, 9],
         [199, 231, 11], [100, 228, 7], [150, 150, 25], [30, 180, 5], [170, 30, 1], [90, 120, 16]],
        dtype=np.float64,
    )
    labels = utils.circle_labels(beads.astype(int), h, w)  # find.py:561
    x = beads[:, 1]
    y = beads[:, 0]
    m = len(beads)
    fg = np.empty((m, length, length), dtype=bool)
    bg = np.empty_like(fg)
    roi = np.empty((m, c, t, length, length), dtype=image.dtype)
    boxes = np.empty((m, 4), dtype=np.int64)
    for i in range(m):
        top, bottom, left, right = utils.bounding_box(round(x[i]), round(y[i]), length, w, h)  # :573-579
        boxes[i] = (top, bottom, left, right)
        sub = labels[top:bottom, left:right]
        fg[i] = sub == i  # :582
        bg[i] = sub == -1  # :584
    for ch in range(c):
        im = image[ch]  # (T, H, W)  :590
        for j in range(m):
            top, bottom, left, right = boxes[j]
            roi[j, ch] = im[..., top:bottom, left:right]  # :601
    # the reference's OWN BeadFinder.__call__ (find.py:471-605) with the centre finder pinned to
    # `beads` must give exactly the arrays replayed above
    from oracle._refload import reference_bead_finder

    ref = reference_bead_finder(image, ["a", "b"], beads, length)
    assert ref is not None
    assert np.array_equal(ref["roi"], roi)
    assert np.array_equal(ref["fg"], np.repeat(fg[:, None], t, 1)) and np.array_equal(ref["bg"], np.repeat(bg[:, None], t, 1))
    assert np.array_equal(ref["x"][:, 0], beads[:, 1]) and np.array_equal(ref["y"][:, 0], beads[:, 0]) and ref["valid"].all()
    np.savez_compressed(
        os.path.join(HERE, "beads.npz"),
        image_salt=1, image_shape=np.array([c, t, h, w]), roi_length=length, beads=beads,
        labels=labels, fg=fg, bg=bg, roi=roi, boxes=boxes,
    )


def golden_chip():
    """Replay ButtonFinder.find_rois' crop + masks (find.py:362-400) for pinned refinement
    results, and the copy-forward crops (find.py:143-176), with the real utils."""
    c, t, h, w = 2, 3, 300, 420
    length, chamber_r, max_r = 72, 30, 15
    image = pattern_image(c, t, h, w, salt=2)
    rows, cols = 2, 3
    # coarse (non-integer) centres incl. exact .5 ties and boxes clipped at the borders
    gx = np.array([[20.5, 190.49, 400.5], [35.5, 200.5, 419.0]])
    gy = np.array([[10.5, 40.2, 36.5], [250.5, 290.0, 299.49]])
    # pinned outcome of the CPU refinement at search timestep 0: (y, x, r) in ROI coords or none
    found = {(0, 1): (30, 41, 9), (1, 0): (36, 36, 12), (1, 2): (50, 60, 14)}
    x = gx.copy()
    y = gy.copy()
    roi = np.empty((rows, cols, c, t, length, length), dtype=image.dtype)
    fg = np.empty((rows, cols, t, length, length), dtype=bool)
    bg = np.empty_like(fg)
    rad = np.full((rows, cols), max_r, dtype=np.int32)
    images0 = image[:, 0]
    for i in range(rows):
        for j in range(cols):
            top, bottom, left, right = utils.bounding_box(round(x[i, j]), round(y[i, j]), length, w, h)  # :327-333
            roi[i, j, :, 0] = images0[..., top:bottom, left:right]
            if (i, j) in found:
                by, bx, br = found[(i, j)]
                y[i, j], x[i, j] = by, bx  # :365
                x[i, j] += left  # :367
                y[i, j] += top  # :368
                top, bottom, left, right = utils.bounding_box(round(x[i, j]), round(y[i, j]), length, w, h)  # :370-376
                roi[i, j, :, 0] = images0[..., top:bottom, left:right]  # :377
                rad[i, j] = br  # :378
            x_rel = round(x[i, j]) - left  # :380
            y_rel = round(y[i, j]) - top  # :381
            bg[i, j, 0] = utils.annulus((length, length), (y_rel, x_rel), outer_radius=chamber_r,
                                        inner_radius=max_r, value=1)  # :384-390
            fg[i, j, 0] = utils.circle((length, length), (y_rel, x_rel), radius=int(rad[i, j]), value=1)  # :392-397
    for tt in range(1, t):  # find.py:143-176 with search_timesteps=[0] -> copy_t = t-1
        for i in range(rows):
            for j in range(cols):
                top, bottom, left, right = utils.bounding_box(round(x[i, j]), round(y[i, j]), length, w, h)
                roi[i, j, :, tt] = image[:, tt, top:bottom, left:right]
        fg[:, :, tt] = fg[:, :, tt - 1]
        bg[:, :, tt] = bg[:, :, tt - 1]
    # the reference's OWN ButtonFinder.__call__ (find.py:55-203, find_rois :308-402, copy-forward
    # :143-181, stack :182) with find_centers / find_circles pinned must give the same arrays
    from oracle._refload import reference_button_finder

    refine = [found.get((i, j)) for i in range(rows) for j in range(cols)]
    ref = reference_button_finder(image, ["a", "b"], np.full((rows, cols), "default"), gx, gy, refine)
    assert ref is not None
    assert np.array_equal(ref["roi"], roi.reshape((rows * cols,) + roi.shape[2:]))
    assert np.array_equal(ref["fg"], fg.reshape((rows * cols,) + fg.shape[2:]))
    assert np.array_equal(ref["bg"], bg.reshape((rows * cols,) + bg.shape[2:]))
    assert np.array_equal(ref["x"][:, 0], x.reshape(-1)) and np.array_equal(ref["y"][:, 0], y.reshape(-1))
    np.savez_compressed(
        os.path.join(HERE, "chip.npz"),
        image_salt=2, image_shape=np.array([c, t, h, w]), roi_length=length, chamber_radius=chamber_r,
        max_button_radius=max_r, coarse_x=gx, coarse_y=gy, x=x, y=y, fg_radius=rad,
        refine=np.array([(-1, -1, -1) if r is None else r for r in refine]),
        roi=roi.reshape((rows * cols,) + roi.shape[2:]),
        fg=fg.reshape((rows * cols,) + fg.shape[2:]), bg=bg.reshape((rows * cols,) + bg.shape[2:]),
    )


def golden_chip_multi():
    """Two search timesteps with different refinement outcomes + copy-forward in between
    (find.py:119-181): produced directly by the reference's own ButtonFinder.__call__."""
    from oracle._refload import reference_button_finder

    c, t, h, w, length = 2, 5, 220, 260, 40
    image = pattern_image(c, t, h, w, salt=3)
    gx = np.array([[30.5, 128.2, 240.0], [41.0, 130.5, 255.5]])
    gy = np.array([[25.0, 30.5, 18.5], [190.5, 200.0, 215.49]])
    rows, cols = gx.shape
    # refinement results in call order: search timestep 1 (row-major), then search timestep 3
    refine = [(20, 22, 7), None, (18, 25, 9), None, (21, 19, 6), (30, 33, 8),
              None, (19, 20, 5), (20, 20, 9), (22, 18, 8), None, (10, 12, 6)]
    ref = reference_button_finder(image, ["a", "b"], np.full((rows, cols), "default"), gx, gy, refine, roi_length=length,
                                  min_button_diameter=8, max_button_diameter=18, chamber_diameter=34,
                                  search_timestep=[1, 3])
    assert ref is not None
    np.savez_compressed(os.path.join(HERE, "chip_multi.npz"), image_salt=3, image_shape=np.array([c, t, h, w]),
                        roi_length=length, chamber_radius=17, max_button_radius=9, search_timesteps=np.array([1, 3]),
                        refine=np.array([(-1, -1, -1) if r is None else r for r in refine]), coarse_x=gx, coarse_y=gy,
                        **ref)


def golden_filter_expression():
    """`filter_expression` (filter.py:11-37) executed in place on u16 ROIs: the composition is the
    reference's; the xarray where/median arithmetic is restated in oracle/_refload.py and the
    outcome is asserted to be independent of the float32-vs-float64 promotion of `where`."""
    from oracle._refload import reference_filter_expression

    rng = np.random.default_rng(5)
    m, c, t, length = 48, 3, 2, 20
    roi = rng.integers(380, 420, (m, c, t, length, length)).astype(np.uint16)
    yy, xx = np.mgrid[0:length, 0:length]
    fg0 = (yy - 10) ** 2 + (xx - 10) ** 2 <= 25
    fg = np.broadcast_to(fg0, (m, t, length, length)).copy()
    bg = np.broadcast_to((yy - 10) ** 2 + (xx - 10) ** 2 > 64, (m, t, length, length)).copy()
    for i in range(0, m, 3):
        roi[i][..., fg0] += np.uint16(50 + i)
    roi[4, 1][..., fg0] += 3000
    roi[13, 2][..., fg0] += 2000
    fg[7] = False          # empty foreground -> NaN median -> comparison False
    valid = rng.random((m, t)) < 0.9
    out = {"roi": roi, "fg": fg, "bg": bg, "valid": valid, "channels": np.array(["a", "b", "c"])}
    cases = [(None, None), (None, 40), ("b", None), ("b", 40), (["a", "c"], None), (["a", "c"], 25)]
    for k, (sc, mc) in enumerate(cases):
        got = reference_filter_expression(roi, fg, bg, valid, ["a", "b", "c"], sc, mc)
        assert got is not None
        assert np.array_equal(got, reference_filter_expression(roi, fg, bg, valid, ["a", "b", "c"], sc, mc, np.float64))
        out[f"case{k}__valid"] = got
        out[f"case{k}__search"] = np.array([] if sc is None else np.atleast_1d(sc))
        out[f"case{k}__min_contrast"] = np.array(-1 if mc is None else mc)
    # identify.py:76-80 (mean of fg - median of bg at time 0) by the reference's own expression,
    # under both promotions xarray's `where` may apply to u16 (float32 by maybe_promote; float64)
    from oracle._refload import reference_mrbles_intensities

    out["intensity_channels"] = np.array(["c", "a"])
    out["intensities_f32"] = reference_mrbles_intensities(roi, fg, bg, ["a", "b", "c"], ["c", "a"])
    out["intensities_f64"] = reference_mrbles_intensities(roi, fg, bg, ["a", "b", "c"], ["c", "a"], np.float64)
    assert out["intensities_f32"].dtype == np.float32 and out["intensities_f64"].dtype == np.float64
    # filter_leaky (filter.py:65-94) on the same ROIs laid out as a 12 x 4 chip, every fourth
    # marker blank (some of them bright, one only in channel c); the reference's own function, both promotions
    from oracle._refload import reference_filter_leaky

    mark_row = np.repeat(np.arange(12), 4)
    tag = np.array(["" if i % 4 == 1 else f"m{i}" for i in range(m)])
    out["tag"], out["mark_row"] = tag, mark_row
    for k, sc in enumerate([None, "b", ["c", "a"]]):
        got = reference_filter_leaky(roi, fg, bg, valid, ["a", "b", "c"], tag, mark_row, sc)
        assert np.array_equal(got, reference_filter_leaky(roi, fg, bg, valid, ["a", "b", "c"], tag, mark_row, sc, np.float64))
        out[f"leaky{k}__valid"] = got
        out[f"leaky{k}__search"] = np.array([] if sc is None else np.atleast_1d(sc))
    np.savez_compressed(os.path.join(HERE, "filter.npz"), **out)


def golden_circles():
    """The reference's own utils.find_circles (utils.py:100-218; unseeded RANSAC) on a fixture with
    six planted discs and clean edges: with 200000 draws the best-scoring rounded circle of every
    disc is found on every run, so circles and scores are reproducible."""
    from oracle import circles as oc
    from oracle._refload import reference_find_circles_stages

    rng = np.random.default_rng(12)
    h, w = 260, 300
    img = (rng.random((h, w)) * 3).astype(np.float64) * 20
    yy, xx = np.mgrid[0:h, 0:w]
    discs = [(50, 60, 18), (60, 200, 12), (150, 130, 25), (200, 40, 9), (210, 240, 15), (120, 260, 10)]
    for cy, cx, r in discs:
        img[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] += 3000
    image = oc.to_uint8(img.astype(np.uint16))
    kw = dict(low_edge_quantile=0.85, high_edge_quantile=0.97, grid_length=20, num_iter=200000, min_radius=6,
              max_radius=30, min_roundness=0.3, min_dist=6)
    edges, circles, scores = reference_find_circles_stages(image, **kw)
    found = {tuple(d) for d in discs if any(abs(c[0] - d[0]) <= 2 and abs(c[1] - d[1]) <= 2 and abs(c[2] - d[2]) <= 2
                                           for c in circles)}
    assert len(found) == len(discs) == len(circles), (found, circles)
    again = reference_find_circles_stages(image, **kw)      # the optimum is found every time on this fixture
    assert {tuple(c) for c in again[1]} == {tuple(c) for c in circles}
    np.savez_compressed(os.path.join(HERE, "circles.npz"), image=image, discs=np.array(discs), edges=edges,
                        circles=circles, scores=scores, **{k: np.array(v) for k, v in kw.items()})


def golden_nonround():
    """`filter_nonround` (filter.py:40-62) executed in place with the real cv2 on elliptical, empty,
    single-pixel and two-component foregrounds."""
    from oracle._refload import reference_filter_nonround

    rng = np.random.default_rng(0)
    m, t, length = 40, 2, 48
    yy, xx = np.mgrid[0:length, 0:length]
    fg = np.zeros((m, t, length, length), bool)
    for i in range(m):
        a, b = rng.uniform(4, 18), rng.uniform(4, 18)
        if i % 5 == 0:
            b = a
        fg[i, :] = ((yy - 24) / a) ** 2 + ((xx - 24) / b) ** 2 <= 1
    fg[3] = False                                          # no contour at all
    fg[7] = False
    fg[7, :, 10, 10] = True                                # one pixel: a contour of length 0
    fg[9, :] |= (np.abs(yy - 5) <= 1) & (xx > 3) & (xx < 40)   # a second component
    fg[11, :] &= ~(((yy - 24) ** 2 + (xx - 24) ** 2) <= 9)     # a hole (ignored by RETR_EXTERNAL)
    valid = rng.random((m, t)) < 0.9
    out = {"fg": fg, "valid": valid}
    for k, mr in enumerate((0.75, 0.5, 0.9)):
        got = reference_filter_nonround(fg, valid, ["a", "b"], mr)
        assert got is not None
        out[f"case{k}__valid"], out[f"case{k}__min_roundness"] = got, np.array(mr)
    np.savez_compressed(os.path.join(HERE, "nonround.npz"), **out)


def golden_bead_field():
    """The reference's own find_circles (default 5e6 unseeded draws) on the 2048^2 / 300-bead field of
    tools/finder_bench.py: a realistic-size run to compare detection sets with (not bit-exact: random)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "tools"))
    import finder_bench as fb
    from oracle import circles as oc
    from oracle._refload import load_reference_utils

    utils = load_reference_utils()
    img = oc.to_uint8(fb.bead_image())
    circles, scores = utils.find_circles(img, **fb.BEADS, gui=None)
    assert len(circles) > 250
    np.savez_compressed(os.path.join(HERE, "bead_field_reference.npz"), circles=circles, scores=scores,
                        **{k: np.array(v) for k, v in fb.BEADS.items()})


def golden_masks_cv():
    """cv.circle rasters through utils.circle / utils.annulus (utils.py:30-52) for clipped and
    unclipped centres -- pins the closed form `dx^2+dy^2 <= r^2` used by oracle and kernel."""
    cases = []
    discs = []
    rings = []
    for (cy, cx, r, ro, ri) in [(36, 36, 10, 30, 15), (0, 0, 15, 30, 15), (71, 71, 7, 30, 15), (5, 66, 12, 30, 16),
                                (36, 2, 15, 33, 15), (40, 40, 0, 20, 20), (-3, 30, 9, 30, 15), (36, 75, 14, 30, 15)]:
        cases.append((cy, cx, r, ro, ri))
        discs.append(utils.circle((72, 72), (cy, cx), r).astype(bool))
        rings.append(utils.annulus((72, 72), (cy, cx), ro, ri, value=1).astype(bool))
    np.savez_compressed(os.path.join(HERE, "masks_cv.npz"), cases=np.array(cases), disc=np.array(discs),
                        ring=np.array(rings))


def golden_flatfield():
    """Run the reference's OWN flatfield_correct source (preprocess.py:62-88, loaded in place with
    stub imports, ndarray operands) on small deterministic stacks."""
    from oracle._refload import reference_flatfield_correct

    rng = np.random.default_rng(99)
    c, t, r, cc, h, w = 2, 2, 1, 2, 32, 40
    tiles = np.clip(rng.normal(2500, 1500, (c, t, r, cc, h, w)), 0, 65535).astype(np.uint16)
    tiles[0, 0, 0, 0, 3, 4] = 65535
    tiles[..., :2, :] = 0
    yy, xx = np.mgrid[0:h, 0:w]
    flat = 1.0 + 0.35 * np.cos(yy / h * 2.0) * np.sin(xx / w * 1.5 + 0.2)
    dark = 100.0 + 5.0 * np.sin(yy * 0.37 + xx * 0.11)
    flat_c = np.stack([flat, flat * 1.1])[:, None, None, None]
    dark_c = np.stack([dark, dark + 3.0])[:, None, None, None]
    cases = {
        "arrays": (flat, dark), "scalars": (0.5, 97.25), "defaults": (1.0, 0.0), "integer_dark": (1.0, 100.0),
        "per_channel": (flat_c, dark_c), "scalar_flat_array_dark": (1.25, dark),
    }
    out = {"tiles": tiles}
    for name, (f, d) in cases.items():
        res = reference_flatfield_correct(tiles, f, d)
        assert res is not None and res.dtype == tiles.dtype and res.shape == tiles.shape
        out[f"{name}__flat"] = np.asarray(f, dtype=np.float64)
        out[f"{name}__dark"] = np.asarray(d, dtype=np.float64)
        out[f"{name}__out"] = res
    tiles32 = (rng.random((1, 1, 1, 2, 16, 24)) * 3000).astype(np.float32)
    out["f32_tiles"] = tiles32
    out["f32__out"] = reference_flatfield_correct(tiles32, flat[:16, :24], dark[:16, :24])
    np.savez_compressed(os.path.join(HERE, "flatfield.npz"), **out)


def golden_stitch():
    """Run the reference's OWN Stitcher source (stitch.py:12-46, loaded in place, named-array
    stand-in for xarray) on small stacks: odd / even / zero overlap, several channels and times."""
    from oracle._refload import reference_stitch

    rng = np.random.default_rng(123)
    out = {}
    for name, shape, ov, dtype in [("odd", (2, 2, 2, 3, 20, 24), 5, np.uint16), ("even", (1, 3, 3, 2, 16, 16), 6, np.float64),
                                   ("zero", (1, 1, 2, 2, 8, 12), 0, np.uint16), ("single", (1, 1, 1, 1, 30, 30), 5, np.float32),
                                   ("max", (1, 1, 2, 2, 10, 10), 9, np.uint8)]:
        tiles = (rng.random(shape) * 200).astype(dtype)
        res = reference_stitch(tiles, ov)
        assert res is not None
        out[name + "__tiles"], out[name + "__overlap"], out[name + "__image"] = tiles, ov, res
    for bad in (-5,):
        try:
            reference_stitch(np.zeros((1, 1, 2, 2, 50, 50)), bad)
            raise AssertionError("reference accepted a negative overlap")
        except ValueError:
            pass
    try:
        reference_stitch(np.zeros((1, 1, 2, 2, 50, 50)), 100)
        raise AssertionError("reference accepted overlap >= tile size")
    except ValueError:
        pass
    np.savez_compressed(os.path.join(HERE, "stitch.npz"), **out)


if __name__ == "__main__":
    golden_geometry()
    golden_flatfield()
    golden_stitch()
    golden_beads()
    golden_chip()
    golden_chip_multi()
    golden_filter_expression()
    golden_circles()
    golden_nonround()
    golden_bead_field()
    golden_masks_cv()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))

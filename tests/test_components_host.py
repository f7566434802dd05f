"""CPU-side behaviour of the component layer: the Assay stand-in, argument errors that must be
raised before anything touches a GPU, and the registry installer's failure mode."""
import numpy as np
import pytest


def test_assay_container():
    from magnify_b200.dataset import Assay

    a = Assay({"tile": (("channel", "y", "x"), np.zeros((2, 4, 5)))}, coords={"channel": (("channel",), np.array(["a", "b"]))})
    assert "tile" in a and "image" not in a and a.sizes == {"channel": 2, "y": 4, "x": 5}
    assert a.tile.dims == ("channel", "y", "x") and a["tile"].shape == (2, 4, 5)
    with pytest.raises(AttributeError):
        a.image
    with pytest.raises(ValueError):
        a["bad"] = (("channel",), np.zeros(3))          # conflicting dimension size
    b = a.assign_coords(valid=(("channel",), np.ones(2, bool)))
    assert "valid" in b and "valid" not in a
    assert b.drop_vars(["tile"]).sizes == {"channel": 2}
    assert a.tile.isel(channel=0).dims == ("y", "x")


def test_constructor_errors_match_reference():
    from magnify_b200.components import BeadFinder, ButtonFinder, Stitcher

    with pytest.raises(ValueError):
        Stitcher(overlap=-1)                                   # stitch.py:8-9
    with pytest.raises(ValueError):
        BeadFinder(min_bead_diameter=20, max_bead_diameter=10)  # find.py:458-459
    with pytest.raises(ValueError):
        ButtonFinder(100, 100, 30, 10, 60)                      # find.py:34-35
    f = ButtonFinder(126.1, 232.9, 16, 30, 60)
    assert (f.min_button_radius, f.max_button_radius, f.chamber_radius, f.roi_length) == (8, 15, 30, 72)  # find.py:37-49
    b = BeadFinder(16, 24)
    assert (b.min_bead_radius, b.max_bead_radius, b.roi_length) == (8, 12, 48)  # find.py:461-467


def test_install_needs_magnify():
    from magnify_b200 import components

    try:
        import magnify  # noqa: F401
    except Exception:
        with pytest.raises(ImportError):
            components.install()
    else:
        names = components.install()
        assert "stitch" in names and "quantify" in names


def test_api_standardize_and_restore_shapes():
    """api._standardized / _restore: the dims bookkeeping of standardize_format / restore_format
    (preprocess.py:11-42, postprocess.py:20-49) without touching the GPU."""
    from magnify_b200 import api
    from magnify_b200.dataset import Var

    arr = np.arange(2 * 3 * 4 * 5, dtype=np.uint16).reshape(3, 2, 4, 5)            # (time, channel, y, x)
    (xp,) = api._standardized(arr, ("time", "channel", "y", "x"), {"channel": ["a", "b"]})
    assert xp["tile"].dims == ("channel", "time", "tile_row", "tile_col", "tile_y", "tile_x")
    assert xp["tile"].values.shape == (2, 3, 1, 1, 4, 5)
    np.testing.assert_array_equal(xp["tile"].values[1, 2, 0, 0], arr[2, 1])
    assert xp.attrs["__original_tile_dims__"] == ["time", "channel", "tile_y", "tile_x"]
    with pytest.raises(ValueError):
        api._standardized(arr, None, None)

    class Labelled:                                   # the duck type of xarray.DataArray
        dims = ("channel", "y", "x")
        values = arr[0]
        coords = {"channel": Var(("channel",), np.array(["bf", "gfp"]))}

    (lab,) = api._standardized(Labelled(), None, None)
    assert lab["tile"].values.shape == (2, 1, 1, 1, 4, 5) and list(lab.coords["channel"].values) == ["bf", "gfp"]
    with pytest.raises(NotImplementedError):
        api._standardized(arr, ("time", "depth", "y", "x"), None)
    # restore: un-stack marks, squeeze the added channel axis, keep the original time axis
    (xp,) = api._standardized(arr[:, 0], ("time", "y", "x"), None)
    xp.data_vars["roi"] = Var(("mark", "channel", "time", "roi_y", "roi_x"), np.zeros((6, 1, 3, 8, 8), np.uint16))
    xp.coords["x"] = Var(("mark", "time"), np.zeros((6, 3)))
    xp.coords["mark_row"] = Var(("mark",), np.repeat(np.arange(2), 3))
    out = api._restore(xp, (2, 3))
    assert out.roi.dims == ("mark_row", "mark_col", "time", "roi_y", "roi_x") and out.roi.shape == (2, 3, 3, 8, 8)
    assert out.x.dims == ("mark_row", "mark_col", "time") and "tile" not in out and "mark_row" not in out
    assert "__original_tile_dims__" not in out.attrs
    only = api._restore(xp, (2, 3), roi_only=True)
    assert only.dims == out.roi.dims and only.shape == out.roi.shape
    kept = api._restore(xp, (2, 3), drop_tiles=False)
    assert kept.tile.dims == ("time", "tile_y", "tile_x") and kept.tile.shape == (3, 4, 5)


def test_pinlist_matches_reference_identify_buttons(tmp_path):
    """api.read_pinlist == the tag array of the reference's identify_buttons (identify.py:13-45, pandas)
    run in place, for names, listed blanks, missing names and a custom blank list."""
    import os

    from magnify_b200 import api
    from oracle._refload import reference_identify_buttons

    path = os.path.join(tmp_path, "pins.csv")
    with open(path, "w") as f:
        f.write("Indices,MutantID,Other\n")
        names = {(1, 1): "wt", (2, 1): "blank", (3, 1): "mutA", (1, 2): "", (2, 2): "BLANK", (3, 2): "a_longer_name_17",
                 (1, 3): "x", (2, 3): "EMPTY", (3, 3): "y"}
        for (col, row), name in names.items():
            f.write(f'"({col},{row})",{name},7\n')
    for blank in (None, ["EMPTY", "x"]):
        mine = api.read_pinlist(path, blank)
        assert mine.shape == (3, 3)
        ref = reference_identify_buttons(4, pinlist=path, blank=blank)
        if ref is None:
            pytest.skip("/root/reference not available (GPU box)")
        np.testing.assert_array_equal(mine, ref[0].astype(mine.dtype))
        assert ref[1].shape == (3, 3, 4) and ref[1].all()
    tag, valid = reference_identify_buttons(2, shape=(2, 5))
    assert tag.shape == (2, 5) and (tag == "default").all() and tag.dtype == np.dtype("<U200")


def test_factories_cover_the_reference_names():
    """The names the predefined pipelines hard-wire (registry.py:243-269, 431-449, 593-610) plus the
    three filters, each a factory(**kwargs) -> callable(assay) like the reference's (registry.py:16-29)."""
    from magnify_b200 import components as comp

    assert set(comp.FACTORIES) == {"flatfield_correct", "stitch", "find_beads", "find_buttons", "filter_expression",
                                   "filter_nonround", "filter_leaky"}
    assert set(comp.EXTRA_FACTORIES) == {"flatfield_stitch_b200", "quantify"}
    assert callable(comp.FACTORIES["stitch"](overlap=10))
    assert callable(comp.FACTORIES["filter_nonround"](min_roundness=0.5, search_channel="egfp"))
    assert callable(comp.FACTORIES["filter_expression"](min_contrast=40))
    with pytest.raises(ValueError):
        comp.FACTORIES["stitch"](overlap=-1)                                         # stitch.py:8-9

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

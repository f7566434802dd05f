"""CPU-side behaviour of the component layer: the Dataset container used when xarray is absent,
argument errors that must be raised before anything touches a GPU, the registry installer."""
import numpy as np
import pytest


def test_dataset_container():
    from magnify_b200.dataset import Dataset

    a = Dataset({"tile": (("channel", "y", "x"), np.zeros((2, 4, 5)))}, coords={"channel": (("channel",), np.array(["a", "b"]))})
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
    assert a.tile.sel(channel="b").dims == ("y", "x") and list(a.channel.values) == ["a", "b"]


def test_dataset_stack_unstack_like_xarray():
    """stack puts the merged dimension last and broadcasts variables that hold only some of the
    dims; unstack puts the level dims last (xarray's orders, which find.py:182 and
    postprocess.py:22 rely on); write-through views (find.py:127-137 assigns into `assay.x[..., t]`)."""
    from magnify_b200.dataset import Dataset

    ds = Dataset({"roi": (("r", "c", "t"), np.arange(24).reshape(2, 3, 4))},
                 coords={"tag": (("r", "c"), np.array([["a", "b", "c"], ["d", "e", "f"]])), "rowinfo": (("r",), [10, 20])})
    st = ds.stack(mark=("r", "c"))
    assert st.roi.dims == ("t", "mark") and st.tag.dims == ("mark",) and st.rowinfo.dims == ("mark",)
    assert list(st.rowinfo.values) == [10, 10, 10, 20, 20, 20] and list(st.r.values) == [0, 0, 0, 1, 1, 1]
    np.testing.assert_array_equal(st.transpose("mark", ...).roi.values, np.arange(24).reshape(6, 4))
    un = st.unstack()
    assert un.roi.dims == ("t", "r", "c")
    np.testing.assert_array_equal(un.roi.transpose("r", "c", "t").values, ds.roi.values)
    ds.roi[:, 1, 2] = -1
    assert (ds["roi"].values[:, 1, 2] == -1).all()
    ds.roi[..., 0] = ds.roi[..., 3]
    np.testing.assert_array_equal(ds.roi.values[..., 0], ds.roi.values[..., 3])
    masked = ds.roi.where(ds.roi > 5)
    assert masked.dtype == np.float64 and np.isnan(masked.values[0, 0, 1])
    assert ds.roi.astype(np.uint16).where(ds.roi > 5).dtype == np.float32          # xarray's maybe_promote
    assert float(masked.median(dim=["r", "c"]).values[1]) == np.nanmedian(masked.values[:, :, 1])


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

    import sys

    if "magnify" not in sys.modules:
        try:
            import magnify  # noqa: F401
        except Exception:
            with pytest.raises(ImportError):
                components.install()
            from magnify_b200 import api

            with pytest.raises(ImportError):
                api.beads(None)

    class Registry(dict):                       # anything with catalogue's `register(name)` works
        def register(self, name):
            return lambda f: self.__setitem__(name, f) or f

    reg = Registry()
    names = components.install(registry=reg)
    assert "stitch" in names and "quantify" in names and reg["stitch"] is components.make_stitch
    assert reg["find_buttons_b200"] is reg["find_buttons"]
    assert "stitch" not in Registry() and components.install(False, Registry()) == names


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


def test_device_array_views_and_pool_on_cpu_tensors():
    """DeviceArray is a duck array: a Dataset keeps it un-materialised through stack / transpose /
    unstack / squeeze / basic indexing (lazy views that replay on the root's single host copy),
    boolean masks expand lazily from their distinct timesteps, nothing holds a reference cycle
    (pinned buffers go back to the pool by refcount), and a host view outlives its DeviceArray."""
    import gc
    import weakref

    import torch

    from magnify_b200.dataset import Dataset
    from magnify_b200.devarray import PINNED, DeviceArray

    t = torch.arange(2 * 3 * 4 * 5, dtype=torch.int32).reshape(2, 3, 4, 5)
    root = DeviceArray(t, extras={"stats": "by-product"})
    ds = Dataset({"roi": (("mark_row", "mark_col", "c", "y"), root)})
    st = ds.stack(mark=("mark_row", "mark_col"), create_index=True).transpose("mark", ...)
    lazy = st.roi.data
    assert isinstance(lazy, DeviceArray) and not lazy.materialised and lazy.extras["stats"] == "by-product"
    assert lazy.tensor.data_ptr() == t.data_ptr() and lazy.tensor.is_contiguous()       # stack + transpose = no copy
    np.testing.assert_array_equal(st.roi.values, t.numpy().reshape(6, 4, 5))
    assert root.materialised                                                            # one host copy, at the root
    un = st.unstack()
    assert un.roi.data.root is root
    np.testing.assert_array_equal(un.roi.transpose("mark_row", "mark_col", ...).values, t.numpy())
    sq = Dataset({"im": (("c", "t", "y", "x"), DeviceArray(t[:1, :1]))}).squeeze("c").squeeze("t")
    assert isinstance(sq.im.data, DeviceArray) and sq.im.shape == (4, 5)
    np.testing.assert_array_equal(sq.im.values, t.numpy()[0, 0])
    masks = DeviceArray(torch.tensor([[[1, 0], [0, 1]], [[1, 1], [0, 0]]], dtype=torch.uint8)[:, None], as_bool=True)
    full = masks.take([0, 0, 0], 1)
    assert full.shape == (2, 3, 2, 2) and full.dtype == np.bool_ and full.numpy().strides[1] == 0   # broadcast, not copied
    assert full.tensor.shape == (2, 3, 2, 2)                                            # and a real tensor for a GPU consumer
    np.testing.assert_array_equal(np.asarray(full)[:, 2], masks.numpy()[:, 0])
    gc.disable()
    try:
        a = DeviceArray(torch.zeros(3, 4))
        w = weakref.ref(a)
        b = a.transpose(1, 0)
        del a
        assert w() is not None
        del b
        assert w() is None                                                              # freed by refcount alone
    finally:
        gc.enable()
    real_empty = torch.empty
    try:                                                                                # pin_memory needs a CUDA runtime
        torch.empty = lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items() if kk != "pin_memory"})
        PINNED.clear()
        block = PINNED.acquire((3, 4), torch.uint16)
        block.tensor.fill_(7)
        view = np.asarray(block)[1:]
        ptr = block.tensor.data_ptr()
        del block
        assert PINNED.cached == 0 and (view == 7).all()                                 # the view keeps the buffer out of the pool
        del view
        assert PINNED.cached == 4096
        again = PINNED.acquire((3, 4), torch.uint16)
        assert again.tensor.data_ptr() == ptr and PINNED.cached == 0
        del again
    finally:
        torch.empty = real_empty
        PINNED.clear()

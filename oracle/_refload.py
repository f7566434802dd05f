"""Loader for the *real* reference helpers (test infrastructure, container-only).

`/root/reference/src/magnify/utils.py` fails to import only because of three GUI
imports (utils.py:9-10,15: napari, napari.types, magnify.plot.vis).  We inject empty
stand-ins for those names and load the file in place -- nothing is copied into this
repo.  `/root/reference` does not exist on the GPU box, so everything that calls
`load_reference_utils()` must tolerate `None` (tests skip; goldens were generated
here by tests/golden/make_golden.py and are committed).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MAGNIFY_REFERENCE_ROOT", "/root/reference")
_cached = None


def load_reference_utils():
    """Return the reference's `magnify.utils` module, or None when it is unavailable."""
    global _cached
    if _cached is not None:
        return _cached
    path = os.path.join(REFERENCE_ROOT, "src", "magnify", "utils.py")
    if not os.path.exists(path):
        return None
    try:
        import cv2  # noqa: F401
        import numba  # noqa: F401
    except Exception:
        return None
    stubs = {}
    if "napari" not in sys.modules:
        napari = types.ModuleType("napari")
        napari_types = types.ModuleType("napari.types")
        napari_types.LayerDataTuple = tuple
        napari.types = napari_types
        stubs["napari"] = napari
        stubs["napari.types"] = napari_types
    if "magnify.plot.vis" not in sys.modules:
        pkg = types.ModuleType("magnify")
        pkg.__path__ = []
        plot = types.ModuleType("magnify.plot")
        plot.__path__ = []
        vis = types.ModuleType("magnify.plot.vis")

        class InteractiveUI:  # placeholder type for annotations only
            pass

        vis.InteractiveUI = InteractiveUI
        stubs.setdefault("magnify", pkg)
        stubs["magnify.plot"] = plot
        stubs["magnify.plot.vis"] = vis
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_magnify_reference_utils", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached = mod
    return mod


_cached_pre = None


class NdarrayAssay:
    """Stand-in for the xarray.Dataset argument of the reference's `flatfield_correct`: `.tile` is a
    plain ndarray (which has the same astype/clip/max/arithmetic surface the function uses) and
    item assignment stores the result."""

    def __init__(self, tile):
        self.tile = tile

    def __setitem__(self, key, value):
        setattr(self, key, value)


def load_reference_preprocess():
    """The reference's `src/magnify/preprocess.py`, loaded in place with stub modules for its
    imports (dask.array, tifffile, xarray, magnify.registry, magnify.utils), or None.  Gives access
    to the reference's OWN `flatfield_correct` source (preprocess.py:62-88), which only needs
    ndarray-like operands for scalar / ndarray flat and dark fields."""
    global _cached_pre
    if _cached_pre is not None:
        return _cached_pre
    path = os.path.join(REFERENCE_ROOT, "src", "magnify", "preprocess.py")
    if not os.path.exists(path):
        return None
    xr = types.ModuleType("xarray")
    xr.Dataset = type("Dataset", (), {})
    xr.DataArray = type("DataArray", (), {})
    dask = types.ModuleType("dask")
    dask_array = types.ModuleType("dask.array")
    dask.array = dask_array
    tifffile = types.ModuleType("tifffile")
    pkg = types.ModuleType("magnify")
    pkg.__path__ = []
    registry = types.ModuleType("magnify.registry")
    registry.component = lambda name: (lambda func: func)   # registry.py:16-29 returns func itself
    utils = types.ModuleType("magnify.utils")
    pkg.registry, pkg.utils = registry, utils
    stubs = {"xarray": xr, "dask": dask, "dask.array": dask_array, "tifffile": tifffile, "magnify": pkg,
             "magnify.registry": registry, "magnify.utils": utils}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_magnify_reference_preprocess", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception:
        mod = None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached_pre = mod
    return mod


def reference_flatfield_correct(tiles, flatfield=1.0, darkfield=0.0):
    """Run the reference's own flatfield_correct on an ndarray tile stack; None if unavailable."""
    mod = load_reference_preprocess()
    if mod is None:
        return None
    xp = NdarrayAssay(tiles)
    mod.flatfield_correct(xp, flatfield=flatfield, darkfield=darkfield)
    return xp.tile


# ---------------------------------------------------------------------------------------------
# The reference's own Stitcher (src/magnify/stitch.py) executed in place on a minimal named-array
# stand-in for xarray.DataArray: only the calls Stitcher.__call__ makes are implemented
# (indexing with Ellipsis + slices, transpose by name, iteration over the first dimension,
# rename, chunk, and xr.concat along an existing dimension).
# ---------------------------------------------------------------------------------------------
class NamedArray:
    def __init__(self, values, dims):
        import numpy as np

        self.values = np.asarray(values)
        self.dims = tuple(dims)
        assert self.values.ndim == len(self.dims)

    shape = property(lambda self: self.values.shape)
    sizes = property(lambda self: dict(zip(self.dims, self.values.shape)))

    def __getitem__(self, key):
        out = self.values[key]
        assert out.ndim == self.values.ndim, "only slicing (no integer indexing) is supported"
        return NamedArray(out, self.dims)

    def __iter__(self):   # like DataArray: iterate over the first dimension, dropping it
        for i in range(self.values.shape[0]):
            yield NamedArray(self.values[i], self.dims[1:])

    def transpose(self, *names):
        order = [self.dims.index(n) for n in names]
        return NamedArray(self.values.transpose(order), names)

    def rename(self, **mapping):
        return NamedArray(self.values, [mapping.get(d, d) for d in self.dims])

    def chunk(self, chunks):
        return self


class NamedAssay:
    def __init__(self, tile=None):
        self._vars = {}
        if tile is not None:
            self._vars["tile"] = tile
        self.mg = type("Mg", (), {"cache": staticmethod(lambda *a, **k: None)})()

    def __contains__(self, name):
        return name in self._vars

    def __getattr__(self, name):
        try:
            return self.__dict__["_vars"][name]
        except KeyError:
            raise AttributeError(name) from None

    def __setitem__(self, name, value):
        self._vars[name] = value

    @property
    def sizes(self):
        out = {}
        for v in self._vars.values():
            out.update(v.sizes)
        return out


def _named_concat(arrays, dim, **_ignored):
    import numpy as np

    arrays = list(arrays)
    axis = arrays[0].dims.index(dim)
    return NamedArray(np.concatenate([a.values for a in arrays], axis=axis), arrays[0].dims)


_cached_stitch = None


def load_reference_stitch():
    """The reference's `src/magnify/stitch.py` loaded in place with stub imports, or None."""
    global _cached_stitch
    if _cached_stitch is not None:
        return _cached_stitch
    path = os.path.join(REFERENCE_ROOT, "src", "magnify", "stitch.py")
    if not os.path.exists(path):
        return None
    xr = types.ModuleType("xarray")
    xr.Dataset = type("Dataset", (), {})
    xr.concat = _named_concat
    pkg = types.ModuleType("magnify")
    pkg.__path__ = []
    registry = types.ModuleType("magnify.registry")
    registry.components = type("Components", (), {"register": staticmethod(lambda name: (lambda f: f))})()
    pkg.registry = registry
    stubs = {"xarray": xr, "magnify": pkg, "magnify.registry": registry}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_magnify_reference_stitch", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception:
        mod = None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached_stitch = mod
    return mod


def reference_stitch(tiles, overlap):
    """Run the reference's own Stitcher on a (C,T,R,Cc,H,W) ndarray; None if unavailable.  Raises
    what the reference raises (ValueError for bad overlaps)."""
    mod = load_reference_stitch()
    if mod is None:
        return None
    dims = ("channel", "time", "tile_row", "tile_col", "tile_y", "tile_x")
    assay = NamedAssay(NamedArray(tiles, dims))
    out = mod.Stitcher(overlap=overlap)(assay)
    image = out.image
    assert image.dims == ("channel", "time", "im_y", "im_x")
    return image.values


# ---------------------------------------------------------------------------------------------
# The reference's own BeadFinder (src/magnify/find.py:445-605) executed in place: dask.array is
# replaced by NumPy, xarray by the labelled stand-ins below, and `utils.find_circles` -- the
# stochastic centre finder, out of scope -- by a function that returns the pinned beads.
# Everything after the centres (find.py:503-605: boxes, label raster, fg/bg, crops, schema) is
# the reference's code.
# ---------------------------------------------------------------------------------------------
class LabelledArray:
    """ndarray + dim names + the owner's coordinate labels (for .sel)."""

    def __init__(self, values, dims, coords):
        import numpy as np

        self.values = values if isinstance(values, np.ndarray) else np.asarray(values)
        self.dims = tuple(dims)
        self.coords = coords

    shape = property(lambda self: self.values.shape)
    dtype = property(lambda self: self.values.dtype)
    sizes = property(lambda self: dict(zip(self.dims, self.values.shape)))

    def to_numpy(self):
        return self.values

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __iter__(self):
        return iter(self.values)

    def __len__(self):
        return len(self.values)

    def __getitem__(self, key):
        import numpy as np

        out = self.values[key]
        if isinstance(out, np.ndarray) and out.ndim > 0:
            return LabelledArray(out, tuple(f"_{i}" for i in range(out.ndim)), self.coords)
        return out

    def __setitem__(self, key, value):
        import numpy as np

        self.values[key] = np.asarray(value)

    def compute(self):
        return self

    def persist(self):
        return self

    def chunk(self, chunks=None):
        return self

    def max(self):
        return self.values.max()

    def sum(self, dim=None):
        if dim is None:
            return self.values.sum()
        axes = tuple(self.dims.index(d) for d in dim)
        return LabelledArray(self.values.sum(axis=axes), tuple(d for d in self.dims if d not in dim), self.coords)

    def _index(self, indexers):
        import numpy as np

        key, dims = [], []
        for d in self.dims:
            if d in indexers:
                k = indexers[d]
                key.append(k)
                if not np.isscalar(k):
                    dims.append(d)
            else:
                key.append(slice(None))
                dims.append(d)
        return LabelledArray(self.values[tuple(key)], dims, self.coords)

    def isel(self, **indexers):
        return self._index(indexers)

    def sel(self, **indexers):
        import numpy as np

        resolved = {}
        for d, label in indexers.items():
            labels = list(self.coords[d])
            if isinstance(label, (list, tuple, np.ndarray, LabelledArray)):
                resolved[d] = [labels.index(v) for v in list(label)]
            else:
                resolved[d] = labels.index(label)
        return self._index(resolved)


class LabelledAssay:
    def __init__(self, variables=None, coords=None):
        self._vars = dict(variables or {})
        self._coords = dict(coords or {})
        self.attrs = {}
        self.mg = type("Mg", (), {"cache": staticmethod(lambda *a, **k: None)})()

    def _wrap(self, name):
        dims, values = self._vars[name]
        return LabelledArray(values, dims, self._coords)

    def __contains__(self, name):
        return name in self._vars or name in self._coords

    def __getattr__(self, name):
        d = self.__dict__
        if name in d.get("_vars", {}):
            return self._wrap(name)
        if name in d.get("_coords", {}):
            return LabelledArray(d["_coords"][name], (name,), d["_coords"])
        raise AttributeError(name)

    def __getitem__(self, name):
        return getattr(self, name)

    def __setitem__(self, name, value):
        if isinstance(value, tuple):
            dims, values = value
        else:
            dims, values = value.dims, value.values
        self._vars[name] = (tuple(dims), values)

    @property
    def sizes(self):
        out = {}
        for dims, values in self._vars.values():
            out.update(dict(zip(dims, values.shape)))
        return out

    def assign_coords(self, **coords):
        new = LabelledAssay(self._vars, self._coords)
        for name, (dims, values) in coords.items():
            new._vars[name] = (tuple(dims), values)
        return new

    def stack(self, create_index=True, **dims):
        """xarray's Dataset.stack for ONE new dimension: every variable that has all the stacked
        dims gets them merged (row-major) into the new one, placed where the first of them was."""
        import numpy as np

        (new_dim, old), = dims.items()
        new = LabelledAssay({}, self._coords)
        for name, (vdims, values) in self._vars.items():
            if all(d in vdims for d in old):
                order = [d for d in vdims if d not in old]
                first = min(vdims.index(d) for d in old)
                perm = [vdims.index(d) for d in order[:first]] + [vdims.index(d) for d in old] + \
                       [vdims.index(d) for d in order[first:]]
                v = np.transpose(values, perm)
                shape = v.shape[:first] + (-1,) + v.shape[first + len(old):]
                new._vars[name] = (tuple(order[:first]) + (new_dim,) + tuple(order[first:]), v.reshape(shape))
            else:
                new._vars[name] = (vdims, values)
        return new

    def transpose(self, *names):
        """`.transpose("mark", ...)`: move the named leading dims to the front of every variable."""
        import numpy as np

        lead = [n for n in names if n is not Ellipsis]
        new = LabelledAssay({}, self._coords)
        for name, (vdims, values) in self._vars.items():
            front = [d for d in lead if d in vdims]
            order = front + [d for d in vdims if d not in front]
            new._vars[name] = (tuple(order), np.transpose(values, [vdims.index(d) for d in order]))
        return new


_cached_find = None


def load_reference_find():
    """The reference's `src/magnify/find.py` loaded in place (NumPy for dask.array, the real
    `utils.py`, stubs for xarray / registry / the GUI), or None."""
    global _cached_find
    if _cached_find is not None:
        return _cached_find
    path = os.path.join(REFERENCE_ROOT, "src", "magnify", "find.py")
    utils = load_reference_utils()
    if not os.path.exists(path) or utils is None:
        return None
    import numpy as np

    xr = types.ModuleType("xarray")
    xr.Dataset = type("Dataset", (), {})
    xr.DataArray = type("DataArray", (), {})
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")
    da.empty = lambda shape, dtype=float, chunks=None: np.empty(shape, dtype=dtype)
    da.empty_like = lambda a, dtype=None, chunks=None: np.empty_like(a, dtype=dtype)
    dask.array = da
    pkg = types.ModuleType("magnify")
    pkg.__path__ = []
    registry = types.ModuleType("magnify.registry")
    registry.components = type("Components", (), {"register": staticmethod(lambda name: (lambda f: f))})()
    plot = types.ModuleType("magnify.plot")
    plot.__path__ = []
    vis = types.ModuleType("magnify.plot.vis")
    vis.InteractiveUI = type("InteractiveUI", (), {})
    pkg.registry, pkg.utils = registry, utils
    stubs = {"xarray": xr, "dask": dask, "dask.array": da, "magnify": pkg, "magnify.registry": registry,
             "magnify.utils": utils, "magnify.plot": plot, "magnify.plot.vis": vis}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_magnify_reference_find", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception:
        mod = None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached_find = mod
    return mod


def reference_bead_finder(image, channels, beads, roi_length, min_bead_diameter=2, max_bead_diameter=60):
    """Run the reference's own BeadFinder.__call__ with the centre finder replaced by the pinned
    `beads` (M,3 rows of (row, col, radius)).  image (C,T,H,W); returns dict(roi, fg, bg, x, y, valid)."""
    import numpy as np

    mod = load_reference_find()
    if mod is None:
        return None
    utils = mod.utils
    pinned = np.round(np.asarray(beads)).astype(np.int32)          # what find_circles returns (utils.py:159)
    real_find_circles = utils.find_circles
    utils.find_circles = lambda img, **kwargs: (pinned, np.ones(len(pinned), dtype=np.float32))
    try:
        finder = mod.BeadFinder(min_bead_diameter=min_bead_diameter, max_bead_diameter=max_bead_diameter,
                                low_edge_quantile=0.1, high_edge_quantile=0.9, num_iter=1, min_roundness=0.3,
                                roi_length=roi_length, search_channel=channels[0], interactive=False)
        assay = LabelledAssay({"image": (("channel", "time", "im_y", "im_x"), np.asarray(image))},
                              {"channel": np.asarray(channels)})
        out = finder(assay)
    finally:
        utils.find_circles = real_find_circles
    return {k: np.asarray(getattr(out, k).values) for k in ("roi", "fg", "bg", "x", "y", "valid")}


def reference_button_finder(image, channels, tag, coarse_x, coarse_y, refine, roi_length=None,
                            min_button_diameter=16, max_button_diameter=30, chamber_diameter=60,
                            search_timestep=0):
    """Run the reference's own ButtonFinder.__call__ (find.py:55-203 incl. find_rois :308-402 and the
    copy-forward loop :143-181) with its two stochastic pieces pinned: `find_centers` returns the
    coarse grid (coarse_x, coarse_y: rows x cols), and `utils.find_circles` returns, for the
    k-th non-blank button in row-major order, `refine[k]` = (y, x, r) in ROI coordinates or None.
    image (C,T,H,W).  Returns dict(roi, fg, bg, x, y, valid) with the stacked `mark` dimension first."""
    import numpy as np

    mod = load_reference_find()
    if mod is None:
        return None
    utils = mod.utils
    calls = {"n": 0}

    def fake_find_circles(img, **kwargs):
        r = refine[calls["n"]]
        calls["n"] += 1
        if r is None:
            return np.empty((0, 3), dtype=np.int32), np.empty(0, dtype=np.float32)
        return np.asarray([r], dtype=np.int32), np.ones(1, dtype=np.float32)

    real_find_circles = utils.find_circles
    utils.find_circles = fake_find_circles
    try:
        finder = mod.ButtonFinder(row_dist=100.0, col_dist=100.0, min_button_diameter=min_button_diameter,
                                  max_button_diameter=max_button_diameter, chamber_diameter=chamber_diameter,
                                  top_chamber=None, left_chamber=None, low_edge_quantile=0.1, high_edge_quantile=0.9,
                                  num_iter=1000, min_roundness=0.2, cluster_penalty=10, roi_length=roi_length,
                                  progress_bar=False, search_timestep=search_timestep, search_channel=channels[0],
                                  interactive=False)
        finder.find_centers = lambda images, assay: (np.array(coarse_x, dtype=float), np.array(coarse_y, dtype=float))
        image = np.asarray(image)
        c, t, h, w = image.shape
        rows, cols = np.asarray(tag).shape
        assay = LabelledAssay(
            {"image": (("channel", "time", "im_y", "im_x"), image),
             "tag": (("mark_row", "mark_col"), np.asarray(tag)),
             "valid": (("mark_row", "mark_col", "time"), np.ones((rows, cols, t), dtype=bool))},
            {"channel": np.asarray(channels), "time": np.arange(t)})
        out = finder(assay)
    finally:
        utils.find_circles = real_find_circles
    return {k: np.asarray(getattr(out, k).values) for k in ("roi", "fg", "bg", "x", "y", "valid")}


# ---------------------------------------------------------------------------------------------
# The reference's own `filter_expression` (src/magnify/filter.py:11-37) executed in place.  The
# function body -- which variables are reduced over which dims, the pairwise-difference
# statistic, the threshold, the and/or with `valid` -- is the reference's code; the xarray
# arithmetic it calls (`where` with NaN fill, skip-NaN `median`) is not in /root/reference and is
# restated here from xarray's published semantics: `where` on an integer array of <= 16 bits
# promotes to float32 (xarray.core.dtypes.maybe_promote), wider integers to float64, and
# `median(dim=...)` skips NaN (nanmedian).  `promote` lets the tests check that the outcome does
# not depend on that promotion rule (medians of u16 values are exact in float32 and float64).
# ---------------------------------------------------------------------------------------------
class ReducibleArray(LabelledArray):
    promote = None   # None = xarray's rule; or a NumPy float dtype

    def _like(self, values, dims=None):
        out = ReducibleArray(values, self.dims if dims is None else dims, self.coords)
        out.promote = self.promote
        return out

    def _index(self, indexers):
        base = LabelledArray._index(self, indexers)
        return self._like(base.values, base.dims)

    def where(self, cond):
        import numpy as np

        dt = self.values.dtype
        if self.promote is not None:
            ft = np.dtype(self.promote)
        elif np.issubdtype(dt, np.integer):
            ft = np.dtype(np.float32 if dt.itemsize <= 2 else np.float64)
        else:
            ft = dt
        c = np.asarray(cond.values if isinstance(cond, LabelledArray) else cond, dtype=bool)
        return self._like(np.where(c, self.values.astype(ft), ft.type(np.nan)))

    def _reduce(self, fn, dim):
        import warnings

        axes = tuple(self.dims.index(d) for d in dim)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)   # all-NaN slices -> NaN, like xarray
            out = fn(self.values, axis=axes)
        return self._like(out, tuple(d for d in self.dims if d not in dim))

    def median(self, dim):
        import numpy as np

        return self._reduce(np.nanmedian, dim)

    def mean(self, dim):
        import numpy as np

        return self._reduce(np.nanmean, dim)

    def __sub__(self, other):
        return self._like(self.values - other.values)

    def __gt__(self, other):
        return self._like(self.values > other)

    def __lt__(self, other):
        return self._like(self.values < other)


class ValidArray(LabelledArray):
    """`assay.valid`: (mark, time) bool; `|=` with a (mark,) array broadcasts by dim name."""

    @property
    def data(self):
        return self.values

    @data.setter
    def data(self, value):
        self.values[...] = value

    def __ior__(self, other):
        self.values |= other.values.reshape((-1,) + (1,) * (self.values.ndim - 1))
        return self

    def __iand__(self, other):
        self.values &= other.values
        return self


class FilterAssay:
    def __init__(self, roi, fg, bg, valid, channels, promote=None, tag=None, mark_row=None):
        self._coords = {"channel": list(channels)}
        self._roi, self._fg, self._bg, self._valid = roi, fg, bg, valid
        self._promote = promote
        self._sel = {}
        self._tag, self._mark_row = tag, mark_row

    tag = property(lambda self: LabelledArray(self._tag, ("mark",), self._coords))
    mark_row = property(lambda self: LabelledArray(self._mark_row, ("mark",), self._coords))

    channel = property(lambda self: list(self._coords["channel"]))
    sizes = property(lambda self: {"mark": self._roi.shape[0]})

    def _view(self, **sel):
        new = FilterAssay(self._roi, self._fg, self._bg, self._valid, self._coords["channel"], self._promote,
                          self._tag, self._mark_row)
        new._sel = {**self._sel, **sel}
        return new

    def isel(self, time):
        return self._view(time=time)

    def sel(self, channel):
        return self._view(channel=self._coords["channel"].index(channel))

    @property
    def roi(self):
        out = IntensityArray(self._roi, ("mark", "channel", "time", "roi_y", "roi_x"), self._coords)
        out.promote = self._promote
        out.masks = {"fg": (("mark", "time", "roi_y", "roi_x"), self._fg), "bg": (("mark", "time", "roi_y", "roi_x"), self._bg)}
        return out._index(self._sel)

    @property
    def fg(self):
        return LabelledArray(self._fg, ("mark", "time", "roi_y", "roi_x"), self._coords)._index(
            {k: v for k, v in self._sel.items() if k == "time"})

    @property
    def bg(self):
        return LabelledArray(self._bg, ("mark", "time", "roi_y", "roi_x"), self._coords)._index(
            {k: v for k, v in self._sel.items() if k == "time"})

    @property
    def valid(self):
        return ValidArray(self._valid, ("mark", "time"), self._coords)

    def __getitem__(self, name):
        return getattr(self, name)

    def __setitem__(self, name, value):
        assert name == "valid"
        self._valid = value.values


_cached_filter = None


def load_reference_filter():
    global _cached_filter
    if _cached_filter is not None:
        return _cached_filter
    path = os.path.join(REFERENCE_ROOT, "src", "magnify", "filter.py")
    utils = load_reference_utils()
    if not os.path.exists(path) or utils is None:
        return None
    import numpy as np

    xr = types.ModuleType("xarray")
    xr.Dataset = type("Dataset", (), {})
    xr.zeros_like = lambda a, dtype=None: ValidArray(np.zeros_like(a.values, dtype=dtype), a.dims, a.coords)
    pkg = types.ModuleType("magnify")
    pkg.__path__ = []
    registry = types.ModuleType("magnify.registry")
    registry.component = lambda name: (lambda f: f)
    pkg.registry, pkg.utils = registry, utils
    stubs = {"xarray": xr, "magnify": pkg, "magnify.registry": registry, "magnify.utils": utils}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_magnify_reference_filter", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception:
        mod = None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached_filter = mod
    return mod


def reference_filter_expression(roi, fg, bg, valid, channels, search_channel=None, min_contrast=None, promote=None):
    """Run the reference's `filter_expression` on (M,C,T,L,L) roi, (M,T,L,L) fg/bg, (M,T) valid.
    Returns the new `valid` (M,T), or None when /root/reference is absent."""
    mod = load_reference_filter()
    if mod is None:
        return None
    assay = FilterAssay(roi, fg, bg, valid.copy(), channels, promote)
    out = mod.filter_expression(assay, search_channel=search_channel, min_contrast=min_contrast)
    return out._valid


# ---------------------------------------------------------------------------------------------
# identify.py:76-80 -- the intensities `identify_mrbles` starts from.  The rest of that function
# (pandas / scipy / sklearn clustering) is out of scope, so only the statement that builds `sel`
# and `intensities` is executed, read from the reference file at run time (located by its text,
# not copied), on the same stand-ins as above.
# ---------------------------------------------------------------------------------------------
class IntensityArray(ReducibleArray):
    """`assay.roi` with its fg/bg coordinates attached (DataArray attribute access `sel.fg`)."""
    masks = None   # {"fg": (dims, values), "bg": (dims, values)}

    def _like(self, values, dims=None):
        out = IntensityArray(values, self.dims if dims is None else dims, self.coords)
        out.promote, out.masks = self.promote, self.masks
        return out

    def _index(self, indexers):
        import numpy as np

        out = ReducibleArray._index(self, indexers)
        masks = {}
        for name, (dims, values) in self.masks.items():
            sub = LabelledArray(values, dims, self.coords)._index({k: v for k, v in indexers.items() if k in dims})
            masks[name] = (sub.dims, sub.values)
        out.masks = masks
        return out

    def __getattr__(self, name):
        masks = self.__dict__.get("masks") or {}
        if name in masks:
            return LabelledArray(masks[name][1], masks[name][0], self.coords)
        raise AttributeError(name)

    def where(self, cond):
        import numpy as np

        # broadcast the condition to self.dims by dimension NAME (xarray alignment)
        c = np.asarray(cond.values, dtype=bool)
        shape = [self.values.shape[self.dims.index(d)] if d in cond.dims else 1 for d in self.dims]
        order = [cond.dims.index(d) for d in self.dims if d in cond.dims]
        c = np.transpose(c, order).reshape(shape)
        return ReducibleArray.where(self, np.broadcast_to(c, self.values.shape))


def reference_filter_leaky(roi, fg, bg, valid, channels, tag, mark_row, search_channel=None, promote=None):
    """Run the reference's `filter_leaky_buttons` (filter.py:65-94); returns the new valid (M,T)."""
    mod = load_reference_filter()
    if mod is None:
        return None
    assay = FilterAssay(roi, fg, bg, valid.copy(), channels, promote, tag, mark_row)
    out = mod.filter_leaky_buttons(assay, search_channel=search_channel)
    return out._valid


def reference_filter_nonround(fg, valid, channels, min_roundness=0.75, search_channel=None):
    """Run the reference's `filter_nonround` (filter.py:40-62; real cv2) -> new valid (M,T)."""
    import numpy as np

    mod = load_reference_filter()
    if mod is None:
        return None
    m, t, length = fg.shape[0], fg.shape[1], fg.shape[-1]
    roi = np.zeros((m, len(channels), t, length, length), np.uint16)
    assay = FilterAssay(roi, fg, fg, valid.copy(), channels)
    out = mod.filter_nonround(assay, min_roundness=min_roundness, search_channel=search_channel)
    return out._valid


def reference_mrbles_intensities(roi, fg, bg, channel_names, channels, promote=None):
    """(mark, len(channels)) intensities by the reference's own expression, or None."""
    import textwrap

    path = os.path.join(REFERENCE_ROOT, "src", "magnify", "identify.py")
    if not os.path.exists(path):
        return None
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip().startswith("sel = assay.roi.isel(time=0)"))
    stop = next(i for i in range(start, len(lines)) if lines[i].strip() == ").to_numpy()")
    code = compile(textwrap.dedent("\n".join(lines[start:stop + 1])), path, "exec")
    arr = IntensityArray(roi, ("mark", "channel", "time", "roi_y", "roi_x"), {"channel": list(channel_names)})
    arr.promote = promote
    arr.masks = {"fg": (("mark", "time", "roi_y", "roi_x"), fg), "bg": (("mark", "time", "roi_y", "roi_x"), bg)}
    scope = {"assay": type("A", (), {"roi": arr})(), "channels": list(channels)}
    exec(code, scope)
    return scope["intensities"]


# ---------------------------------------------------------------------------------------------
# reader.py:80-160 `extract_paths` -- pure Python (re / glob / fnmatch), loaded in place with stub
# modules for the imports the function does not use (bs4, dask, tifffile, xarray).
# ---------------------------------------------------------------------------------------------
_cached_reader = None


def load_reference_reader():
    global _cached_reader
    if _cached_reader is not None:
        return _cached_reader
    path = os.path.join(REFERENCE_ROOT, "src", "magnify", "reader.py")
    utils = load_reference_utils()
    if not os.path.exists(path) or utils is None:
        return None
    xr = types.ModuleType("xarray")
    xr.Dataset = type("Dataset", (), {})
    xr.DataArray = type("DataArray", (), {})
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")
    dask.array = da
    pkg = types.ModuleType("magnify")
    pkg.__path__ = []
    registry = types.ModuleType("magnify.registry")
    registry.readers = type("Readers", (), {"register": staticmethod(lambda name: (lambda f: f))})()
    pkg.registry, pkg.utils = registry, utils
    stubs = {"xarray": xr, "dask": dask, "dask.array": da, "bs4": types.ModuleType("bs4"),
             "tifffile": types.ModuleType("tifffile"), "magnify": pkg, "magnify.registry": registry,
             "magnify.utils": utils}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_magnify_reference_reader", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception:
        mod = None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached_reader = mod
    return mod


# ---------------------------------------------------------------------------------------------
# utils.find_circles (utils.py:100-218) itself, run with a capturing stand-in for the napari GUI:
# `gui.run_widget(fn, ...)` is how the function obtains `edges` and the filtered circles when a
# GUI is present (:135-136, :211-212), so a stand-in that just calls `fn()` exposes both stages.
# ---------------------------------------------------------------------------------------------
class CapturingGui:
    def __init__(self):
        self.stages = []

    def run_widget(self, fn, auto_call=True, last=False):
        out = fn()
        self.stages.append(out)
        return out


def reference_find_circles_stages(img, low_edge_quantile, high_edge_quantile, grid_length, num_iter, min_radius,
                                  max_radius, min_roundness, min_dist):
    """(edges 0/1 uint8, circles, scores) from the reference's own find_circles, or None."""
    utils = load_reference_utils()
    if utils is None:
        return None
    gui = CapturingGui()
    circles, scores = utils.find_circles(img, low_edge_quantile=low_edge_quantile, high_edge_quantile=high_edge_quantile,
                                         grid_length=grid_length, num_iter=num_iter, min_radius=min_radius,
                                         max_radius=max_radius, min_roundness=min_roundness, min_dist=min_dist, gui=gui)
    edges = gui.stages[0][1][0]
    return edges, circles, scores


# ---------------------------------------------------------------------------------------------
# identify.py:13-45 `identify_buttons` (pandas pinlist parsing) loaded in place.
# ---------------------------------------------------------------------------------------------
_cached_identify = None


def load_reference_identify():
    global _cached_identify
    if _cached_identify is not None:
        return _cached_identify
    path = os.path.join(REFERENCE_ROOT, "src", "magnify", "identify.py")
    if not os.path.exists(path):
        return None
    try:
        import numba  # noqa: F401
        import pandas  # noqa: F401
        import scipy  # noqa: F401
    except Exception:
        return None
    pkg = types.ModuleType("magnify")
    pkg.__path__ = []
    registry = types.ModuleType("magnify.registry")
    registry.component = lambda name: (lambda f: f)
    pkg.registry = registry
    stubs = {"magnify": pkg, "magnify.registry": registry}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_magnify_reference_identify", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception:
        mod = None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached_identify = mod
    return mod


def reference_identify_buttons(num_times, shape=None, pinlist=None, blank=None):
    """(tag (rows, cols), valid (rows, cols, T)) from the reference's identify_buttons, or None."""
    import numpy as np

    mod = load_reference_identify()
    if mod is None:
        return None
    assay = LabelledAssay({"image": (("channel", "time", "im_y", "im_x"), np.zeros((1, num_times, 2, 2), np.uint8))}, {})
    out = mod.identify_buttons(assay, shape=shape, pinlist=pinlist, blank=blank)
    return out._vars["tag"][1], out._vars["valid"][1]


# ---------------------------------------------------------------------------------------------
# The WHOLE reference package, unmodified, imported in place: `magnify` is created as a package
# whose __path__ is /root/reference/src/magnify, so every submodule (registry.py, pipeline.py,
# preprocess.py, stitch.py, find.py, identify.py, postprocess.py, reader.py ...) is executed from
# the reference's own files by the normal import system.  Only the third-party packages that
# are not installed here are stood in for:
#   xarray        -> magnify_b200.dataset (the subset of the xarray API the pipeline uses)
#   dask.array    -> NumPy (`empty`, `empty_like`; nothing is lazy)
#   catalogue     -> a 20-line registry (create / register / get / get_all)
#   tifffile, bs4 -> empty modules (only the TIFF reader needs them)
#   napari, magnify.plot -> empty modules (GUI)
# This is how the drop-in claim is exercised: `magnify_b200.components.install()` registers the
# GPU components in the reference's own registry, and the reference's own `*_pipe` builders and
# `Pipeline.__call__` (pipeline.py:14-29) run them.
# ---------------------------------------------------------------------------------------------
class CatalogueRegistry:
    """catalogue.Registry: `register(name)` is a decorator (or `register(name, func=f)`), `get`
    raises for unknown names, re-registering a name replaces it."""

    def __init__(self, namespace):
        self.namespace = tuple(namespace)
        self._items = {}

    def register(self, name, *, func=None):
        def do(f):
            self._items[name] = f
            return f

        return do(func) if func is not None else do

    def get(self, name):
        if name not in self._items:
            raise KeyError(f"Cant't find '{name}' in registry {' -> '.join(self.namespace)}. "
                           f"Available names: {', '.join(sorted(self._items)) or 'none'}")
        return self._items[name]

    def get_all(self):
        return dict(self._items)

    def __contains__(self, name):
        return name in self._items


_cached_pkg = None


def load_reference_package():
    """Import the reference package in place (see above).  Returns the `magnify` module, or None
    when /root/reference is absent.  The stand-in modules stay in sys.modules afterwards (the
    reference's functions look `xr`, `da` ... up at call time through their module globals,
    which are bound at import, so this is only for later `import magnify.x` statements)."""
    global _cached_pkg
    if _cached_pkg is not None:
        return _cached_pkg
    root = os.path.join(REFERENCE_ROOT, "src", "magnify")
    if not os.path.isdir(root):
        return None
    try:
        import cv2  # noqa: F401
        import numba  # noqa: F401
        import pandas  # noqa: F401
        import scipy  # noqa: F401
        import tqdm  # noqa: F401
    except Exception:
        return None
    import numpy as np

    from magnify_b200 import dataset as xr_standin

    mods = {}
    mods["xarray"] = xr_standin
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")
    da.Array = type("Array", (), {})
    da.empty = lambda shape, dtype=float, chunks=None: np.empty(shape, dtype=dtype)
    da.empty_like = lambda a, dtype=None, chunks=None: np.empty_like(np.asarray(a), dtype=dtype)
    dask.array = da
    mods["dask"], mods["dask.array"] = dask, da
    catalogue = types.ModuleType("catalogue")
    catalogue.create = lambda *namespace, entry_points=False: CatalogueRegistry(namespace)
    mods["catalogue"] = catalogue
    mods["tifffile"] = types.ModuleType("tifffile")
    mods["bs4"] = types.ModuleType("bs4")
    napari = types.ModuleType("napari")
    napari_types = types.ModuleType("napari.types")
    napari_types.LayerDataTuple = tuple
    napari.types = napari_types
    mods["napari"], mods["napari.types"] = napari, napari_types
    plot = types.ModuleType("magnify.plot")
    plot.__path__ = []
    vis = types.ModuleType("magnify.plot.vis")
    vis.InteractiveUI = type("InteractiveUI", (), {})
    plot.vis = vis
    mods["magnify.plot"], mods["magnify.plot.vis"] = plot, vis
    for k, v in mods.items():
        sys.modules.setdefault(k, v)
    if sys.modules["xarray"] is not xr_standin:
        return None       # a real xarray is installed: import magnify normally instead
    spec = importlib.util.spec_from_file_location("magnify", os.path.join(root, "__init__.py"),
                                                  submodule_search_locations=[root])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["magnify"] = pkg
    pkg.plot = plot
    try:
        spec.loader.exec_module(pkg)
    except Exception:
        sys.modules.pop("magnify", None)
        for k in [k for k in sys.modules if k.startswith("magnify.")]:
            sys.modules.pop(k, None)
        raise
    _cached_pkg = pkg
    return pkg

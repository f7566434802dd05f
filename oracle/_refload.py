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

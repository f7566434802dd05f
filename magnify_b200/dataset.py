"""Labelled containers with the slice of the xarray API that magnify's components use.

magnify's component contract is `xarray.Dataset -> xarray.Dataset`.  xarray
is not installed in the build image nor on the GPU boxes, so this module provides `Dataset` and
`DataArray` with xarray's semantics for exactly the calls the reference's pipeline makes
(`preprocess.py`, `stitch.py`, `find.py`, `identify.py:13-47`, `postprocess.py`): named dims,
coordinates, attribute access, positional / label indexing with write-through views, `stack` /
`unstack` of one multi-index, `transpose` with Ellipsis, `concat`, NaN-skipping reductions.
`magnify_b200.components` is written against the xarray API only: it runs unchanged on real
xarray objects, and on these when xarray is absent.  The tests also register this module under
the name `xarray` to execute the reference's own, unmodified package in place (oracle/_refload.py).

Like xarray, a variable keeps "duck arrays" (objects with `__array_function__`, such as
`magnify_b200.devarray.DeviceArray`) as they are instead of converting them to NumPy, so
device-resident results travel through a pipeline without a copy until somebody reads them.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, Mapping, Optional, Sequence, Tuple

import numpy as np


def _is_duck(data) -> bool:
    return not isinstance(data, np.ndarray) and hasattr(data, "__array_function__") and hasattr(data, "shape")


def _as_data(data):
    if isinstance(data, (DataArray, Variable)):
        return data.data
    if _is_duck(data):
        return data
    return np.asarray(data)


class Variable:
    """dims + data (+ attrs): what a Dataset stores per name."""

    __slots__ = ("dims", "data", "attrs")

    def __init__(self, dims, data, attrs=None):
        if isinstance(dims, str):
            dims = (dims,)
        self.dims = tuple(dims)
        self.data = _as_data(data)
        self.attrs = dict(attrs or {})
        if len(self.dims) != len(self.data.shape):
            raise ValueError(f"dimensions {self.dims} must have the same length as the number of data "
                             f"dimensions, ndim={len(self.data.shape)}")

    shape = property(lambda self: tuple(self.data.shape))
    dtype = property(lambda self: self.data.dtype)
    ndim = property(lambda self: len(self.dims))
    sizes = property(lambda self: dict(zip(self.dims, self.data.shape)))
    values = property(lambda self: np.asarray(self.data))

    def transpose(self, *dims):
        order = _expand_ellipsis(dims, self.dims)
        if tuple(order) == self.dims:
            return Variable(self.dims, self.data, self.attrs)
        return Variable(order, np.transpose(self.data, [self.dims.index(d) for d in order]), self.attrs)

    def copy(self):
        return Variable(self.dims, self.data, self.attrs)


def _expand_ellipsis(dims, have) -> Tuple[str, ...]:
    dims = list(dims)
    if not dims:
        return tuple(reversed(have))
    if Ellipsis in dims:
        i = dims.index(Ellipsis)
        rest = [d for d in have if d not in dims]
        dims = dims[:i] + rest + dims[i + 1:]
    missing = [d for d in dims if d not in have]
    if missing or len(dims) != len(have):
        raise ValueError(f"{tuple(dims)} must be a permuted list of {tuple(have)}, unless `...` is included")
    return tuple(dims)


def _as_variable(value) -> Variable:
    if isinstance(value, Variable):
        return value
    if isinstance(value, DataArray):
        return value.variable
    if isinstance(value, tuple):
        return Variable(*value)
    data = np.asarray(value)
    if data.ndim == 0:
        return Variable((), data)
    raise TypeError("cannot set a variable without dimension names: pass (dims, data) or a DataArray")


def _indexer_dims(key, dims):
    """Normalise a positional key to one entry per dim; returns (full key, surviving dims)."""
    if not isinstance(key, tuple):
        key = (key,)
    if any(k is Ellipsis for k in key):
        i = next(j for j, k in enumerate(key) if k is Ellipsis)
        fill = len(dims) - (len(key) - 1)
        key = key[:i] + (slice(None),) * fill + key[i + 1:]
    key = key + (slice(None),) * (len(dims) - len(key))
    out_dims = []
    for k, d in zip(key, dims):
        if isinstance(k, (DataArray, Variable)):
            k = np.asarray(k.data)
        if isinstance(k, (int, np.integer)) or (isinstance(k, np.ndarray) and k.ndim == 0):
            continue
        out_dims.append(d)
    key = tuple(np.asarray(k.data) if isinstance(k, (DataArray, Variable)) else k for k in key)
    key = tuple(int(k) if isinstance(k, np.ndarray) and k.ndim == 0 else k for k in key)
    return key, tuple(out_dims)


def _apply_key(data, key):
    """Orthogonal (per-dimension) indexing like xarray's, also for several array indexers."""
    arrays = [i for i, k in enumerate(key) if isinstance(k, (list, np.ndarray))]
    if len(arrays) <= 1:
        return data[key]
    out = data
    for axis in range(len(key) - 1, -1, -1):   # right to left keeps the axis numbers valid
        k = key[axis]
        if isinstance(k, slice) and k == slice(None):
            continue
        out = out[(slice(None),) * axis + (k,)]
    return out


class DataArray:
    """A Variable plus the coordinates that live on (a subset of) its dims."""

    def __init__(self, data=None, coords: Optional[Mapping[str, Any]] = None, dims=None, name=None, attrs=None):
        if isinstance(data, Variable):
            self.variable = data
        else:
            data = _as_data(data)
            if dims is None:
                dims = tuple(f"dim_{i}" for i in range(len(data.shape)))
            self.variable = Variable(dims, data, attrs)
        self._coords: Dict[str, Variable] = {}
        for k, v in (coords or {}).items():
            if isinstance(v, (DataArray, Variable, tuple)):
                self._coords[k] = _as_variable(v)
            else:
                arr = np.asarray(v)
                self._coords[k] = Variable((k,) if arr.ndim else (), arr)
        self.name = name
        self._stacked: Dict[str, Tuple[str, ...]] = {}

    # -- basic properties -----------------------------------------------------------------------
    dims = property(lambda self: self.variable.dims)
    @property
    def data(self):
        return self.variable.data

    @data.setter
    def data(self, value):
        """xarray's `da.data = array`: the array behind the variable is replaced in place, so a
        DataArray taken out of a Dataset writes through to it (filter.py:60, 92 rely on that)."""
        value = _as_data(value)
        if tuple(value.shape) != self.variable.shape:
            raise ValueError(f"replacement data must match the Variable's shape. replacement data has shape "
                             f"{tuple(value.shape)}; Variable has shape {self.variable.shape}")
        self.variable.data = value

    shape = property(lambda self: self.variable.shape)
    dtype = property(lambda self: self.variable.dtype)
    ndim = property(lambda self: self.variable.ndim)
    sizes = property(lambda self: self.variable.sizes)
    size = property(lambda self: int(np.prod(self.variable.shape, dtype=np.int64)))
    attrs = property(lambda self: self.variable.attrs)
    @property
    def values(self):
        return np.asarray(self.variable.data)

    @values.setter
    def values(self, value):
        self.data = np.asarray(value)

    coords = property(lambda self: {k: DataArray(v, name=k) for k, v in self._coords.items()})

    def _new(self, variable: Variable, coords=None, name="__same__") -> "DataArray":
        out = DataArray(variable, name=self.name if name == "__same__" else name)
        out._coords = dict(self._coords if coords is None else coords)
        out._stacked = dict(self._stacked)
        return out

    def to_numpy(self) -> np.ndarray:
        return np.asarray(self.variable.data)

    def __array__(self, dtype=None, copy=None):
        arr = np.asarray(self.variable.data)
        return arr if dtype is None else arr.astype(dtype)

    def item(self):
        return np.asarray(self.variable.data).item()

    def __len__(self):
        if not self.dims:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def __iter__(self):
        if not self.dims:
            raise TypeError("iteration over a 0-d array")
        for i in range(self.shape[0]):
            yield self[i]

    def __bool__(self):
        return bool(np.asarray(self.variable.data))

    def __index__(self):
        return int(np.asarray(self.variable.data))

    def __hash__(self):
        return id(self)

    def __getattr__(self, name):
        coords = self.__dict__.get("_coords", {})
        if name in coords:
            return DataArray(coords[name], name=name)._with_coords_of(self, coords[name].dims)
        var = self.__dict__.get("variable")
        if var is not None:
            if name in var.dims:      # a dimension without a coordinate reads as 0..n-1
                return DataArray(Variable((name,), np.arange(var.sizes[name])), name=name)
            if name in var.attrs:
                return var.attrs[name]
        raise AttributeError(f"{type(self).__name__!r} object has no attribute {name!r}")

    def _with_coords_of(self, other: "DataArray", dims) -> "DataArray":
        self._coords = {k: v for k, v in other._coords.items() if set(v.dims) <= set(dims)}
        return self

    def __repr__(self):
        return f"<magnify_b200.DataArray {self.name!r} {dict(self.sizes)} {self.dtype}>"

    # -- indexing -------------------------------------------------------------------------------
    def __getitem__(self, key) -> "DataArray":
        if isinstance(key, str):
            return getattr(self, key)
        full, out_dims = _indexer_dims(key, self.dims)
        var = Variable(out_dims, _apply_key(self._indexable(), full), self.variable.attrs)
        coords = {}
        for name, cv in self._coords.items():
            ckey = tuple(full[self.dims.index(d)] for d in cv.dims) if all(d in self.dims for d in cv.dims) else None
            if ckey is None:
                continue
            _, cdims = _indexer_dims(ckey, cv.dims)
            coords[name] = Variable(cdims, _apply_key(np.asarray(cv.data), ckey), cv.attrs)
        return self._new(var, coords)

    def _indexable(self):
        return self.variable.data      # a duck array decides itself what it can index lazily

    def __setitem__(self, key, value) -> None:
        full, _ = _indexer_dims(key, self.dims)
        data = self.variable.data
        if not isinstance(data, np.ndarray):
            data = self.variable.data = np.array(data)    # materialise a lazy array before writing
        if isinstance(value, (DataArray, Variable)):
            value = np.asarray(value.data)
        data[full] = value

    def _positions(self, dim: str, label):
        """Positions along `dim` of coordinate label(s); a dim without coordinate is positional."""
        if isinstance(label, (DataArray, Variable)):
            label = np.asarray(label.data)
        if isinstance(label, slice):
            raise NotImplementedError("label slices are not supported")
        index = self._coords.get(dim)
        if index is None or index.dims != (dim,):
            return label.tolist() if isinstance(label, np.ndarray) and label.ndim else (
                int(label) if not isinstance(label, (list, tuple)) else list(label))
        values = np.asarray(index.data)

        def one(v):
            hits = np.nonzero(values == v)[0]
            if len(hits) == 0:
                raise KeyError(f"{v!r} is not a label of dimension {dim!r}")
            return int(hits[0])

        if isinstance(label, (list, tuple)) or (isinstance(label, np.ndarray) and label.ndim):
            return [one(v) for v in list(label)]
        return one(label.item() if isinstance(label, np.ndarray) else label)

    def isel(self, indexers: Optional[Mapping[str, Any]] = None, **kw) -> "DataArray":
        indexers = {**(indexers or {}), **kw}
        unknown = [d for d in indexers if d not in self.dims]
        if unknown:
            raise ValueError(f"dimensions {unknown} do not exist; expected one or more of {self.dims}")
        return self[tuple(indexers.get(d, slice(None)) for d in self.dims)]

    def sel(self, indexers: Optional[Mapping[str, Any]] = None, **kw) -> "DataArray":
        indexers = {**(indexers or {}), **kw}
        return self.isel({d: self._positions(d, v) for d, v in indexers.items()})

    # -- reshaping ------------------------------------------------------------------------------
    def transpose(self, *dims) -> "DataArray":
        return self._new(self.variable.transpose(*dims))

    T = property(lambda self: self.transpose())

    def rename(self, new_name_or_dims=None, **kw) -> "DataArray":
        mapping = dict(new_name_or_dims or {}, **kw) if not isinstance(new_name_or_dims, str) else kw
        var = Variable([mapping.get(d, d) for d in self.dims], self.variable.data, self.variable.attrs)
        coords = {mapping.get(k, k): Variable([mapping.get(d, d) for d in v.dims], v.data, v.attrs)
                  for k, v in self._coords.items()}
        out = self._new(var, coords)
        if isinstance(new_name_or_dims, str):
            out.name = new_name_or_dims
        return out

    def expand_dims(self, dim, axis=0) -> "DataArray":
        dims = [dim] if isinstance(dim, str) else list(dim)
        data = self.variable.data
        for k, d in enumerate(dims):
            data = np.expand_dims(data, axis + k)
        new_dims = list(self.dims)
        new_dims[axis:axis] = dims
        return self._new(Variable(new_dims, data, self.variable.attrs))

    def squeeze(self, dim=None) -> "DataArray":
        dims = [d for d, n in self.sizes.items() if n == 1] if dim is None else ([dim] if isinstance(dim, str) else list(dim))
        for d in dims:
            if self.sizes[d] != 1:
                raise ValueError(f"cannot select a dimension to squeeze out which has length greater than one: {d}")
        return self.isel({d: 0 for d in dims})

    def stack(self, dimensions=None, create_index=True, **kw) -> "DataArray":
        ds = Dataset({"__v__": self}).stack(dimensions, create_index=create_index, **kw)
        return ds["__v__"].rename(None)._renamed(self.name)

    def unstack(self) -> "DataArray":
        ds = Dataset({"__v__": self})
        ds._stacked = dict(self._stacked)
        return ds.unstack()["__v__"]._renamed(self.name)

    def _renamed(self, name):
        self.name = name
        return self

    def chunk(self, *a, **k) -> "DataArray":
        return self

    compute = persist = load = chunk

    def copy(self, deep=True) -> "DataArray":
        data = np.array(self.variable.data) if deep else self.variable.data
        return self._new(Variable(self.dims, data, self.variable.attrs))

    def astype(self, dtype) -> "DataArray":
        return self._new(Variable(self.dims, np.asarray(self.variable.data).astype(dtype), self.variable.attrs))

    def assign_attrs(self, *args, **kw) -> "DataArray":
        out = self._new(Variable(self.dims, self.variable.data, self.variable.attrs))
        out.variable.attrs.update(*args, **kw)
        return out

    def assign_coords(self, coords=None, **kw) -> "DataArray":
        out = self._new(self.variable)
        for k, v in {**(coords or {}), **kw}.items():
            out._coords[k] = _as_variable(v) if isinstance(v, (tuple, DataArray, Variable)) else Variable((k,), np.asarray(v))
        return out

    def drop_vars(self, names, errors="raise") -> "DataArray":
        names = [names] if isinstance(names, str) else list(names)
        return self._new(self.variable, {k: v for k, v in self._coords.items() if k not in names})

    # -- arithmetic -----------------------------------------------------------------------------
    def _binary(self, other, op, reflexive=False) -> "DataArray":
        if isinstance(other, DataArray):
            dims = list(self.dims) + [d for d in other.dims if d not in self.dims]
            a = _broadcast_to_dims(self, dims)
            b = _broadcast_to_dims(other, dims)
            coords = {**other._coords, **self._coords}
        else:
            dims, a, b, coords = list(self.dims), np.asarray(self.variable.data), other, self._coords
        with np.errstate(all="ignore"):
            out = op(b, a) if reflexive else op(a, b)
        return self._new(Variable(dims, out), coords)

    def __add__(self, o): return self._binary(o, np.add)
    def __radd__(self, o): return self._binary(o, np.add, True)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __rsub__(self, o): return self._binary(o, np.subtract, True)
    def __mul__(self, o): return self._binary(o, np.multiply)
    def __rmul__(self, o): return self._binary(o, np.multiply, True)
    def __truediv__(self, o): return self._binary(o, np.true_divide)
    def __rtruediv__(self, o): return self._binary(o, np.true_divide, True)
    def __and__(self, o): return self._binary(o, np.logical_and if self.dtype == bool else np.bitwise_and)
    def __or__(self, o): return self._binary(o, np.logical_or if self.dtype == bool else np.bitwise_or)
    def __invert__(self): return self._new(Variable(self.dims, ~np.asarray(self.variable.data)))
    def __neg__(self): return self._new(Variable(self.dims, -np.asarray(self.variable.data)))
    def __eq__(self, o): return self._binary(o, np.equal)          # noqa: E704
    def __ne__(self, o): return self._binary(o, np.not_equal)
    def __lt__(self, o): return self._binary(o, np.less)
    def __le__(self, o): return self._binary(o, np.less_equal)
    def __gt__(self, o): return self._binary(o, np.greater)
    def __ge__(self, o): return self._binary(o, np.greater_equal)

    def _inplace(self, other, op):
        new = self._binary(other, op)
        if new.dims != self.dims:
            new = new.transpose(*self.dims) if set(new.dims) == set(self.dims) else new
        if new.dims != self.dims:
            raise ValueError("in-place operation would change the dimensions")
        self[...] = new.variable.data
        return self

    def __iand__(self, o): return self._inplace(o, np.logical_and if self.dtype == bool else np.bitwise_and)
    def __ior__(self, o): return self._inplace(o, np.logical_or if self.dtype == bool else np.bitwise_or)

    def clip(self, min=None, max=None) -> "DataArray":
        return self._new(Variable(self.dims, np.clip(np.asarray(self.variable.data), min, max)))

    def where(self, cond, other=None) -> "DataArray":
        """Keep values where cond, NaN elsewhere (integers of <= 16 bits promote to float32, wider
        ones to float64: xarray.core.dtypes.maybe_promote)."""
        dims = list(self.dims)
        if isinstance(cond, DataArray):
            dims += [d for d in cond.dims if d not in dims]
            c = _broadcast_to_dims(cond, dims).astype(bool)
        else:
            c = np.asarray(cond, dtype=bool)
        a = _broadcast_to_dims(self, dims)
        if other is None:
            if np.issubdtype(a.dtype, np.integer) or a.dtype == bool:
                a = a.astype(np.float32 if a.dtype.itemsize <= 2 else np.float64)
            fill = np.array(np.nan, dtype=a.dtype)
        else:
            fill = other
        return self._new(Variable(dims, np.where(c, a, fill)))

    def _reduce(self, fn, nanfn, dim=None, **kw) -> "DataArray":
        import warnings

        data = np.asarray(self.variable.data)
        dims = list(self.dims) if dim is None else ([dim] if isinstance(dim, str) else list(dim))
        axes = tuple(self.dims.index(d) for d in dims)
        use = nanfn if (nanfn is not None and np.issubdtype(data.dtype, np.floating)) else fn
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)     # all-NaN slices -> NaN, like xarray
            out = use(data, axis=axes, **kw)
        keep = [d for d in self.dims if d not in dims]
        coords = {k: v for k, v in self._coords.items() if set(v.dims) <= set(keep)}
        return self._new(Variable(keep, out), coords)

    def max(self, dim=None): return self._reduce(np.max, np.nanmax, dim)
    def min(self, dim=None): return self._reduce(np.min, np.nanmin, dim)
    def sum(self, dim=None): return self._reduce(np.sum, np.nansum, dim)
    def mean(self, dim=None): return self._reduce(np.mean, np.nanmean, dim)
    def median(self, dim=None): return self._reduce(np.median, np.nanmedian, dim)
    def std(self, dim=None): return self._reduce(np.std, np.nanstd, dim)
    def all(self, dim=None): return self._reduce(np.all, None, dim)
    def any(self, dim=None): return self._reduce(np.any, None, dim)

    def equals(self, other) -> bool:
        return isinstance(other, DataArray) and self.dims == other.dims and _same_values(self.data, other.data)


def _broadcast_to_dims(arr: DataArray, dims: Sequence[str]) -> np.ndarray:
    data = np.asarray(arr.variable.data)
    have = [d for d in dims if d in arr.dims]
    data = np.transpose(data, [arr.dims.index(d) for d in have])
    shape = [data.shape[have.index(d)] if d in have else 1 for d in dims]
    return data.reshape(shape)


def _same_values(a, b) -> bool:
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind in "fc" and b.dtype.kind in "fc":
        return bool(np.array_equal(a, b, equal_nan=True))
    return bool(np.array_equal(a, b))


_ACCESSORS: Dict[str, type] = {}


def register_dataset_accessor(name: str):
    """xarray.register_dataset_accessor: `ds.<name>` builds (and caches) accessor(ds)."""

    def decorator(cls):
        _ACCESSORS[name] = cls
        return cls

    return decorator


class Dataset:
    def __init__(self, data_vars: Optional[Mapping[str, Any]] = None, coords: Optional[Mapping[str, Any]] = None,
                 attrs: Optional[dict] = None):
        self._variables: Dict[str, Variable] = {}
        self._coord_names: set = set()
        self._stacked: Dict[str, Tuple[str, ...]] = {}     # stacked dim -> the dims it was made of
        self._accessors: Dict[str, Any] = {}
        self.attrs = dict(attrs or {})
        for k, v in (coords or {}).items():
            self._set_coord(k, v)
        for k, v in (data_vars or {}).items():
            self[k] = v

    # -- containers -----------------------------------------------------------------------------
    def _check(self, name: str, var: Variable) -> Variable:
        sizes = self.sizes
        for d, n in var.sizes.items():
            if d in sizes and sizes[d] != n:
                raise ValueError(f"conflicting sizes for dimension {d!r}: length {n} on {name!r} and length "
                                 f"{sizes[d]} on the dataset")
        return var

    def _set_coord(self, name: str, value) -> None:
        if isinstance(value, (tuple, DataArray, Variable)):
            var = _as_variable(value)
        else:
            arr = np.asarray(value)
            var = Variable((name,) if arr.ndim else (), arr)
        self._variables[name] = self._check(name, var)
        self._coord_names.add(name)

    @property
    def sizes(self) -> Dict[str, int]:
        out: Dict[str, int] = {}
        for var in self._variables.values():
            for d, n in var.sizes.items():
                out.setdefault(d, n)
        return out

    dims = sizes
    variables = property(lambda self: dict(self._variables))
    data_vars = property(lambda self: {k: self[k] for k in self._variables if k not in self._coord_names})
    coords = property(lambda self: {k: self[k] for k in self._variables if k in self._coord_names})

    def __contains__(self, name) -> bool:
        return name in self._variables

    def __iter__(self):
        return iter(k for k in self._variables if k not in self._coord_names)

    def keys(self):
        return list(iter(self))

    def _construct(self, name: str) -> DataArray:
        var = self._variables[name]
        out = DataArray(var, name=name)
        out._coords = {k: self._variables[k] for k in self._coord_names
                       if k != name and set(self._variables[k].dims) <= set(var.dims)}
        if name in self._coord_names and var.dims == (name,):
            out._coords[name] = var
        out._stacked = {d: lv for d, lv in self._stacked.items() if d in var.dims}
        return out

    def __getitem__(self, name) -> DataArray:
        if name in self._variables:
            return self._construct(name)
        if name in self.sizes:       # a dimension without a coordinate reads as 0..n-1
            return DataArray(Variable((name,), np.arange(self.sizes[name])), name=name)
        raise KeyError(name)

    def __setitem__(self, name: str, value) -> None:
        var = _as_variable(value)
        if isinstance(value, DataArray):
            for k, cv in value._coords.items():
                if k not in self._variables and k != name:
                    self._variables[k] = self._check(k, cv)
                    self._coord_names.add(k)
        self._variables[name] = self._check(name, var)

    def __delitem__(self, name: str) -> None:
        del self._variables[name]
        self._coord_names.discard(name)

    def __getattr__(self, name: str):
        d = self.__dict__
        if name in d.get("_variables", {}) or name in self.sizes:
            return self[name]
        if name in d.get("attrs", {}):
            return d["attrs"][name]
        if name in _ACCESSORS and "_accessors" in d:
            if name not in d["_accessors"]:
                d["_accessors"][name] = _ACCESSORS[name](self)
            return d["_accessors"][name]
        raise AttributeError(f"'Dataset' object has no attribute {name!r}")

    def __repr__(self):
        lines = [f"<magnify_b200.Dataset {dict(self.sizes)}>"]
        for k, v in self._variables.items():
            kind = "coord" if k in self._coord_names else "data "
            lines.append(f"  {kind} {k}: {v.dims} {v.dtype}")
        return "\n".join(lines)

    def copy(self, deep: bool = False) -> "Dataset":
        new = Dataset(attrs=self.attrs)
        new._variables = {k: (Variable(v.dims, np.array(v.data), v.attrs) if deep else v) for k, v in self._variables.items()}
        new._coord_names = set(self._coord_names)
        new._stacked = dict(self._stacked)
        return new

    # -- xarray methods the components call -----------------------------------------------------
    def assign_coords(self, coords=None, **kw) -> "Dataset":
        new = self.copy()
        for k, v in {**(coords or {}), **kw}.items():
            new._set_coord(k, v)
        return new

    def assign_attrs(self, *args, **kw) -> "Dataset":
        new = self.copy()
        new.attrs.update(*args, **kw)
        return new

    def assign(self, variables=None, **kw) -> "Dataset":
        new = self.copy()
        for k, v in {**(variables or {}), **kw}.items():
            new[k] = v
        return new

    def drop_vars(self, names, errors: str = "raise") -> "Dataset":
        names = [names] if isinstance(names, str) else list(names)
        new = self.copy()
        for n in names:
            if n in new._variables:
                del new[n]
            elif errors == "raise":
                raise ValueError(f"These variables cannot be found in this dataset: [{n!r}]")
        return new

    def rename(self, name_dict=None, **kw) -> "Dataset":
        mapping = {**(name_dict or {}), **kw}
        new = Dataset(attrs=self.attrs)
        for k, v in self._variables.items():
            new._variables[mapping.get(k, k)] = Variable([mapping.get(d, d) for d in v.dims], v.data, v.attrs)
        new._coord_names = {mapping.get(k, k) for k in self._coord_names}
        new._stacked = {mapping.get(d, d): tuple(mapping.get(x, x) for x in lv) for d, lv in self._stacked.items()}
        return new

    def transpose(self, *dims) -> "Dataset":
        new = self.copy()
        named = [d for d in dims if d is not Ellipsis]
        missing = [d for d in named if d not in self.sizes]
        if missing:
            raise ValueError(f"{tuple(dims)} must be a permuted list of {tuple(self.sizes)}, unless `...` is included")
        for k, v in self._variables.items():
            sub = [d for d in dims if d is Ellipsis or d in v.dims]
            if Ellipsis not in sub and len(sub) != len(v.dims):
                sub = sub + [Ellipsis]
            new._variables[k] = v.transpose(*sub) if v.dims else v
        return new

    def _map_arrays(self, fn) -> "Dataset":
        new = Dataset(attrs=self.attrs)
        new._stacked = dict(self._stacked)
        for k in self._variables:
            out = fn(self._construct(k))
            new._variables[k] = out.variable
        new._coord_names = set(self._coord_names)
        return new

    def isel(self, indexers=None, **kw) -> "Dataset":
        indexers = {**(indexers or {}), **kw}
        return self._map_arrays(lambda a: a.isel({d: i for d, i in indexers.items() if d in a.dims}))

    def sel(self, indexers=None, **kw) -> "Dataset":
        indexers = {**(indexers or {}), **kw}
        pos = {d: self[d]._positions(d, v) if d in self._variables else DataArray(
            Variable((d,), np.arange(self.sizes[d])))._positions(d, v) for d, v in indexers.items()}
        return self.isel(pos)

    def squeeze(self, dim=None) -> "Dataset":
        dims = [d for d, n in self.sizes.items() if n == 1] if dim is None else ([dim] if isinstance(dim, str) else list(dim))
        for d in dims:
            if self.sizes[d] != 1:
                raise ValueError(f"cannot select a dimension to squeeze out which has length greater than one: {d}")
        return self.isel({d: 0 for d in dims})

    def stack(self, dimensions=None, create_index: bool = True, **kw) -> "Dataset":
        """Merge dims into one new LAST dimension per variable (row-major), like xarray: variables
        that hold only some of the dims are broadcast first; the merged dims become level
        coordinates of the new dimension (integer ranges when they had no coordinate)."""
        new = self
        for new_dim, old in {**(dimensions or {}), **kw}.items():
            new = new._stack_once(new_dim, tuple(old))
        return new

    def _stack_once(self, new_dim: str, old: Tuple[str, ...]) -> "Dataset":
        sizes = self.sizes
        out = Dataset(attrs=self.attrs)
        out._stacked = dict(self._stacked)
        out._stacked[new_dim] = old
        out._coord_names = set(self._coord_names)
        levels = {}
        for d in old:
            if d in self._variables and self._variables[d].dims == (d,):
                levels[d] = np.asarray(self._variables[d].data)
            else:
                levels[d] = np.arange(sizes[d])
        for k, v in self._variables.items():
            if k in old and v.dims == (k,):
                continue                                   # replaced by the level coordinate below
            if not any(d in v.dims for d in old):
                out._variables[k] = v
                continue
            data = v.data                                  # duck arrays transpose / reshape lazily
            vdims = list(v.dims)
            for d in old:                                  # broadcast to all the stacked dims
                if d not in vdims:
                    data = np.asarray(data)
                    data = np.broadcast_to(data[..., None], data.shape + (sizes[d],))
                    vdims.append(d)
            other = [d for d in vdims if d not in old]
            data = np.transpose(data, [vdims.index(d) for d in other + list(old)])
            data = np.reshape(data, tuple(data.shape[:len(other)]) + (-1,))
            out._variables[k] = Variable(other + [new_dim], data, v.attrs)
        grids = np.meshgrid(*[levels[d] for d in old], indexing="ij")
        for d, g in zip(old, grids):
            out._variables[d] = Variable((new_dim,), g.reshape(-1))
            out._coord_names.add(d)
        return out

    def unstack(self) -> "Dataset":
        """Split every stacked dimension back into its level dims, which become the LAST dims of
        each variable (xarray's order)."""
        new = self
        for dim in list(self._stacked):
            new = new._unstack_once(dim)
        return new

    def _unstack_once(self, dim: str) -> "Dataset":
        old = self._stacked[dim]
        levels = [np.asarray(self._variables[d].data) for d in old]
        uniques = [np.unique(lv) for lv in levels]
        shape = tuple(len(u) for u in uniques)
        pos = [np.searchsorted(u, lv) for u, lv in zip(uniques, levels)]
        flat = np.ravel_multi_index(pos, shape)
        full = len(flat) == int(np.prod(shape)) and np.array_equal(flat, np.arange(len(flat)))
        out = Dataset(attrs=self.attrs)
        out._stacked = {d: lv for d, lv in self._stacked.items() if d != dim}
        out._coord_names = set(self._coord_names)
        for k, v in self._variables.items():
            if k in old:
                continue
            if dim not in v.dims:
                out._variables[k] = v
                continue
            data = v.data
            other = [d for d in v.dims if d != dim]
            data = np.transpose(data, [v.dims.index(d) for d in other] + [v.dims.index(dim)])
            if full:
                data = np.reshape(data, tuple(data.shape[:-1]) + shape)
            else:
                data = np.asarray(data)
                fill = np.full(data.shape[:-1] + (int(np.prod(shape)),), np.nan,
                               dtype=data.dtype if data.dtype.kind == "f" else np.float64)
                fill[..., flat] = data
                data = fill.reshape(data.shape[:-1] + shape)
            out._variables[k] = Variable(other + list(old), data, v.attrs)
        for d, u in zip(old, uniques):
            out._variables[d] = Variable((d,), u)
            out._coord_names.add(d)
        return out

    def equals(self, other) -> bool:
        if not isinstance(other, Dataset) or set(self._variables) != set(other._variables):
            return False
        if self._coord_names != other._coord_names:
            return False
        return all(self._variables[k].dims == other._variables[k].dims and
                   _same_values(self._variables[k].data, other._variables[k].data) for k in self._variables)

    def identical(self, other) -> bool:
        return self.equals(other) and _attrs_equal(self.attrs, other.attrs)


def _attrs_equal(a: dict, b: dict) -> bool:
    if set(a) != set(b):
        return False
    for k in a:
        x, y = a[k], b[k]
        if isinstance(x, np.ndarray) or isinstance(y, np.ndarray):
            if not np.array_equal(np.asarray(x), np.asarray(y)):
                return False
        elif x != y:
            return False
    return True


def concat(objs: Iterable[DataArray], dim: str, **_ignored) -> DataArray:
    """xarray.concat of DataArrays along an EXISTING dimension (what stitch.py:34-35 does)."""
    objs = list(objs)
    first = objs[0]
    if dim not in first.dims:
        raise NotImplementedError("concat along a new dimension is not supported")
    axis = first.dims.index(dim)
    data = np.concatenate([np.asarray(o.transpose(*first.dims).data) for o in objs], axis=axis)
    coords = {k: v for k, v in first._coords.items() if dim not in v.dims}
    return first._new(Variable(first.dims, data, first.attrs), coords)


def zeros_like(other: DataArray, dtype=None) -> DataArray:
    return other._new(Variable(other.dims, np.zeros(other.shape, dtype=dtype or other.dtype)))


def is_dataset(obj) -> bool:
    """True for this module's Dataset and for xarray.Dataset (when xarray is importable)."""
    if isinstance(obj, Dataset):
        return True
    try:
        import xarray as xr
    except Exception:
        return False
    return isinstance(obj, xr.Dataset)
Assay = Dataset      # the name round 1 used for its stand-in

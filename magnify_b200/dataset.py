"""Minimal labelled container used when xarray is not installed.

`Assay` implements the small slice of the `xarray.Dataset` surface that the hot-path components
touch (`"tile" in assay`, `assay.sizes[...]`, `assay["image"] = (dims, array)`,
`assay.assign_coords(...)`, `assay.image`), so that the components of
`magnify_b200.components` run -- and are tested -- in images without xarray/dask (like the
build image).  With xarray present the same components take and return `xarray.Dataset`.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import numpy as np


class Var:
    """A named array with dimension names (the used subset of xarray.DataArray)."""

    def __init__(self, dims: Iterable[str], values: np.ndarray):
        self.dims = tuple(dims)
        self.values = np.asarray(values)
        if self.values.ndim != len(self.dims):
            raise ValueError(f"{self.values.ndim}-d array given {len(self.dims)} dimension names {self.dims}")

    shape = property(lambda self: self.values.shape)
    dtype = property(lambda self: self.values.dtype)
    sizes = property(lambda self: dict(zip(self.dims, self.values.shape)))

    def to_numpy(self) -> np.ndarray:
        return self.values

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __getitem__(self, key):
        return self.values[key]

    def isel(self, **indexers) -> "Var":
        key = tuple(indexers.get(d, slice(None)) for d in self.dims)
        dims = tuple(d for d in self.dims if not np.isscalar(indexers.get(d, slice(None))))
        return Var(dims, self.values[key])

    def __repr__(self):
        return f"Var(dims={self.dims}, shape={self.values.shape}, dtype={self.values.dtype})"


class Assay:
    """dict-of-Var with data variables, coordinates and attrs."""

    def __init__(self, data_vars: Dict[str, Tuple] | None = None, coords: Dict[str, Tuple] | None = None,
                 attrs: dict | None = None):
        self.data_vars: Dict[str, Var] = {}
        self.coords: Dict[str, Var] = {}
        self.attrs = dict(attrs or {})
        for name, (dims, values) in (data_vars or {}).items():
            self[name] = (dims, values)
        for name, (dims, values) in (coords or {}).items():
            self.coords[name] = self._checked(name, Var(dims, values))

    # -- mapping surface ------------------------------------------------------------------------
    def _checked(self, name: str, var: Var) -> Var:
        sizes = self.sizes
        for d, n in var.sizes.items():
            if d in sizes and sizes[d] != n:
                raise ValueError(f"conflicting size for dimension {d!r}: {n} vs {sizes[d]} (variable {name!r})")
        return var

    def __contains__(self, name: str) -> bool:
        return name in self.data_vars or name in self.coords

    def __getitem__(self, name: str) -> Var:
        if name in self.data_vars:
            return self.data_vars[name]
        if name in self.coords:
            return self.coords[name]
        raise KeyError(name)

    def __setitem__(self, name: str, value) -> None:
        var = value if isinstance(value, Var) else Var(*value)
        self.coords.pop(name, None)
        self.data_vars.pop(name, None)
        self.data_vars[name] = self._checked(name, var)

    def __getattr__(self, name: str) -> Var:
        # attribute access like xarray (`assay.tile`); a missing variable is an AttributeError,
        # which is what the reference's Stitcher test expects for a dataset without `tile`.
        try:
            return self.__dict__["data_vars"][name]
        except KeyError:
            pass
        try:
            return self.__dict__["coords"][name]
        except KeyError:
            raise AttributeError(name) from None

    @property
    def sizes(self) -> Dict[str, int]:
        out: Dict[str, int] = {}
        for var in list(self.data_vars.values()) + list(self.coords.values()):
            out.update(var.sizes)
        return out

    def assign_coords(self, **coords) -> "Assay":
        new = self.copy()
        for name, value in coords.items():
            var = value if isinstance(value, Var) else Var(*value)
            new.data_vars.pop(name, None)
            new.coords[name] = new._checked(name, var)
        return new

    def drop_vars(self, names, errors: str = "raise") -> "Assay":
        new = self.copy()
        for n in names:
            if n in new.data_vars:
                del new.data_vars[n]
            elif n in new.coords:
                del new.coords[n]
            elif errors == "raise":
                raise ValueError(f"no variable {n!r}")
        return new

    def copy(self) -> "Assay":
        new = Assay(attrs=self.attrs)
        new.data_vars = dict(self.data_vars)
        new.coords = dict(self.coords)
        return new

    def __repr__(self):
        lines = ["Assay("]
        lines += [f"  data  {k}: {v!r}" for k, v in self.data_vars.items()]
        lines += [f"  coord {k}: {v!r}" for k, v in self.coords.items()]
        return "\n".join(lines + [")"])

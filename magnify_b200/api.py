"""`mg.beads` / `mg.microfluidic_chip` shaped entry points (src/magnify/registry.py:14-203, 452-612)
over the GPU components: read -> standardize_format -> [identify_buttons] -> flatfield_correct ->
stitch -> find_beads / find_buttons -> drop -> restore_format, on the `Assay` stand-in.

`data` is a path pattern (reader.py's `(channel)_(time)_(row)_(col)` language), an `Assay`, a
labelled array with `.dims` / `.values` / `.coords` (an `xarray.DataArray`), or a NumPy array with
`dims` naming its axes the way the reference's DataArrays do ("channel", "time", "row", "col",
"y", "x" -- or the already standardized "tile_row", ... names).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np

from . import reader
from .components import BeadFinder, ButtonFinder, FlatfieldStitcher
from .dataset import Assay, Var

_RENAME = {"x": "tile_x", "y": "tile_y", "row": "tile_row", "col": "tile_col"}      # preprocess.py:16-19
CHIP_SPACING = {"minichip": (375 / 1.61, 400 / 1.61), "pc": (406 / 3.22, 750 / 3.22), "ps": (375 / 3.22, 655 / 3.22)}


def read_pinlist(pinlist, blank=None) -> np.ndarray:
    """identify.py:14-29: a pinlist CSV with columns `Indices` ("(col, row)", 1-based) and
    `MutantID` -> (rows, cols) array of chamber names; names listed in `blank` (default "",
    "blank", "BLANK") and missing names become "" (an empty chamber)."""
    import csv

    blank = ["", "blank", "BLANK"] if blank is None else ([blank] if isinstance(blank, str) else list(blank))
    cells = []
    with open(pinlist, newline="") as f:
        for row in csv.DictReader(f):
            col, r = (int(v) for v in row["Indices"].replace("(", "").replace(")", "").split(","))
            name = row.get("MutantID") or ""
            cells.append((r - 1, col - 1, "" if name in blank else name))
    rows, cols = max(c[0] for c in cells) + 1, max(c[1] for c in cells) + 1
    width = max(1, max(len(c[2]) for c in cells))
    tag = np.zeros((rows, cols), dtype=f"<U{width}")
    for r, col, name in cells:
        tag[r, col] = name
    return tag


def _standardized(data, dims: Optional[Sequence[str]], coords: Optional[dict]):
    """-> list of standardized assays (tile (channel, time, tile_row, tile_col, tile_y, tile_x)), each
    remembering its original tile dims (preprocess.py:11-42)."""
    if isinstance(data, (str, os.PathLike)):
        return [reader.standardize_format(xp) for xp in reader.Reader()(data)]
    if not isinstance(data, Assay) and hasattr(data, "dims") and hasattr(data, "values") and dims is None:
        # a labelled array such as xarray.DataArray (what the reference's callers pass, tests/test_chip.py:40)
        dims = tuple(data.dims)
        labelled = getattr(data, "coords", {})
        coords = dict(coords or {})
        for name in ("channel", "time"):
            if name in dims and name in labelled and name not in coords:
                coords[name] = np.asarray(labelled[name].values if hasattr(labelled[name], "values") else labelled[name])
        data = np.asarray(data.values)
    if isinstance(data, Assay):
        tile = data["tile"]
        arr, names = np.asarray(tile.values), [_RENAME.get(d, d) for d in tile.dims]
        base = data
    else:
        if dims is None:
            raise ValueError("array input needs dims=(...) naming its axes")
        arr, names = np.asarray(data), [_RENAME.get(d, d) for d in dims]
        base = Assay(coords={k: ((k,), np.asarray(v)) for k, v in (coords or {}).items()})
    if arr.ndim != len(names):
        raise ValueError(f"{arr.ndim}-d array given {len(names)} dimension names {tuple(names)}")
    extra = [d for d in names if d not in reader.TILE_ORDER]
    if extra:
        raise NotImplementedError(f"extra dims {extra} (stacked into time by the reference) are not supported")
    full = [d for d in reader.TILE_ORDER if d in names]
    arr = np.transpose(arr, [names.index(d) for d in full])
    for axis, d in enumerate(reader.TILE_ORDER):
        if d not in names:
            arr = np.expand_dims(arr, axis)
    xp = base.copy()
    xp.attrs = dict(base.attrs, __original_tile_dims__=list(names))
    xp["tile"] = (reader.TILE_ORDER, np.ascontiguousarray(arr))
    return [xp]


def _restore(xp: Assay, grid: Optional[tuple], roi_only: bool = False, drop_tiles: bool = True):
    """drop + restore_format (postprocess.py:6-49): un-stack `mark` into (mark_row, mark_col) for
    chips and squeeze the dims standardize_format had added.  roi_only returns the `roi` variable
    alone (postprocess.py:11-12); drop_tiles=False keeps the tile stack (squeezed like the rest)."""
    original = xp.attrs.get("__original_tile_dims__", list(reader.TILE_ORDER))
    added = {d: d not in original for d in reader.TILE_ORDER}                      # postprocess.py:29-33
    out = Assay(attrs={k: v for k, v in xp.attrs.items() if k != "__original_tile_dims__"})
    variables = list(xp.data_vars.items()) + list(xp.coords.items())
    for name, var in variables:
        if name in ("mark_row", "mark_col") or (drop_tiles and (name == "tile" or name.startswith("tile_"))):
            continue
        dims, values = list(var.dims), np.asarray(var.values)
        if grid is not None and dims and dims[0] == "mark":
            values = values.reshape(tuple(grid) + values.shape[1:])
            dims = ["mark_row", "mark_col"] + dims[1:]
        for d in reader.TILE_ORDER:
            if added[d] and d in dims and values.shape[dims.index(d)] == 1:
                values = np.squeeze(values, axis=dims.index(d))
                dims.remove(d)
        target = out.data_vars if name in xp.data_vars else out.coords
        target[name] = Var(tuple(dims), values)
    return out.data_vars["roi"] if roi_only else out


def beads(data, dims=None, coords=None, flatfield=1.0, darkfield=0.0, overlap: int = 102, min_bead_diameter: int = 10,
          max_bead_diameter: int = 50, low_edge_quantile: float = 0.1, high_edge_quantile: float = 0.9,
          num_iter: int = 5000000, min_roundness: float = 0.3, roi_length: Optional[int] = None, search_channel=None,
          roi_only: bool = False, drop_tiles: bool = True, device=None, seed: int = 0):
    """`mg.beads` (registry.py:452-559): one Assay, or a list when the pattern matches several."""
    results = []
    for xp in _standardized(data, dims, coords):
        xp = FlatfieldStitcher(flatfield, darkfield, overlap, device=device)(xp)
        xp = BeadFinder(min_bead_diameter, max_bead_diameter, low_edge_quantile, high_edge_quantile, num_iter,
                        min_roundness, roi_length, search_channel, device=device, seed=seed)(xp)
        results.append(_restore(xp, None, roi_only, drop_tiles))
    return results[0] if len(results) == 1 else results


def microfluidic_chip(data, dims=None, coords=None, shape=(8, 8), pinlist=None, blank=None, tags=None,
                      overlap: int = 102,
                      row_dist: float = 375 / 1.61, col_dist: float = 400 / 1.61, chip_type: Optional[str] = None,
                      min_button_diameter: int = 8, max_button_diameter: int = 30, chamber_diameter: int = 60,
                      top_chamber=None, left_chamber=None, low_edge_quantile: float = 0.1, high_edge_quantile: float = 0.9,
                      num_iter: int = 5000000, min_roundness: float = 0.2, cluster_penalty: float = 50,
                      roi_length: Optional[int] = None, search_timestep=0, search_channel=None, roi_only: bool = False,
                      drop_tiles: bool = True, flatfield=1.0, darkfield=0.0, device=None, seed: int = 0):
    """`mg.microfluidic_chip` (registry.py:14-203): `pinlist` (CSV, identify.py:18-29), `shape` (all
    chambers "default", identify.py:30-32) or `tags`, a ready (rows, cols) array of chamber names
    with "" for blanks.  The reference's chip pipeline has no flat-field step; `flatfield` / `darkfield` are an
    addition (identity by default)."""
    if chip_type is not None:
        if chip_type not in CHIP_SPACING:
            raise ValueError(f"Invalid chip type: {chip_type}. Must be one of ['pc', 'ps', 'minichip']")
        row_dist, col_dist = CHIP_SPACING[chip_type]
    results = []
    for xp in _standardized(data, dims, coords):
        if pinlist is not None:
            tag = read_pinlist(pinlist, blank)                                        # identify.py:18-29
        elif tags is None:
            tag = np.empty((shape[0], shape[1]), dtype="<U200")
            tag.fill("default")                                                       # identify.py:30-32
        else:
            tag = np.asarray(tags)
        t = xp.sizes["time"]
        xp = xp.assign_coords(tag=(("mark_row", "mark_col"), tag),
                              valid=(("mark_row", "mark_col", "time"), np.ones(tag.shape + (t,), dtype=bool)))
        xp = FlatfieldStitcher(flatfield, darkfield, overlap, device=device)(xp)
        xp = ButtonFinder(row_dist, col_dist, min_button_diameter, max_button_diameter, chamber_diameter, top_chamber,
                          left_chamber, low_edge_quantile, high_edge_quantile, num_iter, min_roundness, cluster_penalty,
                          roi_length, False, search_timestep, search_channel, device=device, seed=seed)(xp)
        results.append(_restore(xp, tag.shape, roi_only, drop_tiles))
    return results[0] if len(results) == 1 else results

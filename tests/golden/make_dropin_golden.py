"""Generator of tests/golden/dropin_*.npz (run in the build container, where /root/reference is):

    python tests/golden/make_dropin_golden.py

For every case of tests/dropin_cases.py the reference's OWN pipeline (its registry, builders,
`Pipeline`, and its own components, imported in place by oracle/_refload.load_reference_package)
is run component by component and two states are saved:

  * `in__*`  -- the dataset entering the first hot-path component (after `standardize_format`
                [+ `identify_buttons`]): tile stack, coordinates, tag / valid;
  * `out__*` -- the dataset leaving the last hot-path component (`find_buttons` / `find_beads`),
                before the reference's `drop` / `restore_format`: image, roi, fg, bg, x, y, valid;

  * `valid_after__<filter>` (chip cases) -- `valid` after the reference's own filter_expression,
                filter_nonround(min_roundness=0.6) and filter_leaky applied in that order to the
                `out__` state;

plus the marker centres (and fg radii) the reference arrived at, so that the GPU side can pin
them (`centers=` hook) -- the reference's own centre search is random and unseeded.
tests/test_gpu_dropin.py replays the hot-path components of `magnify_b200.components` with the
real kernels on the `in__` state and requires the `out__` state exactly.
"""
import inspect
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pytest  # noqa: E402

from dropin_cases import bead_input, chip_input, load_mg, pin_circle_finders  # noqa: E402

HOT = {"flatfield_correct", "stitch", "rotate", "find_buttons", "find_beads"}


def pack(prefix, ds, out, skip=()):
    coords = set(ds.coords)
    names = []
    for name, var in ds.variables.items():
        if name in skip:
            continue
        values = np.array(var.values, copy=True)       # a snapshot: later components write into `valid` in place
        if values.dtype.kind in "OU":
            values = values.astype(str)
        out[f"{prefix}__{name}"] = values
        out[f"{prefix}__{name}__dims"] = np.array(list(var.dims), dtype=str)
        names.append(("coord:" if name in coords else "data:") + name)
    out[f"{prefix}__names"] = np.array(names, dtype=str)


def run_case(mg, kind, case):
    data, kwargs = (chip_input if kind == "chip" else bead_input)(case)
    builder = mg.microfluidic_chip_pipe if kind == "chip" else mg.beads_pipe
    accepted = set(inspect.signature(builder).parameters)
    pipe = builder(**{k: v for k, v in kwargs.items() if k in accepted})
    radii = []
    real_circle = mg.utils.circle

    def recording_circle(shape, center, radius, value=1):
        radii.append(int(radius))
        return real_circle(shape, center, radius, value=value)

    beads_seen = []
    real_labels = mg.utils.circle_labels

    def recording_labels(circles, num_rows, num_cols):
        beads_seen.append(np.array(circles))
        return real_labels(circles, num_rows, num_cols)

    mg.utils.circle = recording_circle
    mg.utils.circle_labels = recording_labels
    try:
        (ds,) = list(pipe.reader(data=data))
        out = {}
        entered = False
        for name, comp in pipe.components:
            if name in HOT and not entered:
                pack("in", ds, out)
                entered = True
            if name in ("drop", "restore_format"):
                break
            ds = comp(ds)
        pack("out", ds, out, skip=("tile",))      # the tile stack is the input, already saved
        if kind == "chip":
            # the reference's own consumers of the crops, one after the other on that state
            # (filter.py:11-94); only `valid` changes
            reg = mg.registry.components
            for name, fkw in (("filter_expression", {}), ("filter_nonround", {"min_roundness": 0.6}), ("filter_leaky", {})):
                ds = reg.get(name)(**fkw)(ds)
                out[f"valid_after__{name}"] = np.asarray(ds["valid"].values).copy()
    finally:
        mg.utils.circle = real_circle
        mg.utils.circle_labels = real_labels
    defaults = {k: v.default for k, v in inspect.signature(builder).parameters.items() if v.default is not inspect._empty}
    full = {**defaults, **{k: v for k, v in kwargs.items() if k in accepted}}
    # arrays (flat / dark fields) travel as arrays, the rest as JSON
    arrays = {k: np.asarray(v) for k, v in full.items() if isinstance(v, np.ndarray)}
    for k, v in arrays.items():
        out[f"kwarg__{k}"] = v
    out["kwargs_json"] = np.array(json.dumps({k: v for k, v in full.items() if k not in arrays and k != "pinlist"}))
    if kind == "chip":
        n_search = len(np.atleast_1d(full["search_timestep"]))
        # per button: annulus(outer), annulus(inner), then the foreground disc (find.py:384-397)
        out["fg_radius"] = np.asarray(radii[2::3], dtype=np.int32).reshape(n_search, -1)   # per search timestep, row-major
    else:
        out["beads"] = (beads_seen[0] if beads_seen else np.empty((0, 3))).astype(np.float64)   # rows (row, col, radius)
    return out


def main():
    mg = load_mg()
    if mg is None:
        raise SystemExit("needs /root/reference")
    mp = pytest.MonkeyPatch()
    pin_circle_finders(mp, mg)
    try:
        for kind, cases in (("chip", ["chip_single", "chip_series", "chip_tiles", "chip_blank_float"]),
                            ("beads", ["beads_single", "beads_flatfield_tiles", "beads_none"])):
            for case in cases:
                out = run_case(mg, kind, case)
                path = os.path.join(HERE, f"dropin_{case}.npz")
                np.savez_compressed(path, **out)
                print(f"{path}: {os.path.getsize(path) / 1e3:.0f} kB, {sorted(k for k in out if k.startswith('out__') and '__dims' not in k)}")
    finally:
        mp.undo()


if __name__ == "__main__":
    main()

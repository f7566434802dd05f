"""pytest plugin for running the REFERENCE'S OWN test files (/root/reference/tests/test_*.py, read
in place, never copied) against this package's components:

    python -m pytest -p refsuite_plugin -p no:cacheprovider /root/reference/tests

(tests/test_reference_suite.py does exactly that in a subprocess).  Before collection the plugin

  * imports the reference's unmodified package in place on the stand-ins for xarray / dask /
    catalogue (oracle/_refload.py::load_reference_package), so `import magnify as mg` and
    `import xarray as xr` in the reference's tests resolve to it;
  * calls `magnify_b200.components.install()` into the reference's registry, so `mg.beads`,
    `mg.microfluidic_chip`, `mg.mrbles` build their pipelines from THIS package's
    flatfield_correct / stitch / find_beads / find_buttons / filters, and points
    `magnify.stitch.Stitcher` (which tests/test_stitch.py imports directly) at this package's class;
  * MGB_REFSUITE=reference leaves the reference's own components in place instead (the control run).

This container has no GPU, so the array kernels under the components are the oracle-backed
stand-ins of tests/cpu_ops.py, and this package's circle finder (CUDA) is stood in for by the
reference's own `utils.find_circles` -- the real, randomised one, exactly what the reference's
components use in the control run: what is exercised is the component layer -- dataset schema,
argument handling, errors, copy-forward, grid fit, refinement logic -- under the reference's own
assertions."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for path in (ROOT, HERE):
    if path not in sys.path:
        sys.path.insert(0, path)

import pytest  # noqa: E402

_patch = pytest.MonkeyPatch()
_calls = {}


def _counted(name, fn):
    def wrapper(*args, **kwargs):
        _calls[name] = _calls.get(name, 0) + 1
        return fn(*args, **kwargs)

    return wrapper


def _use_reference_finder(mg):
    """magnify_b200.circles.find_circles / to_uint8 -> the reference's own (utils.py:83-218)."""
    import numpy as np
    import torch

    from magnify_b200 import circles

    def to_uint8(x, batched=False):
        arr = x.detach().cpu().numpy()
        if batched:
            return torch.from_numpy(np.stack([mg.utils.to_uint8(a) for a in arr]) if len(arr) else arr.astype(np.uint8))
        return torch.from_numpy(mg.utils.to_uint8(arr))

    def find_circles(image, low_edge_quantile, high_edge_quantile, grid_length, num_iter, min_radius, max_radius,
                     min_roundness, min_dist, seed=0):
        arr = image.detach().cpu().numpy()

        def one(a):
            return mg.utils.find_circles(a, low_edge_quantile, high_edge_quantile, grid_length, num_iter, min_radius,
                                         max_radius, min_roundness, min_dist, None)

        return [one(a) for a in arr] if arr.ndim == 3 else one(arr)

    _patch.setattr(circles, "to_uint8", to_uint8)
    _patch.setattr(circles, "find_circles", find_circles)


def pytest_configure(config):
    from dropin_cases import load_mg

    mg = load_mg()
    if mg is None:
        raise pytest.UsageError("the reference package is not importable here")
    if os.environ.get("MGB_REFSUITE", "b200") == "b200":
        import cpu_ops
        from magnify_b200 import components

        for name in ("stitch", "flatfield_stitch", "roi_gather_stats", "chip_masks", "bead_masks"):
            _patch.setattr(cpu_ops, name, _counted(name, getattr(cpu_ops, name)))
        cpu_ops.patch_components(_patch)
        _use_reference_finder(mg)
        components.install()
        import magnify.stitch as ref_stitch

        _patch.setattr(ref_stitch, "Stitcher", components.Stitcher)
        assert mg.registry.components.get("find_beads") is components.make_find_beads
        assert mg.registry.components.get("find_buttons") is components.make_find_buttons


def pytest_terminal_summary(terminalreporter):
    if _calls:
        terminalreporter.write_line("magnify_b200 array ops called by the reference's tests: "
                                    + ", ".join(f"{k} x{v}" for k, v in sorted(_calls.items())))


def pytest_sessionfinish(session, exitstatus):
    if os.environ.get("MGB_REFSUITE", "b200") == "b200" and session.testscollected and not _calls.get("roi_gather_stats"):
        session.exitstatus = 1          # the replaced components never ran: the run proves nothing


def pytest_unconfigure(config):
    _patch.undo()

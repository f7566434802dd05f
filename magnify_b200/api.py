"""`mg.beads` / `mg.mrbles` / `mg.microfluidic_chip` with the GPU components: thin callers of the
reference's OWN builders (src/magnify/registry.py:32-693).  `install()` registers the components
of `magnify_b200.components` under the reference's names, then the reference assembles and runs
its pipeline exactly as it always does -- nothing of registry.py / pipeline.py / preprocess.py's
`standardize_format` / postprocess.py is re-implemented here.  Needs `magnify` importable."""
from __future__ import annotations

from . import components


def _magnify():
    try:
        import magnify
    except Exception as e:
        raise ImportError("magnify_b200.api calls the reference's own builders: `magnify` (with xarray, dask, "
                          "catalogue) must be importable.  Without it, chain the components of "
                          "magnify_b200.components by hand (examples/chip_demo.py).") from e
    components.install()
    return magnify


def microfluidic_chip(data, **kwargs):
    """`mg.microfluidic_chip(data, **kwargs)` (registry.py:32-193) on the GPU components."""
    return _magnify().microfluidic_chip(data, **kwargs)


def beads(data, **kwargs):
    """`mg.beads(data, **kwargs)` (registry.py:454-559) on the GPU components."""
    return _magnify().beads(data, **kwargs)


def mrbles(data, spectra, codes, **kwargs):
    """`mg.mrbles(data, spectra, codes, **kwargs)` (registry.py:274-399) on the GPU components."""
    return _magnify().mrbles(data, spectra, codes, **kwargs)


def microfluidic_chip_pipe(**kwargs):
    return _magnify().microfluidic_chip_pipe(**kwargs)


def beads_pipe(**kwargs):
    return _magnify().beads_pipe(**kwargs)


def mrbles_pipe(spectra, codes, **kwargs):
    return _magnify().mrbles_pipe(spectra, codes, **kwargs)

"""magnify_b200 -- B200-native per-marker quantification hot path for FordyceLab/magnify.

Array-level API in `magnify_b200.ops` (torch CUDA tensors in, torch CUDA tensors out, all work
done by hand-written sm_100a kernels behind the C ABI of include/magnify_b200.h), the
reference-facing components in `magnify_b200.components`, the multi-timepoint drivers in
`magnify_b200.pipeline`.  Next to the hot path: `magnify_b200.reader` (TIFF tiles -> pinned
staging), `magnify_b200.circles` / `chipgrid` (the circle finder behind find_beads /
find_buttons) and `magnify_b200.api` (`beads`, `microfluidic_chip` with the reference's
top-level signatures).  No CPU fallback: importing `ops` without the built library raises.
"""
__version__ = "0.2.0"

from . import _lib  # noqa: F401


def build(force: bool = False, verbose: bool = False) -> str:
    from .build import build_library

    return build_library(force=force, verbose=verbose)

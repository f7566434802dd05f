"""The reference's OWN test-suite (/root/reference/tests/test_beads.py, test_chip.py,
test_stitch.py: 39 tests, read in place) run against this package's components.

tests/refsuite_plugin.py imports the reference's unmodified package on the stand-ins for xarray /
dask / catalogue, calls `components.install()` into its registry and points
`magnify.stitch.Stitcher` at this package's class; then pytest runs the reference's test files
as they are.  Every assertion the reference makes about `mg.beads`, `mg.microfluidic_chip` and
`Stitcher` -- dataset type and dims, marker counts, positions and areas within its tolerances,
copy-forward equality across timesteps, dtype preservation, blank chambers, error cases -- must
hold for the replaced components.  The control run (the reference's own components on the same
stand-ins) shows the stand-ins themselves carry the reference's suite.

Build container only (needs /root/reference and no GPU: the array kernels under the components
are the oracle-backed stand-ins of tests/cpu_ops.py, the circle finder is the reference's own)."""
import os
import re
import subprocess
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_TESTS = "/root/reference/tests"


def run_suite(mode):
    env = dict(os.environ, MGB_REFSUITE=mode, PYTHONPATH=os.pathsep.join([HERE, ROOT, os.environ.get("PYTHONPATH", "")]))
    with tempfile.TemporaryDirectory(prefix="mgb_refsuite_") as cwd:     # /root/reference is read-only: no cache, no rootdir there
        out = subprocess.run([sys.executable, "-m", "pytest", "-p", "refsuite_plugin", "-p", "no:cacheprovider",
                              "--rootdir", cwd, "-q", REF_TESTS], cwd=cwd, env=env, capture_output=True, text=True,
                             timeout=1500)
    tail = "\n".join(out.stdout.strip().splitlines()[-25:])
    summary = re.search(r"(\d+) passed", out.stdout)
    return out.returncode, int(summary.group(1)) if summary else 0, tail


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="/root/reference not available (GPU box)")
@pytest.mark.parametrize("mode", ["b200", "reference"])
def test_reference_suite_passes(mode):
    rc, passed, tail = run_suite(mode)
    assert rc == 0, tail
    assert passed >= 39, tail
    if mode == "b200":        # the plugin reports which of this package's array ops the reference's tests reached
        assert "roi_gather_stats x" in tail and "stitch x" in tail and "chip_masks x" in tail and "bead_masks x" in tail, tail

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # A GPU test on a box without a GPU is a hard error, not a skip: the CUDA path is the product.
    pass


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    return load


def pattern_image(c, t, h, w, salt=0):
    """Same deterministic uint16 stack as tests/golden/make_golden.py::pattern_image."""
    cc, tt, yy, xx = np.meshgrid(np.arange(c), np.arange(t), np.arange(h), np.arange(w), indexing="ij")
    v = cc * 7919 + tt * 10473 + yy * 131 + xx * 31 + (yy * xx) % 977 + salt * 2221
    return (v % 65536).astype(np.uint16)


@pytest.fixture(scope="session")
def make_pattern_image():
    return pattern_image


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device")
    from magnify_b200 import _lib

    _lib.load()  # fail loudly if the CUDA library was not built
    return torch.device("cuda:0")

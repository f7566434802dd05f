"""The ctypes stub printed in INTEGRATION.md section 1, executed as written (raw CDLL, no
magnify_b200._lib): what a magnify maintainer would paste to call the C ABI directly."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import stitch as o_st

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "magnify_b200", "libmagnify_b200.so")


def test_integration_stub_stitch(cuda_device):
    lib = ctypes.CDLL(LIB)
    lib.mgb_stitch.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 8 + \
                              [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
    lib.mgb_stitch.restype = ctypes.c_int
    lib.mgb_error_string.restype = ctypes.c_char_p

    def stitch(tiles: torch.Tensor, overlap: int) -> torch.Tensor:
        c, t, r, cc, h, w = tiles.shape
        image = torch.empty((c, t, r * (h - overlap), cc * (w - overlap)), dtype=tiles.dtype, device=tiles.device)
        rc = lib.mgb_stitch(tiles.data_ptr(), image.data_ptr(), 0, c, t, r, cc, h, w, overlap,
                            tiles.element_size(), None, torch.cuda.current_stream().cuda_stream)
        if rc:
            raise RuntimeError(lib.mgb_error_string(rc).decode())
        return image

    rng = np.random.default_rng(0)
    for dtype, shape, overlap in ((np.uint16, (2, 2, 2, 3, 40, 48), 6), (np.float64, (1, 1, 2, 2, 25, 25), 8),
                                  (np.uint8, (1, 2, 1, 2, 20, 20), 0)):
        tiles = (rng.random(shape) * 200).astype(dtype)
        got = stitch(torch.from_numpy(tiles).to(cuda_device), overlap).cpu().numpy()
        np.testing.assert_array_equal(got, o_st.stitch(tiles, overlap))
    with pytest.raises(RuntimeError):
        stitch(torch.zeros((1, 1, 2, 2, 50, 50), dtype=torch.uint16, device=cuda_device), 100)   # stitch.py:16-20

"""oracle/tiff.py (struct-level TIFF / BigTIFF decoder) pinned against OpenCV's libtiff decoder
and against two independent producers (tests/tiffgen.py and cv2.imwrite)."""
import os

import numpy as np
import pytest

from oracle import tiff as ot
from tiffgen import write_tiff

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("byteorder", ["<", ">"])
def test_struct_decoder_matches_libtiff(tmp_path, big, byteorder):
    rng = np.random.default_rng(0)
    for dtype in (np.uint8, np.uint16, np.float32):
        pages = [(rng.random((37, 53)) * 250).astype(dtype) for _ in range(3)]
        for rps, scatter in ((None, False), (5, False), (16, True)):
            path = write_tiff(os.path.join(tmp_path, "a.tif"), pages, big=big, byteorder=byteorder, rows_per_strip=rps,
                              scatter=scatter, pad_strips=3 if scatter else 0, description="some text")
            ok, decoded = cv2.imreadmulti(path, flags=cv2.IMREAD_UNCHANGED)
            assert ok and len(decoded) == 3
            for k in range(3):
                got = ot.read_page(path, k)
                assert got.dtype == np.dtype(dtype)
                np.testing.assert_array_equal(got, pages[k])
                np.testing.assert_array_equal(decoded[k], pages[k])
            assert ot.description(path).rstrip(b"\0") == b"some text"


def test_struct_decoder_reads_libtiff_output(tmp_path):
    rng = np.random.default_rng(1)
    pages = [rng.integers(0, 65535, (40, 64), dtype=np.uint16) for _ in range(3)]
    single, multi = os.path.join(tmp_path, "s.tif"), os.path.join(tmp_path, "m.tif")
    assert cv2.imwrite(single, pages[0], [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    assert cv2.imwritemulti(multi, pages, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    np.testing.assert_array_equal(ot.read_page(single, 0), pages[0])
    for k in range(3):
        np.testing.assert_array_equal(ot.read_page(multi, k), pages[k])
    assert cv2.imwrite(single, pages[0])                         # libtiff default: LZW
    with pytest.raises(ValueError):
        ot.read_page(single, 0)

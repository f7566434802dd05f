"""Pure-Python TIFF / BigTIFF page decoder (uncompressed pages) -- TEST INFRASTRUCTURE.

The reference reads tile pages with `tifffile.TiffFile(f).pages[i].asarray()`
(src/magnify/reader.py:265-279).  tifffile (pinned 2026.1.28 in the reference's uv.lock) is not
in /root/reference and not installable here, so this restates the published TIFF 6.0 baseline
layout + the BigTIFF extension with `struct`, field by field, independently of the C++ reader
in magnify_b200/csrc/tiff_pages.cpp.  It is pinned against OpenCV's libtiff decoder
(`cv2.imreadmulti(..., IMREAD_UNCHANGED)`, cv2 4.13 in-image) in tests/test_oracle_tiff.py:
for an uncompressed page tifffile, libtiff and this file must all return the stored samples.
"""
from __future__ import annotations

import struct

import numpy as np

TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
            13: "I", 16: "Q", 17: "q", 18: "Q"}


def read_ifds(path: str):
    """List of {tag: values-tuple} for the main IFD chain, plus (byteorder, bigtiff)."""
    data = open(path, "rb").read()
    bo = {b"II": "<", b"MM": ">"}[data[:2]]
    magic = struct.unpack(bo + "H", data[2:4])[0]
    if magic == 42:
        big, off = False, struct.unpack(bo + "I", data[4:8])[0]
    elif magic == 43:
        assert struct.unpack(bo + "HH", data[4:8]) == (8, 0)
        big, off = True, struct.unpack(bo + "Q", data[8:16])[0]
    else:
        raise ValueError("not a TIFF file")
    ifds = []
    while off:
        n = struct.unpack(bo + ("Q" if big else "H"), data[off : off + (8 if big else 2)])[0]
        pos = off + (8 if big else 2)
        tags = {}
        for _ in range(n):
            if big:
                tag, typ, count = struct.unpack(bo + "HHQ", data[pos : pos + 12])
                raw, size = data[pos + 12 : pos + 20], 20
            else:
                tag, typ, count = struct.unpack(bo + "HHI", data[pos : pos + 8])
                raw, size = data[pos + 8 : pos + 12], 12
            pos += size
            if typ not in TYPE_FMT:
                continue
            fmt = TYPE_FMT[typ]
            nbytes = struct.calcsize("=" + fmt) * count
            if nbytes > len(raw):
                where = struct.unpack(bo + ("Q" if big else "I"), raw)[0]
                raw = data[where : where + nbytes]
            if typ == 2:
                tags[tag] = raw[:nbytes]
            else:
                tags[tag] = struct.unpack(bo + fmt * count, raw[:nbytes])
        ifds.append(tags)
        off = struct.unpack(bo + ("Q" if big else "I"), data[pos : pos + (8 if big else 4)])[0]
    return ifds, bo, big, data


def page_dtype(tags, bo: str) -> np.dtype:
    bits = tags.get(258, (1,))[0]
    fmt = tags.get(339, (1,))[0]
    kind = {1: "u", 2: "i", 3: "f"}[fmt]
    return np.dtype(f"{bo}{kind}{bits // 8}")


def read_page(path: str, page: int) -> np.ndarray:
    """Samples of page `page` as a (height, width[, samples]) array in native byte order."""
    ifds, bo, _, data = read_ifds(path)
    tags = ifds[page]
    if tags.get(259, (1,))[0] != 1:
        raise ValueError("compressed page")
    width, height = tags[256][0], tags[257][0]
    spp = tags.get(277, (1,))[0]
    rps = min(tags.get(278, (height,))[0], height)
    dt = page_dtype(tags, bo)
    row_bytes = width * spp * dt.itemsize
    out = bytearray()
    for s, (off, cnt) in enumerate(zip(tags[273], tags[279])):
        rows = min(rps, height - s * rps)
        assert cnt >= rows * row_bytes
        out += data[off : off + rows * row_bytes]
    arr = np.frombuffer(bytes(out), dtype=dt).astype(dt.newbyteorder("="))
    return arr.reshape((height, width) if spp == 1 else (height, width, spp))


def description(path: str, page: int = 0) -> bytes:
    ifds, *_ = read_ifds(path)
    return ifds[page].get(270, b"")

"""Minimal TIFF / BigTIFF writer for the reader tests (strips or tiles, uncompressed or Deflate,
optional horizontal differencing and ImageDescription).  Independent of both decoders under test; cv2.imwrite is the second producer."""
from __future__ import annotations

import struct
import zlib

import numpy as np


def write_tiff(path, pages, *, big=False, byteorder="<", rows_per_strip=None, description=None, scatter=False,
               pad_strips=0, header_extra=b"", tile=None, deflate=False, predictor=False):
    """pages: list of 2-D (or H x W x S) arrays of one dtype.  scatter=True stores the strips of a
    page in reverse file order (non-contiguous), pad_strips adds slack bytes to every StripByteCount
    region.  header_extra is inserted right after the 8/16-byte header (e.g. Micro-Manager blocks).
    tile=(th, tw) writes tiled pages (edge tiles zero padded) instead of strips; deflate=True
    compresses every strip / tile with zlib (Compression 8); predictor=True applies horizontal
    differencing (Predictor 2) first."""
    bo = byteorder
    off_fmt, off_size = ("Q", 8) if big else ("I", 4)
    buf = bytearray()
    buf += b"II" if bo == "<" else b"MM"
    if big:
        buf += struct.pack(bo + "HHH", 43, 8, 0) + b"\0" * 8
    else:
        buf += struct.pack(bo + "H", 42) + b"\0" * 4
    first_ifd_pos = len(buf) - off_size
    buf += header_extra
    prev_next_pos = first_ifd_pos
    for k, page in enumerate(pages):
        page = np.asarray(page)
        h, w = page.shape[:2]
        spp = 1 if page.ndim == 2 else page.shape[2]
        raw = page.astype(page.dtype.newbyteorder(bo)).tobytes()
        row_bytes = w * spp * page.dtype.itemsize
        rps = h if rows_per_strip is None else rows_per_strip
        nstrips = (h + rps - 1) // rps
        def encode(block):
            """block: 2-D/3-D array in the file's byte order -> stored bytes."""
            if predictor:
                flat = block.reshape(block.shape[0], -1).astype(block.dtype.newbyteorder("="))
                diff = flat.copy()
                diff[:, spp:] = flat[:, spp:] - flat[:, :-spp]
                block = diff.astype(block.dtype)
            data = block.tobytes()
            return zlib.compress(data) if deflate else data

        stored = page.astype(page.dtype.newbyteorder(bo))
        if tile is not None:
            th, tw = tile
            chunks = []
            for ty in range(0, h, th):
                for tx in range(0, w, tw):
                    block = np.zeros((th, tw) + stored.shape[2:], dtype=stored.dtype)
                    part = stored[ty:ty + th, tx:tx + tw]
                    block[:part.shape[0], :part.shape[1]] = part
                    chunks.append(encode(block))
            nstrips = len(chunks)
        else:
            chunks = [encode(stored[s * rps:min((s + 1) * rps, h)]) for s in range(nstrips)]
        offsets = [0] * nstrips
        order = range(nstrips - 1, -1, -1) if scatter else range(nstrips)
        for s in order:
            if len(buf) % 2:
                buf += b"\0"
            offsets[s] = len(buf)
            buf += chunks[s] + b"\xee" * pad_strips
        counts = [len(c) + pad_strips for c in chunks]
        fmt = {"u": 1, "i": 2, "f": 3}[page.dtype.kind]
        entries = [(256, 4, [w]), (257, 4, [h]), (258, 3, [page.dtype.itemsize * 8] * spp), (259, 3, [8 if deflate else 1]),
                   (262, 3, [1 if spp == 1 else 2])]
        if predictor:
            entries.append((317, 3, [2]))
        if description is not None and (k == 0 or isinstance(description, list)):
            text = description[k] if isinstance(description, list) else description
            entries.append((270, 2, text.encode() + b"\0"))
        if tile is not None:
            entries += [(277, 3, [spp]), (322, 4, [tile[1]]), (323, 4, [tile[0]]), (324, 16 if big else 4, offsets),
                        (325, 16 if big else 4, counts), (339, 3, [fmt] * spp)]
        else:
            entries += [(273, 16 if big else 4, offsets), (277, 3, [spp]), (278, 4, [rps]),
                        (279, 16 if big else 4, counts), (339, 3, [fmt] * spp)]
        entries.sort(key=lambda e: e[0])
        # out-of-line values first
        packed = []
        for tag, typ, vals in entries:
            if typ == 2:
                payload, count = bytes(vals), len(vals)
            else:
                code = {3: "H", 4: "I", 16: "Q"}[typ]
                payload, count = struct.pack(bo + code * len(vals), *vals), len(vals)
            if len(payload) > off_size:
                if len(buf) % 2:
                    buf += b"\0"
                where = len(buf)
                buf += payload
                payload = struct.pack(bo + off_fmt, where)
            packed.append((tag, typ, count, payload.ljust(off_size, b"\0")))
        if len(buf) % 2:
            buf += b"\0"
        ifd_pos = len(buf)
        buf[prev_next_pos : prev_next_pos + off_size] = struct.pack(bo + off_fmt, ifd_pos)
        buf += struct.pack(bo + ("Q" if big else "H"), len(packed))
        for tag, typ, count, payload in packed:
            buf += struct.pack(bo + "HH" + ("Q" if big else "I"), tag, typ, count) + payload
        prev_next_pos = len(buf)
        buf += b"\0" * off_size
    with open(path, "wb") as f:
        f.write(bytes(buf))
    return path


def ome_xml(size_x, size_y, size_c=1, size_t=1, size_z=1, order="XYCZT", dtype="uint16", delta_t_ms=None,
            channel_names=None):
    planes = ""
    if delta_t_ms is not None:
        idx = 0
        for t in range(size_t):
            for c in range(size_c):
                planes += f'<Plane TheC="{c}" TheT="{t}" TheZ="0" DeltaT="{delta_t_ms[idx]}" DeltaTUnit="ms"/>'
                idx += 1
    chans = "".join(f'<Channel ID="Channel:0:{i}" Name="{n}"/>' for i, n in enumerate(channel_names or []))
    return ('<?xml version="1.0" encoding="UTF-8"?><OME xmlns="http://www.openmicroscopy.org/Schemas/OME/2016-06">'
            f'<Image ID="Image:0"><Pixels ID="Pixels:0" DimensionOrder="{order}" Type="{dtype}" SizeX="{size_x}" '
            f'SizeY="{size_y}" SizeC="{size_c}" SizeT="{size_t}" SizeZ="{size_z}">{chans}<TiffData/>{planes}'
            '</Pixels></Image></OME>')

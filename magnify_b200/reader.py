"""TIFF tile reader feeding the staging ring.

Mirrors the reference's `read` component for TIFF inputs (src/magnify/reader.py): the path
pattern language of `extract_paths` (:80-160), the assembly of path dimensions and in-file
dimensions into the canonical tile stack of `read_tiffs` (:163-326), and -- instead of one
`tifffile` call per dask chunk (:265-279) -- native page reads that land directly in pinned
staging buffers (`include/magnify_b200.h`, `mgb_tiff_*`; `csrc/tiff_pages.cpp`).

What is and is not reproduced:
  * pixels: classic TIFF / BigTIFF pages, either byte order, strips or tiles, uncompressed (the
    direct path: pread into the pinned slot) or LZW / Deflate / PackBits with optional horizontal
    differencing (decoded on the reading threads); bit-exact.  JPEG, the floating-point predictor
    and planar multi-sample pages raise;
  * in-file axes: single-page files ("YX") and OME-TIFF series described by the OME-XML `Pixels`
    element of the first page (DimensionOrder / SizeC / SizeT / SizeZ, length-1 axes squeezed, the
    way tifffile presents `series[0]`); a multi-page file without OME-XML has the tifffile axis
    "I", which the reference cannot map either (KeyError at reader.py:208) -- same error here;
  * Micro-Manager `Summary` metadata (StartTime, ChNames) from the file header block and OME
    `Plane@DeltaT` times, as reader.py:210-246 uses them;
  * zarr directories (reader.py:56-65) are storage, out of scope: NotImplementedError.
"""
from __future__ import annotations

import collections
import ctypes
import datetime
import fnmatch
import glob
import json
import os
import re
import struct
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .dataset import DataArray, Dataset

TILE_ORDER = ("channel", "time", "tile_row", "tile_col", "tile_y", "tile_x")
LETTER_TO_DIM = {"C": "channel", "T": "time", "Z": "depth", "Y": "tile_y", "X": "tile_x", "R": "tile_pos"}


# ---------------------------------------------------------------------------------------------
# native file handle
# ---------------------------------------------------------------------------------------------
class PageInfo(collections.namedtuple("PageInfo", "width height bits samples sample_format compression nbytes "
                                                   "status description_bytes strips bigtiff big_endian")):
    @property
    def dtype(self) -> np.dtype:
        kind = {1: "u", 2: "i", 3: "f"}.get(self.sample_format, "u")
        return np.dtype(f"{kind}{self.bits // 8}")

    @property
    def shape(self) -> Tuple[int, ...]:
        return (self.height, self.width) if self.samples == 1 else (self.height, self.width, self.samples)


class TiffFile:
    """An open TIFF file: its main IFD chain parsed once by the native reader."""

    def __init__(self, path):
        self.path = os.fspath(path)
        handle = ctypes.c_void_p()
        _lib.call("mgb_tiff_open", self.path.encode(), ctypes.byref(handle))
        self._handle = handle
        n = ctypes.c_int64()
        _lib.call("mgb_tiff_page_count", self._handle, ctypes.byref(n))
        self.num_pages = int(n.value)

    def close(self) -> None:
        if self._handle is not None:
            _lib.call("mgb_tiff_close", self._handle)
            self._handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def page_info(self, page: int = 0) -> PageInfo:
        info = (ctypes.c_int64 * 12)()
        _lib.call("mgb_tiff_page_info", self._handle, int(page), info)
        return PageInfo(*[int(v) for v in info])

    def description(self, page: int = 0) -> bytes:
        n = self.page_info(page).description_bytes
        buf = ctypes.create_string_buffer(max(n, 1))
        _lib.call("mgb_tiff_description", self._handle, int(page), buf, n)
        return buf.raw[:n].rstrip(b"\0")

    def read_pages(self, pages: Sequence[int], out: Optional[np.ndarray] = None, threads: int = 4) -> np.ndarray:
        """Pages `pages` into out[(len(pages),) + page shape] (allocated when None)."""
        pages = [int(p) for p in pages]
        info = self.page_info(pages[0]) if pages else self.page_info(0)
        if out is None:
            out = np.empty((len(pages),) + info.shape, dtype=info.dtype)
        _check_destination(out, len(pages), info)
        idx = (ctypes.c_int64 * len(pages))(*pages)
        _lib.call("mgb_tiff_read_pages", self._handle, idx, len(pages), ctypes.c_void_p(out.ctypes.data),
                  info.nbytes, int(threads))
        return out

    def asarray(self, page: int = 0) -> np.ndarray:
        return self.read_pages([page], threads=1)[0]


def _check_destination(out: np.ndarray, n: int, info: PageInfo) -> None:
    if not out.flags.c_contiguous or not out.flags.writeable:
        raise ValueError("destination must be a writable C-contiguous array")
    if out.dtype != info.dtype or out.size != n * info.nbytes // info.dtype.itemsize:
        raise ValueError(f"destination {out.shape} {out.dtype} does not hold {n} pages of {info.shape} {info.dtype}")


def read_files(paths: Sequence[str], out: np.ndarray, page: int = 0, threads: int = 8) -> np.ndarray:
    """Page `page` of every file into out[i] -- one file per tile, the usual acquisition layout."""
    if len(paths) == 0:
        return out
    if not out.flags.c_contiguous or not out.flags.writeable:
        raise ValueError("destination must be a writable C-contiguous array")
    if out.shape[0] != len(paths) or out.ndim < 3:
        raise ValueError(f"destination {out.shape} does not hold {len(paths)} pages")
    height, width = out.shape[-2:]
    arr = (ctypes.c_char_p * len(paths))(*[os.fspath(p).encode() for p in paths])
    stride = out[0].nbytes
    _lib.call("mgb_tiff_read_files", arr, len(paths), int(page), int(width), int(height), out.dtype.itemsize * 8,
              ctypes.c_void_p(out.ctypes.data), stride, int(threads))
    return out


def write_tiff(path, pages: np.ndarray, description: Optional[str] = None, bigtiff: Optional[bool] = None,
               threads: int = 8) -> str:
    """Save (H, W) or (N, H, W) pages as an uncompressed TIFF / BigTIFF (one strip per page, parallel
    writes) -- e.g. a stitched image or a stack of crops taken from the pinned result buffers.  The
    reference has no TIFF output (it caches to zarr); this is the inverse of the page reader."""
    arr = np.ascontiguousarray(pages)
    if arr.ndim == 2:
        arr = arr[None]
    if arr.ndim != 3:
        raise ValueError(f"pages must be (H, W) or (N, H, W), got {arr.shape}")
    kind = {"u": 1, "i": 2, "f": 3}.get(arr.dtype.kind)
    if kind is None or arr.dtype.itemsize not in (1, 2, 4, 8) or arr.dtype.byteorder == ">":
        raise TypeError(f"unsupported dtype {arr.dtype}")
    _lib.call("mgb_tiff_write", os.fspath(path).encode(), ctypes.c_void_p(arr.ctypes.data), arr.shape[0], arr.shape[1],
              arr.shape[2], arr.dtype.itemsize * 8, kind, -1 if bigtiff is None else int(bool(bigtiff)),
              None if description is None else description.encode(), int(threads))
    return os.fspath(path)


# ---------------------------------------------------------------------------------------------
# path patterns  (reader.py:80-160)
# ---------------------------------------------------------------------------------------------
def _format_time(text, fmt):
    return datetime.datetime.strptime(text, fmt if fmt else "%Y%m%d-%H%M%S")


FORMATTERS: Dict[str, Callable] = {
    "": lambda x, y: x,
    "str": lambda x, y: x,
    "time": _format_time,
    "int": lambda x, y: int(x),
    "float": lambda x, y: float(x),
}


def extract_paths(pattern, **kwargs):
    """Expand a search pattern with named groups into `{index tuple: absolute path}` plus the
    per-key metadata maps -- same contract as reader.py:80-160.

    `(key)` / `(key|format)` captures one path-component substring as the index of `key`;
    `(name_key|formatter|format)` captures extra metadata `name` attached to `key`.  Keys that do
    not occur in the pattern give `None` in the index tuple.  Globbing is recursive, matching is
    case-insensitive, two files with the same index raise ValueError."""
    keys = {k: (f if callable(f) else FORMATTERS[f]) for k, f in kwargs.items()}
    all_keys = list(keys)
    pattern = os.path.expanduser(os.fspath(pattern))

    # Claim the parenthesised groups key by key, in key order, first the ones that start with the
    # key, then the ones that contain "_key" -- the precedence of the reference's substitutions.
    groups = [(m.start(), m.end(), m.group(1)) for m in re.finditer(r"\(([^()]*)\)", pattern)]
    claimed: Dict[int, Optional[str]] = {}      # group number -> regex group name (None: anonymous)
    meta: Dict[str, Dict[str, Callable]] = collections.defaultdict(dict)
    active: Dict[str, Callable] = {}
    for key, formatter in keys.items():
        for g, (_, _, text) in enumerate(groups):
            if g not in claimed and text.startswith(key):
                claimed[g] = key
        for g, (_, _, text) in enumerate(groups):
            if g not in claimed and ("_" + key) in text:
                claimed[g] = text[: text.index("_" + key)]
        spec = re.search(rf"\({re.escape(key)}(?:\s*\|\s*(.*?))?\)", pattern)
        if spec:
            active[key] = (lambda x, y=spec.group(1), f=formatter: f(x, y))
        for name, formatter_name, fmt in re.findall(
                rf"\(([^\(]*?)_{re.escape(key)}(?:\s*\|\s*(.*?))?(?:\s*\|\s*(.*?))?\)", pattern):
            meta[key][name] = (lambda x, y=fmt, f=FORMATTERS[formatter_name]: f(x, y))

    # Build the glob (groups -> *) and the regex (groups -> named wildcards inside one component).
    glob_parts, rx_parts, pos, used = [], [], 0, set()
    for g, (start, end, _) in enumerate(groups):
        if g not in claimed:
            continue
        glob_parts.append(pattern[pos:start] + "*")
        name = claimed[g]
        token = f"MGBxGROUPx{g}x"
        rx_parts.append(pattern[pos:start] + token)
        used.add((token, name))
        pos = end
    glob_path = "".join(glob_parts) + pattern[pos:]
    regex = fnmatch.translate("".join(rx_parts) + pattern[pos:])
    seen_names = set()
    for token, name in sorted(used):
        if name and name.isidentifier() and name not in seen_names:
            regex = regex.replace(token, rf"(?P<{name}>[^/\\]*?)")
            seen_names.add(name)
        else:
            regex = regex.replace(token, r"[^/\\]*?")
    matcher = re.compile(regex, re.IGNORECASE)

    path_dict: Dict[tuple, str] = {}
    meta_dict: Dict[tuple, dict] = collections.defaultdict(dict)
    for path in glob.glob(glob_path, recursive=True):
        match = matcher.fullmatch(path)
        idxs = []
        for key in all_keys:
            if key not in active:
                idxs.append(None)
                continue
            idx = active[key](match.group(key))
            idxs.append(idx)
            for name, formatter in meta[key].items():
                meta_dict[name, key][idx] = formatter(match.group(name))
        idxs = tuple(idxs)
        if idxs in path_dict:
            raise ValueError(f"{path} and {path_dict[idxs]} map to the same index.")
        path_dict[idxs] = os.path.abspath(path)
    return path_dict, meta_dict


# ---------------------------------------------------------------------------------------------
# in-file metadata
# ---------------------------------------------------------------------------------------------
def micromanager_summary(path) -> Optional[dict]:
    """The Micro-Manager `Summary` JSON of a file, or None.  Micro-Manager writes four header
    blocks after the 8-byte TIFF header; the fourth is (magic 2355492, length) at byte 32
    followed by the summary JSON (the block tifffile exposes as
    `micromanager_metadata["Summary"]`, reader.py:210-246)."""
    with open(path, "rb") as f:
        head = f.read(40)
        if len(head) < 40 or head[:2] not in (b"II", b"MM"):
            return None
        bo = "<" if head[:2] == b"II" else ">"
        magic, length = struct.unpack(bo + "II", head[32:40])
        if magic != 2355492 or length <= 0 or length > (64 << 20):
            return None
        try:
            return json.loads(f.read(length).rstrip(b"\0").decode("utf-8", "replace"))
        except ValueError:
            return None


class SeriesLayout(collections.namedtuple("SeriesLayout", "axes shape delta_t_ms")):
    """axes/shape of the first series the way tifffile reports them (squeezed), plus the DeltaT
    (ms) of every OME plane in file order (empty when absent)."""


def series_layout(tif: TiffFile) -> SeriesLayout:
    info = tif.page_info(0)
    text = tif.description(0)
    page_axes = "YX" if info.samples == 1 else "YXS"
    if text.lstrip()[:5] == b"<?xml" and b"<OME" in text[:4096]:
        xml = text.decode("utf-8", "replace")
        pixels = re.search(r"<Pixels\b([^>]*)>", xml)
        if pixels:
            attrs = dict(re.findall(r'(\w+)="([^"]*)"', pixels.group(1)))
            order = attrs.get("DimensionOrder", "XYCZT")
            sizes = {a: int(attrs.get("Size" + a, 1)) for a in "XYCZT"}
            axes, shape = "", ()
            for a in reversed(order):                    # slowest-varying first
                if a in "XY" or sizes[a] > 1:            # tifffile squeezes length-1 axes
                    axes += a
                    shape += (sizes[a],)
            image = re.search(r"<Image\b.*?</Image>", xml, re.S)
            planes = re.findall(r"<Plane\b([^>]*)/?>", image.group(0) if image else xml)
            delta = []
            for p in planes:
                pa = dict(re.findall(r'(\w+)="([^"]*)"', p))
                if "DeltaT" in pa:
                    if pa.get("DeltaTUnit", "ms") != "ms":
                        raise AssertionError("OME Plane DeltaTUnit must be ms")   # reader.py:226
                    delta.append(float(pa["DeltaT"]))
            return SeriesLayout(axes, shape, tuple(delta))
    if tif.num_pages == 1:
        return SeriesLayout(page_axes, info.shape, ())
    return SeriesLayout("I" + page_axes, (tif.num_pages,) + info.shape, ())


# ---------------------------------------------------------------------------------------------
# lazy tile stack
# ---------------------------------------------------------------------------------------------
class TiffTiles:
    """Lazy (dims_in_path + dims_in_file) tile array: one TIFF page per innermost (tile_y, tile_x)
    plane, like the dask array of reader.py:265-292, but read natively.

    `read(index, out)` fills `out` (any C-contiguous host array, normally pinned) with the planes
    selected by `index` over the leading (non-page) dims; `np.asarray(tiles)` reads everything."""

    def __init__(self, filenames: List[str], outer_shape: Tuple[int, ...], inner_shape: Tuple[int, ...],
                 dims: Tuple[str, ...], dtype: np.dtype, threads: int = 8):
        self.filenames = list(filenames)
        self.outer_shape, self.inner_shape = tuple(outer_shape), tuple(inner_shape)
        self.dims = tuple(dims)
        self.dtype = np.dtype(dtype)
        self.shape = self.outer_shape + self.inner_shape
        self.threads = threads
        self._perm = tuple(range(len(self.shape)))     # storage axis of every presented axis

    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: int(np.prod(self.shape)))
    nbytes = property(lambda self: self.size * self.dtype.itemsize)

    def transpose(self, dims: Sequence[str]) -> "TiffTiles":
        """Reorder the leading dims; names not present become new length-1 dims (the
        `expand_dims` of standardize_format, preprocess.py:35-38).  Page dims stay last."""
        dims = tuple(dims)
        if dims[-2:] != self.dims[-2:] or set(self.dims) - set(dims):
            raise ValueError(f"cannot present dims {self.dims} as {dims}: the two page dims must stay last "
                             "and no dim may be dropped")
        new = TiffTiles(self.filenames, self.outer_shape, self.inner_shape, dims, self.dtype, self.threads)
        new.shape = tuple(self.shape[self.dims.index(d)] if d in self.dims else 1 for d in dims)
        new._perm = tuple(self._perm[self.dims.index(d)] if d in self.dims else None for d in dims)
        return new

    def _planes(self, lead_index: Tuple[int, ...]) -> Tuple[int, int]:
        """(file number, page number) of the plane at presented leading index `lead_index`."""
        storage = [0] * (len(self.outer_shape) + len(self.inner_shape) - 2)
        for axis, i in zip(self._perm[:-2], lead_index):
            if axis is not None:
                storage[axis] = i
        n_outer = len(self.outer_shape)
        file_idx = int(np.ravel_multi_index(storage[:n_outer], self.outer_shape)) if n_outer else 0
        inner_lead = self.inner_shape[:-2]
        page_idx = int(np.ravel_multi_index(storage[n_outer:], inner_lead)) if inner_lead else 0
        return file_idx, page_idx

    def read(self, index: Tuple = (), out: Optional[np.ndarray] = None) -> np.ndarray:
        """Planes under the integer prefix `index` of the leading dims, in presented order."""
        lead = self.shape[:-2]
        index = tuple(int(i) for i in index)
        rest = lead[len(index):]
        shape = tuple(rest) + self.shape[-2:]
        if out is None:
            out = np.empty(shape, dtype=self.dtype)
        if out.dtype != self.dtype or out.size != int(np.prod(shape)) or not out.flags.c_contiguous:
            raise ValueError(f"destination {out.shape} {out.dtype} does not hold block {shape} {self.dtype}")
        flat = out.reshape((-1,) + self.shape[-2:])
        planes = [self._planes(index + tuple(sub)) for sub in np.ndindex(*rest)] if rest else [self._planes(index)]
        if all(p == 0 for _, p in planes):
            read_files([self.filenames[f] for f, _ in planes], flat, page=0, threads=self.threads)
            return out
        by_file: Dict[int, List[Tuple[int, int]]] = collections.defaultdict(list)
        for k, (f, p) in enumerate(planes):
            by_file[f].append((k, p))
        for f, items in by_file.items():
            with TiffFile(self.filenames[f]) as tif:
                ks = [k for k, _ in items]
                if ks == list(range(ks[0], ks[0] + len(ks))):
                    tif.read_pages([p for _, p in items], flat[ks[0]: ks[0] + len(ks)], threads=self.threads)
                else:
                    for k, p in items:
                        tif.read_pages([p], flat[k: k + 1], threads=1)
        return out

    def __array__(self, dtype=None, copy=None):
        arr = self.read(())
        return arr if dtype is None else arr.astype(dtype)

    def to_numpy(self) -> np.ndarray:
        return self.read(())

    def __array_function__(self, func, types, args, kwargs):
        # duck array: a Dataset keeps the lazy stack as it is; NumPy functions read it
        conv = lambda o: o.read(()) if isinstance(o, TiffTiles) else o   # noqa: E731
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in kwargs.items()})

    def __getitem__(self, key):
        return self.read(())[key]

    def blocks(self) -> Iterator[Tuple[Tuple[int, int], Callable[[np.ndarray], None]]]:
        """((channel, time), fill) for every block of a canonical 6-d stack: `fill(dst)` reads the
        block's pages straight into `dst` -- the chunk protocol of pipeline.ChunkStager.feed."""
        if self.dims != TILE_ORDER:
            raise ValueError(f"blocks() needs dims {TILE_ORDER}, got {self.dims}")
        for ci in range(self.shape[0]):
            for ti in range(self.shape[1]):
                yield (ci, ti), (lambda dst, ci=ci, ti=ti: self.read((ci, ti), dst))

    def __repr__(self):
        return f"TiffTiles(dims={self.dims}, shape={self.shape}, dtype={self.dtype}, files={len(self.filenames)})"


# ---------------------------------------------------------------------------------------------
# read_tiffs  (reader.py:163-326)
# ---------------------------------------------------------------------------------------------
def read_tiffs(xp_dict: Dict[tuple, str], name: str, meta_dict, threads: int = 8) -> Dataset:
    channel_idxs, time_idxs, row_idxs, col_idxs = (sorted(set(idx)) for idx in zip(*xp_dict.keys()))
    dims_in_path, outer_shape = [], ()
    for idxs, dim in ((channel_idxs, "channel"), (time_idxs, "time"), (row_idxs, "tile_row"), (col_idxs, "tile_col")):
        if idxs[0] != -1:
            dims_in_path.append(dim)
            outer_shape += (len(idxs),)
    times = time_idxs if "time" in dims_in_path else None
    channels = channel_idxs if "channel" in dims_in_path else None

    first = next(iter(xp_dict.values()))
    with TiffFile(first) as tif:
        layout = series_layout(tif)
        info = tif.page_info(0)
        if info.status != 0:
            raise _lib.MagnifyB200Error("mgb_tiff_page_info", info.status,
                                        f"{first}: page encoding not supported (compression {info.compression}); "
                                        "only uncompressed pages are staged natively")
        dtype, inner_shape = info.dtype, tuple(layout.shape)
        dims_in_file = [LETTER_TO_DIM[c] for c in layout.axes]          # KeyError for "I"/"S" like the reference
        summary = micromanager_summary(first)
        if times is None and summary is not None and "StartTime" in summary:
            start = datetime.datetime.strptime(summary["StartTime"][:-6], "%Y-%m-%d %H:%M:%S.%f")
            if "time" in dims_in_file:
                stamps = [start + datetime.timedelta(milliseconds=d) for d in layout.delta_t_ms]
                stride = inner_shape[dims_in_file.index("channel")] if "channel" in dims_in_file else 1
                assert len(stamps) % stride == 0
                times = stamps[::stride]
            else:
                times = [start]
        if channels is None and summary is not None and "ChNames" in summary:
            channels = summary["ChNames"]
    if "tile_pos" in dims_in_file:
        k = dims_in_file.index("tile_pos")
        inner_shape = inner_shape[:k] + inner_shape[k + 1:]
        dims_in_file = dims_in_file[:k] + dims_in_file[k + 1:]
    if "depth" in dims_in_file:
        raise ValueError("tiff files with a Z dimension are not yet supported.")
    if "tile_y" not in dims_in_file or "tile_x" not in dims_in_file:
        raise ValueError("tiff files must contain an X and Y dimension.")
    if set(dims_in_file).intersection(dims_in_path):
        raise ValueError("Dimensions specified in the path names and inside the tiff file overlap.")

    filenames = [path for _, path in sorted(xp_dict.items())]
    dims = tuple(dims_in_path + dims_in_file)
    tiles = TiffTiles(filenames, outer_shape, inner_shape, dims, dtype, threads)
    ordered = tuple(d for d in TILE_ORDER if d in dims) + tuple(d for d in dims if d not in TILE_ORDER)
    tiles = tiles.transpose(ordered)

    coords = {}
    if channels is not None:
        coords["channel"] = (("channel",), np.asarray(channels))
    if times is not None:
        coords["time"] = (("time",), np.asarray([int(t.timestamp()) for t in times]))
    for (meta_name, dim), values in meta_dict.items():
        if dim == "time":
            dim_idxs = [datetime.datetime.fromtimestamp(int(i)) for i in coords[dim][1]]
        else:
            dim_idxs = list(coords[dim][1])
        coords[meta_name] = ((dim,), np.asarray([values[i] for i in dim_idxs]))
    return Dataset({"tile": (ordered, tiles)}, coords=coords, attrs={"name": name})


def standardize_format(xp: Dataset) -> Dataset:
    """`standardize_format` (preprocess.py:11-42) for a lazily read tile stack: remember the
    original dims, add the missing canonical dims as length-1 and order them
    (channel, time, tile_row, tile_col, tile_y, tile_x) without reading a page.  Extra
    (non-canonical) dims, which the reference stacks into `time`, are not produced by `read_tiffs`
    and are rejected."""
    tile = xp["tile"]
    tiles = tile.data
    if not isinstance(tiles, TiffTiles):
        raise TypeError("standardize_format here handles the lazy TIFF tile stack of read_tiffs")
    extra = [d for d in tile.dims if d not in TILE_ORDER]
    if extra:
        raise ValueError(f"unexpected tile dims {extra}")
    new = xp.copy()
    new.attrs = dict(xp.attrs, __original_tile_dims__=list(tile.dims))
    new["tile"] = (TILE_ORDER, tiles.transpose(TILE_ORDER))
    return new


class Reader:
    """`read` of the reference (reader.py:23-78) for TIFF inputs and ready-made assays."""

    def __init__(self, threads: int = 8):
        self.threads = threads

    def __call__(self, data) -> Iterator[Dataset]:
        items = [data] if isinstance(data, (str, os.PathLike, Dataset, DataArray)) else list(data)
        for d in items:
            if isinstance(d, (Dataset, DataArray)):
                yield d
                continue
            path_dict, meta_dict = extract_paths(d, assay="str", channel="str", time="time", row="int", col="int")
            if len(path_dict) == 0:
                raise FileNotFoundError(f"The pattern {d} did not lead to any files.")
            path_dict = {(("",) + k[1:]) if k[0] is None else k: v for k, v in path_dict.items()}
            for xp_name in sorted({k[0] for k in path_dict}, key=natural_sort_key):
                xp_dict = {tuple(-1 if x is None else x for x in k[1:]): v
                           for k, v in path_dict.items() if k[0] == xp_name}
                first = next(iter(xp_dict.values()))
                if len(xp_dict) == 1 and os.path.isdir(first):
                    raise NotImplementedError("zarr assays (reader.py:56-65) are storage, not part of the staged path")
                yield read_tiffs(xp_dict, name=xp_name, meta_dict=meta_dict, threads=self.threads)


def natural_sort_key(s: str):
    """utils.natural_sort_key: digit runs compare as numbers."""
    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", s)]

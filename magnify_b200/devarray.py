"""Device-resident arrays that look like (lazy) host arrays to a Dataset.

The components of `magnify_b200.components` exchange `xarray.Dataset`s like the reference's, but
what they put into them are `DeviceArray`s: handles of torch CUDA tensors that satisfy the
duck-array protocol xarray keeps as-is (`shape`, `dtype`, `__array_function__`).  The next GPU
component takes the tensor straight from the handle (no host round trip: the stitched image that
`stitch` emits is the tensor `find_buttons` gathers from), and the host copy is made at most once,
when somebody reads the values -- into pinned memory, on a copy stream, started in the background
as soon as the array is emitted (`prefetch`) so that the device->host traffic of one assay
overlaps the host->device traffic of the next.  This is the role the temp-zarr cache plays in
the reference (`accessor.py:18-35`: results are materialised once and read back lazily).

`transpose` / `reshape` / `expand_dims` / `squeeze` / `moveaxis` stay lazy (torch views), which
is what `Dataset.stack(...).transpose(...)` (find.py:182) needs; anything else materialises.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

_NP_OF = {torch.uint8: np.uint8, torch.uint16: np.uint16, torch.int16: np.int16, torch.int32: np.int32,
          torch.int64: np.int64, torch.float32: np.float32, torch.float64: np.float64, torch.bool: np.bool_}

PREFETCH = True      # start the device->host copy of emitted results in the background
TRACE = None         # diagnostic (bench.py MGB_E2E_TIMELINE): a list collecting (shape, bytes, start, end events) per download
PREFETCH_SKIP = ()   # names of dataset variables that are NOT copied in the background (e.g. ("image",) when the
#                      caller only wants the crops: the reference's drop(roi_only=True), postprocess.py:6-17)


class Streams:
    """The copy streams of one device (created on first use)."""

    _by_device: Dict[int, "Streams"] = {}

    def __init__(self, device: torch.device):
        self.h2d = torch.cuda.Stream(device=device)
        self.d2h = torch.cuda.Stream(device=device)

    @classmethod
    def of(cls, device: torch.device) -> "Streams":
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in cls._by_device:
            cls._by_device[idx] = Streams(torch.device("cuda", idx))
        return cls._by_device[idx]


class _PinnedBlock:
    """One pinned host buffer on loan from the pool.  NumPy arrays made from it (`np.asarray(block)`,
    through `__array_interface__`) keep it alive as their base object; when the last of them -- and
    the DeviceArray that filled it -- is gone, the buffer goes back to the pool.  So a host view a
    caller still holds is never overwritten by a later result."""

    def __init__(self, pool: "PinnedPool", raw: torch.Tensor, shape, dtype: torch.dtype):
        self._pool, self._raw = pool, raw
        self.tensor = raw[: int(np.prod(shape, dtype=np.int64)) * torch.empty(0, dtype=dtype).element_size()].view(dtype).view(tuple(shape))
        self.last_copy = None                        # event after the last device->host copy into the buffer
        npdt = np.dtype(_NP_OF[dtype])
        self.__array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": npdt.str,
                                    "data": (self.tensor.data_ptr(), False), "version": 3}

    def __del__(self):
        pool, raw = self.__dict__.get("_pool"), self.__dict__.get("_raw")
        if pool is not None and raw is not None:
            pool._give_back(raw, self.last_copy)


class PinnedPool:
    """Pinned host buffers, recycled by size.  Page-locking memory costs about half a second per
    GB (measured: 1.7-2.2 s for the 3.9 GB image of an 8-timepoint chip stack, every time), far
    more than copying into it, so the result buffers of one assay are reused by the next."""

    def __init__(self, max_cached_bytes: int = 96 << 30):
        self.max_cached_bytes = max_cached_bytes
        self.free = {}                                # nbytes -> [(raw uint8 tensor, last-copy event)]
        self.cached = 0

    def acquire(self, shape, dtype: torch.dtype, stream=None) -> _PinnedBlock:
        nbytes = max(1, int(np.prod(shape, dtype=np.int64)) * torch.empty(0, dtype=dtype).element_size())
        nbytes = -(-nbytes // 4096) * 4096
        bucket = self.free.get(nbytes)
        if bucket:
            raw, last = bucket.pop()
            self.cached -= nbytes
            if last is not None and stream is not None:
                stream.wait_event(last)               # a copy into the buffer may still be in flight
        else:
            raw = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        return _PinnedBlock(self, raw, shape, dtype)

    def _give_back(self, raw: torch.Tensor, last_copy) -> None:
        n = raw.numel()
        if self.cached + n <= self.max_cached_bytes:
            self.free.setdefault(n, []).append((raw, last_copy))
            self.cached += n

    def clear(self) -> None:
        self.free.clear()
        self.cached = 0


PINNED = PinnedPool()


class LazyArray:
    """Base of the duck arrays: NumPy functions work on the materialised values."""

    def numpy(self) -> np.ndarray:
        raise NotImplementedError

    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: int(np.prod(self.shape, dtype=np.int64)))
    nbytes = property(lambda self: self.size * np.dtype(self.dtype).itemsize)

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        arr = self.numpy()
        return arr if dtype is None else arr.astype(dtype)

    def __array_function__(self, func, types, args, kwargs):
        return func(*_materialise(args), **_materialise(kwargs))

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        return getattr(ufunc, method)(*_materialise(inputs), **_materialise(kwargs))

    def __getitem__(self, key):
        return self.numpy()[key]

    def astype(self, dtype, **kw):
        return self.numpy().astype(dtype, **kw)

    def __repr__(self):
        return f"<{type(self).__name__} shape={tuple(self.shape)} dtype={np.dtype(self.dtype).name}>"


def _materialise(obj):
    if isinstance(obj, LazyArray):
        return obj.numpy()
    if isinstance(obj, (list, tuple)):
        return type(obj)(_materialise(o) for o in obj)
    if isinstance(obj, dict):
        return {k: _materialise(v) for k, v in obj.items()}
    return obj


class DeviceArray(LazyArray):
    """A torch tensor (normally on a CUDA device) standing for a host array of the same shape.

    tensor: any strided view (the x-padded stitched image is one).  as_bool: the tensor holds 0/1
    bytes that read as a boolean array (fg / bg masks).  `extras` carries by-products of the
    kernel that made the array for later components (the summaries the gather already computed).

    Views made by basic indexing / transpose / reshape / expand_dims / squeeze / moveaxis /
    `take` stay lazy and remember how they derive from their root array, as a pair of functions
    per step: one for torch tensors (evaluated on demand for a GPU consumer) and one for NumPy
    arrays.  Only roots are copied to the host (once); a view replays its steps on the root's
    host copy, where they are free NumPy views."""

    def __init__(self, tensor: torch.Tensor, as_bool: bool = False, extras: Optional[dict] = None,
                 _root: Optional["DeviceArray"] = None, _ops: tuple = (), _meta: Optional[torch.Tensor] = None):
        self.as_bool = as_bool
        self.extras = extras if extras is not None else {}
        self._root = _root                         # None for a root (no self-reference: arrays must die by refcount)
        self._ops = _ops
        self._tensor = tensor                      # None for a view until somebody asks for it
        self._meta = _meta if _meta is not None else torch.empty(tuple(tensor.shape), dtype=tensor.dtype, device="meta")
        self._host: Optional[np.ndarray] = None
        self._pending = None                       # root only: (pinned tensor, event) of a copy in flight

    shape = property(lambda self: tuple(self._meta.shape))
    dtype = property(lambda self: np.dtype(np.bool_ if self.as_bool else _NP_OF[self._meta.dtype]))
    root = property(lambda self: self if self._root is None else self._root)

    @property
    def tensor(self) -> torch.Tensor:
        if self._tensor is None:
            t = self.root._tensor
            for torch_fn, _ in self._ops:
                t = torch_fn(t)
            self._tensor = t
        return self._tensor

    @property
    def materialised(self) -> bool:
        return self.root._host is not None

    def prefetch(self) -> "DeviceArray":
        """Queue the device->host copy behind the work already queued on the current stream."""
        root = self.root
        t = root._tensor
        if root._host is not None or root._pending is not None or not t.is_cuda or t.numel() == 0:
            return self
        streams = Streams.of(t.device)
        produced = torch.cuda.Event()
        produced.record(torch.cuda.current_stream(t.device))
        block = PINNED.acquire(tuple(t.shape), t.dtype, streams.d2h)
        with torch.cuda.stream(streams.d2h):
            streams.d2h.wait_event(produced)
            if TRACE is not None:
                began = torch.cuda.Event(enable_timing=True)
                began.record(streams.d2h)
            _copy_to_host(t, block.tensor)
            done = torch.cuda.Event(enable_timing=TRACE is not None)
            done.record(streams.d2h)
            if TRACE is not None:
                TRACE.append((tuple(t.shape), t.numel() * t.element_size(), began, done))
        block.last_copy = done
        t.record_stream(streams.d2h)
        root._pending = (block, done)
        return self

    def _root_numpy(self) -> np.ndarray:
        root = self.root
        if root._host is None:
            t = root._tensor
            if t.is_cuda and t.numel() > 0:
                if root._pending is None:
                    root.prefetch()
                block, done = root._pending
                done.synchronize()
                root._pending = None
                root._host = np.asarray(block)       # the array keeps the pinned block alive (and out of the pool)
            else:
                root._host = t.contiguous().cpu().numpy() if t.is_cuda else t.contiguous().numpy()
        return root._host

    def numpy(self) -> np.ndarray:
        if self._host is None:
            arr = self._root_numpy()
            for _, numpy_fn in self._ops:
                arr = numpy_fn(arr)
            if self is self.root:
                return arr.view(np.bool_) if self.as_bool else arr
            self._host = arr
        return self._host.view(np.bool_) if self.as_bool else self._host

    def _view(self, torch_fn, numpy_fn) -> "DeviceArray":
        return DeviceArray(None, self.as_bool, self.extras, _root=self.root, _ops=self._ops + ((torch_fn, numpy_fn),),
                           _meta=torch_fn(self._meta))

    def __getitem__(self, key):
        """Basic indexing (integers, slices, Ellipsis) stays lazy; anything else reads the values."""
        keys = key if isinstance(key, tuple) else (key,)
        if all(isinstance(k, (int, np.integer, slice)) or k is Ellipsis for k in keys):
            keys = tuple(int(k) if isinstance(k, np.integer) else k for k in keys)
            return self._view(lambda t: t[keys], lambda a: a[keys])
        return self.numpy()[key]

    def take(self, index, axis: int) -> "DeviceArray":
        """Lazy `np.take(self, index, axis)`; a length-1 axis is broadcast instead of copied on the
        host (time-invariant masks, find.py:585-586)."""
        index = np.asarray(index, dtype=np.int64)
        n_src = self.shape[axis]

        def torch_fn(t):
            return t.index_select(axis, torch.as_tensor(index, device=t.device))

        def numpy_fn(a):
            if n_src == 1:
                shape = list(a.shape)
                shape[axis] = len(index)
                return np.broadcast_to(a, shape)
            return np.take(a, index, axis=axis)

        return self._view(torch_fn, numpy_fn)

    # -- lazy view operations -------------------------------------------------------------------
    def transpose(self, *axes):
        axes = axes[0] if len(axes) == 1 and not isinstance(axes[0], (int, np.integer)) else axes
        if axes is None or len(axes) == 0:
            axes = tuple(reversed(range(self.ndim)))
        axes = tuple(int(a) for a in axes)
        return self._view(lambda t: t.permute(*axes), lambda a: np.transpose(a, axes))

    def reshape(self, *shape, **kw):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], (int, np.integer)) else shape
        shape = tuple(int(s) for s in shape)
        return self._view(lambda t: t.reshape(shape), lambda a: np.reshape(a, shape))   # views whenever strides allow

    def _np_transpose(self, a, axes=None):
        return self.transpose(*(axes if axes is not None else ()))

    def _np_reshape(self, a, *args, **kw):
        shape = args[0] if args else kw.get("shape", kw.get("newshape"))
        return self.reshape(shape)

    def _np_moveaxis(self, a, source, destination):
        return self._view(lambda t: torch.movedim(t, source, destination), lambda x: np.moveaxis(x, source, destination))

    def _np_expand_dims(self, a, axis):
        axes = sorted([int(axis)] if isinstance(axis, (int, np.integer)) else [int(x) for x in axis])

        def torch_fn(t):
            for ax in axes:
                t = t.unsqueeze(ax)
            return t

        return self._view(torch_fn, lambda x: np.expand_dims(x, axis))

    def _np_squeeze(self, a, axis=None):
        if axis is None:
            return self._view(lambda t: t.squeeze(), lambda x: np.squeeze(x))
        axes = sorted([int(axis)] if isinstance(axis, (int, np.integer)) else [int(x) for x in axis], reverse=True)

        def torch_fn(t):
            for ax in axes:
                t = t.squeeze(ax)
            return t

        return self._view(torch_fn, lambda x: np.squeeze(x, axis))

    def __array_function__(self, func, types, args, kwargs):
        handler = {np.transpose: self._np_transpose, np.reshape: self._np_reshape, np.moveaxis: self._np_moveaxis,
                   np.expand_dims: self._np_expand_dims, np.squeeze: self._np_squeeze}.get(func)
        if handler is not None and args and args[0] is self:
            return handler(*args, **kwargs)
        return func(*_materialise(args), **_materialise(kwargs))


def _copy_to_host(t: torch.Tensor, host: torch.Tensor) -> None:
    """Device tensor (any strides) -> dense pinned host tensor on the current stream.  An x-padded
    image goes down as ONE pitched copy; other strided views are densified on the device first."""
    from . import ops

    if t.is_contiguous():
        host.copy_(t, non_blocking=True)
        return
    if t.dim() == 4:
        try:
            ops.to_host_dense(t, out=host)
            return
        except ValueError:
            pass
    host.copy_(t.contiguous(), non_blocking=True)


def device_tensor(data) -> Optional[torch.Tensor]:
    """The torch tensor behind a variable's data when it is device-resident, else None."""
    return data.tensor if isinstance(data, DeviceArray) else None

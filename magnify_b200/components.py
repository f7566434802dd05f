"""Reference-facing components: the same names, arguments, exceptions and dataset schema as
magnify's own (`src/magnify/preprocess.py:62-88`, `stitch.py:6-50`, `find.py:13-629`), with the
pixel work done on the GPU through `magnify_b200.ops`.

A component is a callable `Dataset -> Dataset`.  "Dataset" is an `xarray.Dataset` when xarray
is installed, or `magnify_b200.dataset.Assay` (a minimal stand-in with the same accessors).
`install()` registers the factories in `magnify.registry.components` under the reference's own
names so that `mg.mrbles`, `mg.beads` and `mg.microfluidic_chip` pick them up unchanged
(INTEGRATION.md).

Centre finding (`utils.find_circles`, a random search the reference cannot reproduce from run to
run) runs on the GPU with a seeded sampler (`magnify_b200.circles`,
SURVEY.md section 8f rows N1 / N4); `BeadFinder` / `ButtonFinder` also take a `centers` hook that
pins the centres, which is what the bit-exact ROI / mask parity tests use.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Optional

import numpy as np
import torch

from . import chipgrid, circles, ops, pipeline
from .dataset import Assay  # noqa: F401  (re-exported: the stand-in Dataset type of this module)

TILE_DIMS = ("channel", "time", "tile_row", "tile_col", "tile_y", "tile_x")
IMAGE_DIMS = ("channel", "time", "im_y", "im_x")


# ---------------------------------------------------------------------------------------------
# dataset adapters (xarray.Dataset or Assay)
# ---------------------------------------------------------------------------------------------
def _to_numpy(var) -> np.ndarray:
    return var.to_numpy() if hasattr(var, "to_numpy") else np.asarray(var)


def _device(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("magnify_b200 components need a CUDA device (there is no CPU fallback)")
    return torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")


def _image_to_device(assay, dev) -> torch.Tensor:
    """Upload assay["image"] into an x-padded device image (aligned rows for the staged gather)."""
    arr = np.ascontiguousarray(_to_numpy(assay["image"]))
    host = torch.from_numpy(arr)
    image = ops.alloc_image(host.shape, host.dtype, dev)
    image.copy_(host)
    return image


def _tiles_to_device(assay, dev) -> torch.Tensor:
    tile = assay["tile"]
    if tuple(tile.dims) != TILE_DIMS:
        raise ValueError(f"tile must have dims {TILE_DIMS} (run standardize_format first), got {tuple(tile.dims)}")
    values = getattr(tile, "values", None)
    if hasattr(values, "blocks"):
        # lazily read TIFF tiles (reader.TiffTiles): pages land in a pinned buffer, then HBM
        host = torch.empty(tuple(values.shape), dtype=getattr(torch, str(values.dtype)), pin_memory=True)
        values.read((), host.numpy())
        return host.to(dev, non_blocking=True)
    return torch.from_numpy(np.ascontiguousarray(_to_numpy(tile))).to(dev)


def _roi_to_device(roi: np.ndarray, dev) -> torch.Tensor:
    """Upload a roi for the reductions: uint16 as is; any other dtype the reference accepts
    (float images, tests/test_chip.py:76-96; 8-bit) as float32, which represents uint8 / int16 /
    float32 values exactly."""
    roi = np.ascontiguousarray(roi)
    if roi.dtype == np.uint16:
        return torch.from_numpy(roi).to(dev)
    if roi.dtype in (np.float32, np.uint8, np.int8, np.int16):
        return torch.from_numpy(roi.astype(np.float32)).to(dev)
    raise TypeError(f"roi dtype {roi.dtype} is not supported by the GPU reductions (uint16, float32, 8/16-bit integers)")


def _read_tiff(path) -> np.ndarray:
    """Flat-field / dark-field image file (preprocess.py:75-81 reads it with tifffile): first page,
    through the native uncompressed-TIFF reader."""
    from . import reader

    with reader.TiffFile(os.fspath(path)) as tif:
        return tif.asarray(0)


# ---------------------------------------------------------------------------------------------
# flatfield_correct  (preprocess.py:62-88)
# ---------------------------------------------------------------------------------------------
def flatfield_correct(xp, flatfield=1.0, darkfield=0.0, device=None):
    if isinstance(flatfield, (str, os.PathLike)):
        flatfield = _read_tiff(os.path.expanduser(flatfield))
    if isinstance(darkfield, (str, os.PathLike)):
        darkfield = _read_tiff(os.path.expanduser(darkfield))
    dev = _device(device)
    tiles = _tiles_to_device(xp, dev)
    out = ops.flatfield_correct(tiles, flatfield, darkfield)
    xp["tile"] = (TILE_DIMS, out.cpu().numpy())
    return xp


def make_flatfield_correct(flatfield=1.0, darkfield=0.0, device=None):
    return lambda xp: flatfield_correct(xp, flatfield=flatfield, darkfield=darkfield, device=device)


# ---------------------------------------------------------------------------------------------
# stitch  (stitch.py:6-50)
# ---------------------------------------------------------------------------------------------
class Stitcher:
    def __init__(self, overlap: int = 102, device=None):
        if overlap < 0:
            raise ValueError("Overlap must be non-negative.")  # stitch.py:8-9
        self.overlap = overlap
        self.device = device

    def __call__(self, assay):
        if "tile" not in assay:
            raise AttributeError("Dataset must contain 'tile' data variable.")  # stitch.py:13-14
        sizes = assay.sizes
        ops.check_overlap(self.overlap, sizes["tile_y"], sizes["tile_x"])  # stitch.py:16-20
        dev = _device(self.device)
        image = ops.stitch(_tiles_to_device(assay, dev), self.overlap)
        assay["image"] = (IMAGE_DIMS, ops.to_host_dense(image, non_blocking=False).numpy())
        return assay


def make_stitch(overlap: int = 102, device=None):
    return Stitcher(overlap=overlap, device=device)


class FlatfieldStitcher:
    """flatfield_correct + stitch in one pass over the tiles (fused kernel); equivalent to the
    two reference components back to back, minus the corrected `tile` variable (which the
    predefined pipelines drop anyway, postprocess.py:6-17)."""

    def __init__(self, flatfield=1.0, darkfield=0.0, overlap: int = 102, device=None):
        if overlap < 0:
            raise ValueError("Overlap must be non-negative.")
        self.flatfield, self.darkfield, self.overlap, self.device = flatfield, darkfield, overlap, device

    def __call__(self, assay):
        if "tile" not in assay:
            raise AttributeError("Dataset must contain 'tile' data variable.")
        sizes = assay.sizes
        ops.check_overlap(self.overlap, sizes["tile_y"], sizes["tile_x"])
        flat, dark = self.flatfield, self.darkfield
        if isinstance(flat, (str, os.PathLike)):
            flat = _read_tiff(os.path.expanduser(flat))
        if isinstance(dark, (str, os.PathLike)):
            dark = _read_tiff(os.path.expanduser(dark))
        dev = _device(self.device)
        image = ops.flatfield_stitch(_tiles_to_device(assay, dev), flat, dark, overlap=self.overlap)
        assay["image"] = (IMAGE_DIMS, ops.to_host_dense(image, non_blocking=False).numpy())
        return assay


# ---------------------------------------------------------------------------------------------
# find_beads  (find.py:445-629)
# ---------------------------------------------------------------------------------------------
class BeadFinder:
    """ROI/mask half of the reference BeadFinder on the GPU (find.py:503-605).

    centers: ndarray (M,3) of (row, col, radius) or a callable `assay -> ndarray` pins the bead
    centres; when omitted they are found on the GPU (`magnify_b200.circles.find_circles`, the
    reference's `utils.find_circles` with a seeded sampler; find.py:476-501)."""

    def __init__(self, min_bead_diameter: int, max_bead_diameter: int, low_edge_quantile: float = 0.1,
                 high_edge_quantile: float = 0.9, num_iter: int = 5000000, min_roundness: float = 0.3,
                 roi_length: Optional[int] = None, search_channel=None, interactive: bool = False,
                 centers=None, device=None, seed: int = 0):
        if min_bead_diameter > max_bead_diameter:
            raise ValueError("min_bead_diameter must be <= max_bead_diameter.")  # find.py:458-459
        self.min_bead_radius = math.floor(min_bead_diameter / 2)
        self.max_bead_radius = math.ceil(max_bead_diameter / 2)
        self.low_edge_quantile, self.high_edge_quantile = low_edge_quantile, high_edge_quantile
        self.num_iter, self.min_roundness = num_iter, min_roundness
        self.roi_length = roi_length if roi_length is not None else 2 * max_bead_diameter  # find.py:467
        self.search_channels = [] if search_channel is None else (
            [search_channel] if isinstance(search_channel, str) else list(search_channel))
        self.interactive = interactive
        self.centers = centers
        self.device = device
        self.seed = seed

    def find_centers(self, assay, image: Optional[torch.Tensor] = None) -> np.ndarray:
        """(M, 3) float64 (row, col, radius): the pinned `centers`, or the GPU circle finder run on
        time 0 of every search channel with the reference's parameters (find.py:476-501)."""
        if self.centers is not None:
            beads = self.centers(assay) if callable(self.centers) else self.centers
            return np.asarray(beads, dtype=np.float64).reshape(-1, 3)
        from scipy.spatial import cKDTree

        dev = _device(self.device)
        if image is None:
            image = _image_to_device(assay, dev)
        names = list(_to_numpy(assay["channel"])) if "channel" in assay else list(range(image.shape[0]))
        channels = self.search_channels or names
        beads = np.empty((0, 3))
        for k, ch in enumerate(channels):
            img = circles.to_uint8(image[names.index(ch), 0].contiguous())
            found = circles.find_circles(img, self.low_edge_quantile, self.high_edge_quantile, 20, self.num_iter,
                                         self.min_bead_radius, self.max_bead_radius, self.min_roundness,
                                         self.min_bead_radius, seed=self.seed + k)[0].astype(np.float64)
            if len(beads) > 0 and len(found) > 0:   # drop beads already seen in another channel, find.py:491-500
                near = cKDTree(beads[:, :2]).query_ball_point(found[:, :2], 2 * self.min_bead_radius)
                found = found[~np.array([len(n) > 0 for n in near], dtype=bool)]
            beads = np.concatenate([beads, found])
        return beads

    def __call__(self, assay):
        dev = _device(self.device)
        image = _image_to_device(assay, dev)
        beads = self.find_centers(assay, image)
        c, t, him, wim = image.shape
        length = self.roi_length
        m = len(beads)
        x = np.repeat(beads[:, 1:2], t, axis=1)  # find.py:543-550
        y = np.repeat(beads[:, 0:1], t, axis=1)
        valid = np.ones((m, t), dtype=bool)  # find.py:551-554
        if m == 0:  # find.py:557-558
            roi = np.empty((0, c, t, length, length), dtype=_to_numpy(assay["image"]).dtype)
            fg = np.empty((0, t, length, length), dtype=bool)
            bg = fg.copy()
        else:
            xd = torch.from_numpy(x).to(dev).contiguous()
            yd = torch.from_numpy(y).to(dev).contiguous()
            boxes = ops.bounding_boxes(xd, yd, length, wim, him)
            labels = ops.bead_labels(torch.from_numpy(beads.astype(np.int64).astype(np.int32)).to(dev), him, wim)
            fg_d, bg_d = ops.bead_masks(labels, boxes[:, 0].contiguous(), length)
            roi = ops.roi_gather(image, boxes, length).cpu().numpy()
            # masks are time invariant (find.py:585-586): broadcast views, not copies
            fg = np.broadcast_to(fg_d.cpu().numpy().astype(bool)[:, None], (m, t, length, length))
            bg = np.broadcast_to(bg_d.cpu().numpy().astype(bool)[:, None], (m, t, length, length))
        assay["roi"] = (("mark", "channel", "time", "roi_y", "roi_x"), roi)
        return assay.assign_coords(
            fg=(("mark", "time", "roi_y", "roi_x"), fg), bg=(("mark", "time", "roi_y", "roi_x"), bg),
            x=(("mark", "time"), x), y=(("mark", "time"), y), valid=(("mark", "time"), valid))


def make_find_beads(min_bead_diameter: int, max_bead_diameter: int, low_edge_quantile: float,
                    high_edge_quantile: float, num_iter: int, min_roundness: float, roi_length: int,
                    search_channel, interactive: bool, centers=None, device=None, seed: int = 0):
    return BeadFinder(min_bead_diameter=min_bead_diameter, max_bead_diameter=max_bead_diameter,
                      low_edge_quantile=low_edge_quantile, high_edge_quantile=high_edge_quantile, num_iter=num_iter,
                      min_roundness=min_roundness, roi_length=roi_length, search_channel=search_channel,
                      interactive=interactive, centers=centers, device=device, seed=seed)


# ---------------------------------------------------------------------------------------------
# find_buttons  (find.py:13-442)
# ---------------------------------------------------------------------------------------------
class ButtonFinder:
    """ROI/mask half of the reference ButtonFinder on the GPU (find.py:143-181, 308-402).

    centers: callable `(assay, t) -> (x, y, fg_radius)` giving, for a search timestep t, the final
    button centres in image coordinates (rows x cols float64 each) and the foreground radii
    (rows x cols int; max_button_radius where refinement found nothing, find.py:363,378).  When
    omitted, centres are found here: full-image circle finding on the GPU, the reference's grid
    clustering on the host (`chipgrid`), and the per-chamber refinement as one GPU batch
    (find.py:205-306, 336-378; seeded, unlike the reference)."""

    def __init__(self, row_dist: float, col_dist: float, min_button_diameter: int, max_button_diameter: int,
                 chamber_diameter: int, top_chamber=None, left_chamber=None, low_edge_quantile: float = 0.1,
                 high_edge_quantile: float = 0.9, num_iter: int = 5000000, min_roundness: float = 0.20,
                 cluster_penalty: float = 10, roi_length: Optional[int] = None, progress_bar: bool = False,
                 search_timestep=0, search_channel=None, interactive: bool = False, centers: Optional[Callable] = None,
                 device=None, seed: int = 0):
        if min_button_diameter > max_button_diameter:
            raise ValueError("min_button_diameter must be <= max_button_diameter.")  # find.py:34-35
        self.row_dist, self.col_dist = row_dist, col_dist
        self.min_button_radius = math.floor(min_button_diameter / 2)
        self.max_button_radius = math.ceil(max_button_diameter / 2)
        self.chamber_radius = round(chamber_diameter / 2)
        self.roi_length = roi_length if roi_length is not None else round(1.2 * chamber_diameter)  # find.py:49
        self.search_timesteps = sorted(np.atleast_1d(search_timestep).astype(int).tolist())
        self.kwargs = dict(row_dist=row_dist, col_dist=col_dist, min_button_diameter=min_button_diameter,
                           max_button_diameter=max_button_diameter, chamber_diameter=chamber_diameter,
                           top_chamber=top_chamber, left_chamber=left_chamber, low_edge_quantile=low_edge_quantile,
                           high_edge_quantile=high_edge_quantile, num_iter=num_iter, min_roundness=min_roundness,
                           cluster_penalty=cluster_penalty, roi_length=roi_length, progress_bar=progress_bar,
                           search_timestep=search_timestep, search_channel=search_channel, interactive=interactive)
        self.centers = centers
        self.device = device
        self.seed = seed

    def _search_channel_indexes(self, assay, n_channels: int):
        names = list(_to_numpy(assay["channel"])) if "channel" in assay else list(range(n_channels))
        wanted = self.kwargs["search_channel"]
        if wanted is None:
            return list(range(n_channels))
        wanted = [wanted] if isinstance(wanted, str) or np.isscalar(wanted) else list(wanted)
        return [names.index(ch) for ch in wanted]

    def find_centers(self, assay, image: torch.Tensor, t: int):
        """find.py:205-306: circles of every search channel's full image at time t (GPU), merged
        across channels, clustered into the chip grid and intersected (host) -> (mark_x, mark_y)."""
        kw = self.kwargs
        tag = _to_numpy(assay["tag"])
        points = np.empty((0, 2))
        for k, ch in enumerate(self._search_channel_indexes(assay, image.shape[0])):
            img = circles.to_uint8(image[ch, t].contiguous())
            found = circles.find_circles(img, kw["low_edge_quantile"], kw["high_edge_quantile"], 20, kw["num_iter"],
                                         self.min_button_radius, self.max_button_radius, kw["min_roundness"],
                                         self.chamber_radius, seed=self.seed + 1000 * t + k)[0]
            points = chipgrid.merge_channel_points(points, found[:, :2].astype(np.float64), self.chamber_radius)
        return chipgrid.grid_centers(points, tag, tuple(image.shape[-2:]), self.row_dist, self.col_dist,
                                     self.chamber_radius, kw["top_chamber"], kw["left_chamber"], kw["cluster_penalty"])

    def refine(self, assay, image: torch.Tensor, t: int, x: np.ndarray, y: np.ndarray):
        """find.py:324-378: crop every chamber at its grid position, look for the best circle in
        the crop of each search channel (batched on the GPU) and recentre on it.  Returns refined
        x, y (rows, cols) and the foreground radius per chamber (max_button_radius when nothing
        was found or the chamber is blank)."""
        kw = self.kwargs
        dev = image.device
        rows, cols = x.shape
        m, length = rows * cols, self.roi_length
        him, wim = image.shape[-2:]
        tag = _to_numpy(assay["tag"]).reshape(m)
        chans = self._search_channel_indexes(assay, image.shape[0])
        planes = ops.alloc_image((len(chans), 1, him, wim), image.dtype, dev)
        for k, ch in enumerate(chans):
            planes[k, 0].copy_(image[ch, t])
        xd = torch.from_numpy(np.ascontiguousarray(x.reshape(m, 1))).to(dev)
        yd = torch.from_numpy(np.ascontiguousarray(y.reshape(m, 1))).to(dev)
        boxes = ops.bounding_boxes(xd, yd, length, wim, him)
        crops = ops.roi_gather(planes, boxes, length)                           # (m, channels, 1, L, L)
        batch = circles.to_uint8(crops.reshape(m * len(chans), length, length), batched=True)   # find.py:343
        high_q = 1 - np.pi * self.min_button_radius / length**2                 # find.py:345-347
        found = circles.find_circles(batch, kw["low_edge_quantile"], high_q, 20, kw["num_iter"] // (rows * cols),
                                     self.min_button_radius, self.max_button_radius, kw["min_roundness"], 0,
                                     seed=self.seed + 1000 * t + 500)
        top_left = boxes[:, 0].cpu().numpy()
        x, y = x.reshape(m).copy(), y.reshape(m).copy()
        radius = np.full(m, self.max_button_radius, dtype=np.int32)
        for i in range(m):
            if tag[i] == "":
                continue
            best, best_score = None, -np.inf
            for k in range(len(chans)):
                cand, scores = found[i * len(chans) + k]
                if len(cand) > 0 and scores[0] > best_score:                    # best first: [0] is the argmax
                    best, best_score = cand[0], scores[0]
            if best is not None:
                y[i], x[i] = best[0] + top_left[i, 0], best[1] + top_left[i, 1]
                radius[i] = best[2]
        return x.reshape(rows, cols), y.reshape(rows, cols), radius

    def _gpu_centers(self, assay, image: torch.Tensor, t: int):
        x, y = self.find_centers(assay, image, t)
        return self.refine(assay, image, t, np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))

    def __call__(self, assay):
        dev = _device(self.device)
        image = _image_to_device(assay, dev)
        c, t, him, wim = image.shape
        rows, cols = assay["tag"].shape
        m = rows * cols
        length = self.roi_length
        src = pipeline.copy_forward_sources(t, self.search_timesteps)  # find.py:143-151
        search = sorted(set(src.tolist()))
        x = np.empty((rows, cols, t))
        y = np.empty((rows, cols, t))
        radius = np.empty((m, len(search)), dtype=np.int32)
        for k, ts in enumerate(search):
            xs, ys, rs = self.centers(assay, ts) if self.centers is not None else self._gpu_centers(assay, image, ts)
            x[..., ts], y[..., ts], radius[:, k] = xs, ys, np.asarray(rs).reshape(m)
        for ti in range(t):  # copy-forward of the centres, find.py:156-157,174-175
            x[..., ti], y[..., ti] = x[..., src[ti]], y[..., src[ti]]
        xd = torch.from_numpy(np.ascontiguousarray(x.reshape(m, t))).to(dev)
        yd = torch.from_numpy(np.ascontiguousarray(y.reshape(m, t))).to(dev)
        boxes, rel = ops.bounding_boxes(xd, yd, length, wim, him, want_rel=True)
        fgs, bgs = [], []
        for k, ts in enumerate(search):
            f, b = ops.chip_masks(rel[:, ts].contiguous(), torch.from_numpy(radius[:, k].copy()).to(dev),
                                  self.max_button_radius, self.chamber_radius, length)
            fgs.append(f.cpu().numpy().astype(bool))
            bgs.append(b.cpu().numpy().astype(bool))
        index = {ts: k for k, ts in enumerate(search)}
        mask_t = np.array([index[int(s)] for s in src])
        fg = np.stack(fgs, 1)[:, mask_t]  # find.py:172-173
        bg = np.stack(bgs, 1)[:, mask_t]
        roi = ops.roi_gather(image, boxes, length).cpu().numpy()
        valid = _to_numpy(assay["valid"]) if "valid" in assay else np.ones((rows, cols, t), dtype=bool)
        valid = valid[:, :, src] if valid.ndim == 3 else valid
        return self._emit(assay, roi, fg, bg, x, y, valid, rows, cols, t, length)

    @staticmethod
    def _emit(assay, roi, fg, bg, x, y, valid, rows, cols, t, length):
        c = roi.shape[1]
        try:
            import xarray as xr
        except Exception:
            xr = None
        if xr is not None and isinstance(assay, xr.Dataset):
            assay["roi"] = (("mark_row", "mark_col", "channel", "time", "roi_y", "roi_x"),
                            roi.reshape(rows, cols, c, t, length, length))
            assay = assay.assign_coords(
                fg=(("mark_row", "mark_col", "time", "roi_y", "roi_x"), fg.reshape(rows, cols, t, length, length)),
                bg=(("mark_row", "mark_col", "time", "roi_y", "roi_x"), bg.reshape(rows, cols, t, length, length)),
                x=(("mark_row", "mark_col", "time"), x), y=(("mark_row", "mark_col", "time"), y),
                valid=(("mark_row", "mark_col", "time"), valid.reshape(rows, cols, t)))
            return assay.stack(mark=("mark_row", "mark_col"), create_index=True).transpose("mark", ...)  # find.py:182
        # Assay stand-in: the stacked `mark` dimension directly, row-major like xarray's stack
        mr, mc = np.divmod(np.arange(rows * cols), cols)
        assay = assay.drop_vars(["tag", "valid"], errors="ignore").assign_coords(
            mark_row=(("mark",), mr), mark_col=(("mark",), mc),
            tag=(("mark",), _to_numpy(assay["tag"]).reshape(-1)) if "tag" in assay else (("mark",), np.full(rows * cols, "default")),
            fg=(("mark", "time", "roi_y", "roi_x"), fg), bg=(("mark", "time", "roi_y", "roi_x"), bg),
            x=(("mark", "time"), x.reshape(rows * cols, t)), y=(("mark", "time"), y.reshape(rows * cols, t)),
            valid=(("mark", "time"), np.asarray(valid).reshape(rows * cols, t)))
        assay["roi"] = (("mark", "channel", "time", "roi_y", "roi_x"), roi)
        return assay


def make_find_buttons(row_dist, col_dist, min_button_diameter, max_button_diameter, chamber_diameter, top_chamber,
                      left_chamber, low_edge_quantile, high_edge_quantile, num_iter, min_roundness, cluster_penalty,
                      roi_length, progress_bar, search_timestep, search_channel, interactive, centers=None, device=None):
    return ButtonFinder(row_dist, col_dist, min_button_diameter, max_button_diameter, chamber_diameter, top_chamber,
                        left_chamber, low_edge_quantile, high_edge_quantile, num_iter, min_roundness, cluster_penalty,
                        roi_length, progress_bar, search_timestep, search_channel, interactive, centers=centers,
                        device=device)


# ---------------------------------------------------------------------------------------------
# quantify: the per-marker summaries the reference's consumers compute with xarray expressions
# (identify.py:76-80, filter.py:21-22,51,74,82, README.md:21-22)
# ---------------------------------------------------------------------------------------------
def quantify(assay, median: bool = True, device=None):
    """Adds fg_count/bg_count (mark,time) and fg_sum, bg_sum, fg_mean, bg_mean[, fg_median,
    bg_median] (mark,channel,time) computed from roi/fg/bg: `roi.where(fg).mean(["roi_x","roi_y"])`
    etc.  uint16 roi only."""
    dev = _device(device)
    roi = _roi_to_device(_to_numpy(assay["roi"]), dev)
    m, c, t, length, _ = roi.shape
    fg = torch.from_numpy(np.ascontiguousarray(_to_numpy(assay["fg"])).view(np.uint8)).to(dev)
    bg = torch.from_numpy(np.ascontiguousarray(_to_numpy(assay["bg"])).view(np.uint8)).to(dev)
    dims = ("mark", "channel", "time")
    if m == 0:
        empty = np.empty((0, c, t))
        for name in ("fg_sum", "bg_sum", "fg_mean", "bg_mean") + (("fg_median", "bg_median") if median else ()):
            assay[name] = (dims, empty.copy())
        return assay
    stats = ops.roi_stats(roi, fg, bg)
    s = stats.cpu().numpy()
    assay["fg_sum"], assay["bg_sum"] = (dims, s[..., 2]), (dims, s[..., 3])
    assay["fg_mean"], assay["bg_mean"] = (dims, s[..., 4]), (dims, s[..., 5])
    assay["fg_count"], assay["bg_count"] = (("mark", "time"), s[:, 0, :, 0]), (("mark", "time"), s[:, 0, :, 1])
    if median:
        assay["fg_median"] = (dims, ops.roi_median(roi, fg).cpu().numpy())
        assay["bg_median"] = (dims, ops.roi_median(roi, bg).cpu().numpy())
    return assay


def _time0_medians(assay, channel_index: int, dev):
    """GPU medians of fg and bg at time 0 for one channel -> two (M,) float64 arrays."""
    roi = np.ascontiguousarray(_to_numpy(assay["roi"])[:, channel_index : channel_index + 1, :1])
    fg = np.ascontiguousarray(_to_numpy(assay["fg"])[:, :1]).view(np.uint8)
    bg = np.ascontiguousarray(_to_numpy(assay["bg"])[:, :1]).view(np.uint8)
    roi_d = _roi_to_device(roi, dev)
    fgm = ops.roi_median(roi_d, torch.from_numpy(fg).to(dev)).cpu().numpy()[:, 0, 0]
    bgm = ops.roi_median(roi_d, torch.from_numpy(bg).to(dev)).cpu().numpy()[:, 0, 0]
    return fgm, bgm


def filter_expression(assay, search_channel=None, min_contrast=None, device=None):
    """`filter_expression` of the reference (filter.py:11-37) with the masked medians computed on
    the GPU.  The pairwise-difference statistic is the reference's
    own expression (O(M^2) on the host, like the reference)."""
    dev = _device(device)
    channels = list(_to_numpy(assay["channel"])) if "channel" in assay else list(range(assay["roi"].shape[1]))
    wanted = channels if search_channel is None else ([search_channel] if isinstance(search_channel, str)
                                                      or np.isscalar(search_channel) else list(search_channel))
    valid = _to_numpy(assay["valid"]).astype(bool)
    expressed = np.zeros_like(valid)
    for ch in wanted:
        fgm, bgm = _time0_medians(assay, channels.index(ch), dev)
        if min_contrast is None:
            diffs = bgm[:, np.newaxis] - bgm[np.newaxis, :]
            offdiag = np.ones_like(diffs, dtype=bool) & (~np.eye(len(diffs), dtype=bool))
            upper = 4 * diffs[offdiag].std()
        else:
            upper = min_contrast
        hit = fgm - bgm > upper
        expressed |= hit.reshape((-1,) + (1,) * (valid.ndim - 1))
    return assay.assign_coords(valid=(tuple(assay["valid"].dims), valid & expressed))


def filter_leaky(assay, search_channel=None, device=None):
    """`filter_leaky` of the reference (filter.py:65-94) with the masked medians computed on the
    GPU: tagged markers whose blank neighbour (previous / next mark in stacked order) is not
    "empty" are invalidated.  Needs the stacked chip schema (`tag`, `mark_row` per mark)."""
    dev = _device(device)
    channels = list(_to_numpy(assay["channel"])) if "channel" in assay else list(range(assay["roi"].shape[1]))
    wanted = channels if search_channel is None else ([search_channel] if isinstance(search_channel, str)
                                                      or np.isscalar(search_channel) else list(search_channel))
    tag = _to_numpy(assay["tag"])
    rows = _to_numpy(assay["mark_row"])
    valid = _to_numpy(assay["valid"]).astype(bool).copy()
    top = rows.max() if rows.size else 0
    for ch in wanted:
        fgm, bgm = _time0_medians(assay, channels.index(ch), dev)
        diffs = bgm[:, np.newaxis] - bgm[np.newaxis, :]
        offdiag = np.ones_like(diffs, dtype=bool) & (~np.eye(len(diffs), dtype=bool))
        empty = fgm - bgm < 5 * diffs[offdiag].std()
        for i in range(len(tag)):
            if tag[i] == "":
                continue
            if rows[i] > 0 and tag[i - 1] == "":
                valid[i] &= empty[i - 1]
            if rows[i] < top and tag[i + 1] == "":
                valid[i] &= empty[i + 1]
    return assay.assign_coords(valid=(tuple(assay["valid"].dims), valid))


def filter_nonround(assay, min_roundness: float = 0.75, search_channel=None, device=None):
    """`filter_nonround` of the reference (filter.py:40-62): markers whose time-0 foreground has a
    roundness 4 pi area / perimeter^2 not above `min_roundness` (or no contour at all) are
    invalidated.  Perimeters come from the GPU border-following kernel (`ops.mask_perimeters`).  The
    reference repeats the same test once per search channel although `fg` has no channel axis; the
    outcome is that of a single pass."""
    dev = _device(device)
    fg0 = np.ascontiguousarray(_to_numpy(assay["fg"])[:, 0])
    valid = _to_numpy(assay["valid"]).astype(bool).copy()
    if fg0.shape[0]:
        perimeter = ops.mask_perimeters(torch.from_numpy(fg0.view(np.uint8)).to(dev)).cpu().numpy()
        areas = fg0.sum(axis=(1, 2))
        with np.errstate(divide="ignore", invalid="ignore"):
            roundness = 4 * np.pi * areas.astype(np.float64) / perimeter**2
        keep = (perimeter != 0) & (roundness > min_roundness)
        valid &= keep.reshape((-1,) + (1,) * (valid.ndim - 1))
    return assay.assign_coords(valid=(tuple(assay["valid"].dims), valid))


def mrbles_intensities(assay, channels=None, device=None) -> np.ndarray:
    """The per-bead intensities `identify_mrbles` starts from (identify.py:76-80): mean of the
    foreground minus median of the background at time 0, (mark, channel)."""
    dev = _device(device)
    names = list(_to_numpy(assay["channel"])) if "channel" in assay else list(range(assay["roi"].shape[1]))
    idx = list(range(len(names))) if channels is None else [names.index(c) for c in channels]
    roi = _roi_to_device(_to_numpy(assay["roi"])[:, idx, :1], dev)
    fg = torch.from_numpy(np.ascontiguousarray(_to_numpy(assay["fg"])[:, :1]).view(np.uint8)).to(dev)
    bg = torch.from_numpy(np.ascontiguousarray(_to_numpy(assay["bg"])[:, :1]).view(np.uint8)).to(dev)
    mean_fg = ops.roi_stats(roi, fg, bg)[:, :, 0, 4]
    med_bg = ops.roi_median(roi, bg)[:, :, 0]
    return (mean_fg - med_bg).cpu().numpy()


def make_quantify(median: bool = True, device=None):
    return lambda xp: quantify(xp, median=median, device=device)


# ---------------------------------------------------------------------------------------------
# registration
# ---------------------------------------------------------------------------------------------
def _component(fn):
    """The reference's `@registry.component` idiom (registry.py:16-29): factory(**kwargs) returns a
    callable that applies `fn(assay, **kwargs)`."""
    return lambda **kwargs: (lambda xp: fn(xp, **kwargs))


FACTORIES = {
    "flatfield_correct": make_flatfield_correct,
    "stitch": make_stitch,
    "find_beads": make_find_beads,
    "find_buttons": make_find_buttons,
    "filter_expression": _component(filter_expression),      # filter.py:11-37
    "filter_nonround": _component(filter_nonround),          # filter.py:40-62
    "filter_leaky": _component(filter_leaky),                # filter.py:65-94
}
EXTRA_FACTORIES = {
    "flatfield_stitch_b200": lambda flatfield=1.0, darkfield=0.0, overlap=102, device=None: FlatfieldStitcher(
        flatfield, darkfield, overlap, device),
    "quantify": make_quantify,
}


def install(override: bool = True):
    """Register the GPU components in magnify's registry (registry.py:12-13).  With override the
    reference's own names are replaced, so the predefined pipelines use them unchanged; the
    `<name>_b200` aliases and "quantify" are always added."""
    try:
        import magnify.registry as registry
    except Exception as e:
        raise ImportError("magnify (and its dependencies xarray, dask, catalogue) must be importable to "
                          "install the magnify_b200 components into its registry") from e
    for name, factory in FACTORIES.items():
        registry.components.register(name + "_b200")(factory)
        if override:
            registry.components.register(name)(factory)
    for name, factory in EXTRA_FACTORIES.items():
        registry.components.register(name)(factory)
    return sorted(list(FACTORIES) + [n + "_b200" for n in FACTORIES] + list(EXTRA_FACTORIES))

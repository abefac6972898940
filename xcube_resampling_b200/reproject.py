"""reproject_dataset: regular source grid -> regular target grid in another CRS.

Same entry point, arguments and error behaviour as the reference's
``xcube_resampling/reproject.py:51-186``.  Underneath, the reference's four dask stages -- PROJ
transform of every target pixel centre (``:472-496``), per-tile source windows (``:385-469``), the
padded / re-tiled source copy (``:499-530``) and the numpy fancy-index gather (``:268-335``) --
collapse into ONE kernel of ``libxrs.so``: ``xrs_reproject`` computes the target -> source CRS
transform in fp64 registers and gathers every band in the same pass, reading the source in place.
What survives of the per-tile logic is exactly what shapes the numbers: each reference tile's
window origin is a FLOAT32 coordinate (``:427-450``) and fractional indices are measured from it.

The device-level functions (:class:`ReprojectPlan`, :func:`transform_points_dev`) take and return
torch tensors used purely as device buffers.
"""

from __future__ import annotations

import ctypes
import math
from collections.abc import Iterable

import numpy as np
import torch

from . import _dev
from ._lib import XrsProj, check, load
from .constants import DTYPE_CODES, INTERP_CODES, SCALE_LIMIT
from .crs import normalize_crs
from .dataset import DataArray, Dataset, from_any, to_like
from .gridmapping import GridMapping
from .utils import (
    _get_fill_value,
    _get_interp_method_str,
    _prep_interp_methods_downscale,
    _select_variables,
    clip_dataset_by_bbox,
    normalize_grid_mapping,
)

_DENSIFY_PTS = 21  # pyproj.Transformer.transform_bounds default


# ---------------------------------------------------------------------------
# device level: point transforms
# ---------------------------------------------------------------------------
def transform_points_dev(x: torch.Tensor, y: torch.Tensor, from_crs, to_crs, out=None):
    """``Transformer.from_crs(from_crs, to_crs, always_xy=True).transform(x, y)`` on float64 device
    tensors of any (equal) shape; returns two new tensors (or writes ``out=(ox, oy)``)."""
    lib = load()
    if x.dtype != torch.float64 or y.dtype != torch.float64:
        raise TypeError("coordinates must be float64 device tensors")
    if x.shape != y.shape:
        raise ValueError("x and y must have the same shape")
    x, y = x.contiguous(), y.contiguous()
    ox, oy = (torch.empty_like(x), torch.empty_like(y)) if out is None else out
    p_from, p_to = XrsProj.from_crs(normalize_crs(from_crs)), XrsProj.from_crs(normalize_crs(to_crs))
    check(lib.xrs_transform_points(ctypes.addressof(p_from), ctypes.addressof(p_to), _dev.ptr(x), _dev.ptr(y),
                                   _dev.ptr(ox), _dev.ptr(oy), x.numel(), _dev.stream_ptr(x.device)),
          "xrs_transform_points")
    return ox, oy


def transform_points(x, y, from_crs, to_crs, device=None):
    """Host arrays in, host arrays out (float64), through the device kernel."""
    xs = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    ys = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    xs, ys = np.broadcast_arrays(xs, ys)
    ox, oy = transform_points_dev(_dev.to_device(xs, device), _dev.to_device(ys, device), from_crs, to_crs)
    return _dev.to_host(ox), _dev.to_host(oy)


def _edge_points(boxes: np.ndarray, densify_pts: int = _DENSIFY_PTS):
    """Boundary samples of (n, 4) boxes the way PROJ's proj_trans_bounds walks them: every edge in
    ``densify_pts + 1`` steps (from memory -- source not in container).  Returns (n, 4 * steps)."""
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
    steps = densify_pts + 1
    k = np.arange(steps, dtype=np.float64)[None, :]
    west, south, east, north = (boxes[:, c:c + 1] for c in range(4))
    dx, dy = (east - west) / steps, (north - south) / steps
    full = np.ones_like(k)
    xs = np.concatenate([west + k * dx, east * full, east - k * dx, west * full], axis=1)
    ys = np.concatenate([south * full, south + k * dy, north * full, north - k * dy], axis=1)
    return xs, ys


def _unwrap_longitudes(lon: np.ndarray) -> np.ndarray:
    """Remove +-360 degree jumps between consecutive samples along the last axis (NaNs pass through)."""
    lon = np.asarray(lon, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        step = np.diff(lon, axis=-1)
        turn = np.where(np.abs(step) > 180.0, -np.sign(step) * 360.0, 0.0)
    shift = np.concatenate([np.zeros(lon.shape[:-1] + (1,)), np.cumsum(turn, axis=-1)], axis=-1)
    return lon + shift


def transform_bounds(from_crs, to_crs, boxes, densify_pts: int = _DENSIFY_PTS, device=None) -> np.ndarray:
    """``Transformer.transform_bounds`` for a batch of (x_min, y_min, x_max, y_max) boxes: densified
    edges through the device transform, then min / max over the finite results.  (n, 4) float64.

    Geographic output: the point transform wraps longitudes into [-180, 180] like PROJ, so a box
    that touches the antimeridian (e.g. a web-Mercator tile ending at x = +20037508.34) would get
    a boundary sample at -180 next to ones at +179.99.  Consecutive samples are never half a turn
    apart, so such jumps are undone (the walk stays contiguous: 179.99 -> 180.00001) and the box
    keeps its real extent.  PROJ's own antimeridian result (x_min > x_max) and its pole-enclosing
    case are not reproduced."""
    xs, ys = _edge_points(boxes, densify_pts)
    tx, ty = transform_points(xs, ys, from_crs, to_crs, device)
    if normalize_crs(to_crs).is_geographic:
        tx = _unwrap_longitudes(tx)
        # the walk is contiguous but may have started on the far side of the antimeridian (its first
        # sample wrapped): bring every box as a whole back so that its centre lies in [-180, 180]
        with np.errstate(invalid="ignore"):
            centre = 0.5 * (np.nanmin(np.where(np.isfinite(tx), tx, np.nan), axis=1)
                            + np.nanmax(np.where(np.isfinite(tx), tx, np.nan), axis=1))
        turns = np.where(np.isfinite(centre), np.round(centre / 360.0), 0.0)
        tx = tx - 360.0 * turns[:, None]
    ok = np.isfinite(tx) & np.isfinite(ty)
    big = np.inf
    out = np.stack([np.where(ok, tx, big).min(axis=1), np.where(ok, ty, big).min(axis=1),
                    np.where(ok, tx, -big).max(axis=1), np.where(ok, ty, -big).max(axis=1)], axis=1)
    if not np.all(np.isfinite(out)):
        raise ValueError("transform_bounds: a box has no transformable boundary point")
    return out


# ---------------------------------------------------------------------------
# host: per-tile source windows (reproject.py:385-469)
# ---------------------------------------------------------------------------
class SourceWindows:
    """Per reference tile of the target: where its source window starts, the common window size
    and the float32 coordinate of window element (0, 0)."""

    def __init__(self, i0, j0, win_w, win_h, x0, y0, pad):
        self.i0, self.j0 = i0, j0          # (nty, ntx) int64, unpadded source indices (may be < 0)
        self.win_w, self.win_h = win_w, win_h
        self.x0, self.y0 = x0, y0          # (nty, ntx) float32
        self.pad = pad                     # ((top, bottom), (left, right)) the reference would pad


def get_source_windows(source_gm: GridMapping, target_gm: GridMapping, device=None) -> SourceWindows:
    """``_get_scr_bboxes_indices`` (reproject.py:385-469)."""
    ntx = math.ceil(target_gm.width / target_gm.tile_width)
    nty = math.ceil(target_gm.height / target_gm.tile_height)
    x_axis_src, y_axis_src = source_gm.x_values, source_gm.y_values
    ox, oy = float(x_axis_src[0]), float(y_axis_src[0])
    x_res, y_res = source_gm.x_res, source_gm.y_res
    boxes = transform_bounds(target_gm.crs, source_gm.crs, target_gm.xy_bboxes, device=device)
    lo_i = np.empty((nty, ntx), dtype=np.int64)
    lo_j, hi_i, hi_j = np.empty_like(lo_i), np.empty_like(lo_i), np.empty_like(lo_i)
    for k in range(nty * ntx):
        ty, tx = divmod(k, ntx)
        bx0, by0, bx1, by1 = (float(v) for v in boxes[k])
        lo_i[ty, tx] = math.floor((bx0 - ox) / x_res)
        hi_i[ty, tx] = math.ceil((bx1 - ox) / x_res)
        lo_j[ty, tx] = math.floor((oy - by1) / y_res)
        hi_j[ty, tx] = math.ceil((oy - by0) / y_res)
    ext_i, ext_j = hi_i - lo_i, hi_j - lo_j
    win_w, win_h = int(ext_i.max()) + 1, int(ext_j.max()) + 1
    i0 = lo_i - (win_w - ext_i) // 2
    j0 = lo_j - (win_h - ext_j) // 2
    i_min, i_max = int(i0.min()), int(i0.max()) + win_w
    j_min, j_max = int(j0.min()), int(j0.max()) + win_h
    # coordinate axes over the union of all windows, then float32 per tile (reproject.py:427-450)
    x_axis = np.arange(ox + i_min * x_res, ox + i_max * x_res, x_res)
    y_step = float(y_axis_src[1] - y_axis_src[0])
    y_axis = np.arange(oy + j_min * y_step, oy + j_max * y_step, y_step)
    x0 = x_axis[i0 - i_min].astype(np.float32)
    y0 = y_axis[j0 - j_min].astype(np.float32)
    pad = ((-min(0, j_min), max(0, j_max - source_gm.height)), (-min(0, i_min), max(0, i_max - source_gm.width)))
    return SourceWindows(i0, j0, win_w, win_h, x0, y0, pad)


# ---------------------------------------------------------------------------
# device level: the fused kernel
# ---------------------------------------------------------------------------
class ReprojectPlan:
    """Device-resident tables of one (source grid, target grid) pair; :meth:`run` only enqueues."""

    def __init__(self, source_gm: GridMapping, target_gm: GridMapping, device=None,
                 rows: tuple[int, int] | None = None, windows: SourceWindows | None = None):
        self.lib = load()
        self.device = _dev.require_cuda(device)
        self.source_gm, self.target_gm = source_gm, target_gm
        self.rows = (0, target_gm.height) if rows is None else (int(rows[0]), int(rows[1]))
        self.windows = windows if windows is not None else get_source_windows(source_gm, target_gm, self.device)
        w = self.windows
        self._src_proj = XrsProj.from_crs(source_gm.crs)
        self._dst_proj = XrsProj.from_crs(target_gm.crs)
        self._dst_x = _dev.to_device(target_gm.x_values, self.device, dtype=np.float64)
        self._dst_y = _dev.to_device(target_gm.y_values, self.device, dtype=np.float64)
        self._x0 = _dev.to_device(w.x0.astype(np.float64).ravel(), self.device)
        self._y0 = _dev.to_device(w.y0.astype(np.float64).ravel(), self.device)
        self._i0 = _dev.to_device(w.i0.astype(np.int32).ravel(), self.device)
        self._j0 = _dev.to_device(w.j0.astype(np.int32).ravel(), self.device)

    def footprint(self) -> tuple[int, int, int, int] | None:
        """Source window (i0, j0, i1, j1), end-exclusive and clipped to the source, that the tiles
        intersecting ``rows`` can read; ``None`` if they lie entirely outside the source."""
        from .bands import reproject_band_footprint

        w = self.windows
        return reproject_band_footprint(w.i0, w.j0, w.win_w, w.win_h, self.target_gm, self.rows,
                                        (self.source_gm.width, self.source_gm.height))

    def run(self, src: torch.Tensor, interp_method: str, fill_value, out: torch.Tensor | None = None,
            out_dtype=None, window_origin: tuple[int, int] = (0, 0)) -> torch.Tensor:
        """Reproject (bands, h, w) or (h, w) ``src`` -- the whole source image, or the window of it
        that starts at pixel ``window_origin=(i0, j0)`` and covers :meth:`footprint`.

        ``out_dtype``: ``None`` = what the reference returns (float64 for bilinear, else the source
        dtype); pass the source dtype to get bilinear results cast once to it."""
        if interp_method not in INTERP_CODES:
            raise NotImplementedError(
                f"interp_methods must be one of 0, 1, 'nearest', 'bilinear', "
                f"'triangular', was '{interp_method}'."
            )
        squeeze = src.dim() == 2
        src3 = src.unsqueeze(0) if squeeze else src
        if src3.stride(2) != 1:
            src3 = src3.contiguous()
        np_dtype = np.dtype(str(src3.dtype).replace("torch.", ""))
        if out_dtype is None:
            out_np_dtype = np.dtype(np.float64) if interp_method == "bilinear" else np_dtype
        else:
            out_np_dtype = np.dtype(out_dtype)
        bands, win_h, win_w = src3.shape
        gm, sgm, w = self.target_gm, self.source_gm, self.windows
        n_rows = self.rows[1] - self.rows[0]
        if out is None:
            out = _dev.empty((bands, n_rows, gm.width), out_np_dtype, src3.device)
        src_planes = _dev.plane_ptr_array(src3)
        dst_planes = _dev.plane_ptr_array(out)
        check(self.lib.xrs_reproject(
            src_planes, dst_planes, bands, DTYPE_CODES[np_dtype], DTYPE_CODES[out_np_dtype], sgm.height, sgm.width,
            src3.stride(1), int(window_origin[0]), int(window_origin[1]), win_w, win_h,
            ctypes.addressof(self._src_proj), ctypes.addressof(self._dst_proj), _dev.ptr(self._dst_x),
            _dev.ptr(self._dst_y), gm.height, gm.width, gm.tile_height, gm.tile_width, _dev.ptr(self._x0),
            _dev.ptr(self._y0), _dev.ptr(self._i0), _dev.ptr(self._j0), w.win_w, w.win_h, float(sgm.x_res),
            float(sgm.y_res), INTERP_CODES[interp_method], float(fill_value), self.rows[0], self.rows[1],
            _dev.stream_ptr(src3.device)), "xrs_reproject")
        return out[0] if squeeze else out


# ---------------------------------------------------------------------------
# dataset level
# ---------------------------------------------------------------------------
def reproject_dataset(
    source_ds,
    target_gm: GridMapping,
    source_gm: GridMapping | None = None,
    variables: str | Iterable[str] | None = None,
    interp_methods=None,
    agg_methods=None,
    recover_nans=False,
    fill_values=None,
    *,
    devices: Iterable | None = None,
):
    """Reproject a dataset on a regular grid to a regular grid in another CRS.

    Drop-in for ``xcube_resampling.reproject.reproject_dataset`` (reproject.py:51-186): same
    arguments, defaults and errors; always eager (numpy in, numpy out).

    ``devices`` (keyword only, not in the reference): CUDA devices to spread the call over.  The
    target is cut into one row band per device; every device uploads only the source rectangle its
    band's tiles can read (``ReprojectPlan.footprint``) and fills its rows of the result.  No
    exchange step is involved."""
    user_ds = source_ds
    source_ds = from_any(source_ds)
    if source_gm is None:
        source_gm = GridMapping.from_dataset(source_ds)
    if source_gm.is_j_axis_up:  # reproject.py:115-118
        source_ds = _flip_y(source_ds, source_gm.xy_dim_names[1])
        source_gm = GridMapping.from_dataset(source_ds, crs=source_gm.crs)
    source_ds = normalize_grid_mapping(source_ds, source_gm)
    source_ds = _select_variables(source_ds, variables)

    source_ds, source_gm = _downscale_source_dataset(source_ds, source_gm, target_gm, interp_methods, agg_methods,
                                                     recover_nans)

    # output coordinates (reproject.py:152-159)
    sx_name, sy_name = source_gm.xy_var_names
    coords = {n: v for n, v in source_ds.coords.items() if n not in (sx_name, sy_name)}
    tx_name, ty_name = target_gm.xy_var_names
    coords[tx_name] = target_gm.x_coords
    coords[ty_name] = target_gm.y_coords
    coords["spatial_ref"] = DataArray(np.array(0), dims=(), attrs=target_gm.crs.to_cf())
    target_ds = Dataset(coords=coords, attrs=source_ds.attrs)

    from ._pipeline import Target, group_by_buffer

    yx_dims = (source_gm.xy_dim_names[1], source_gm.xy_dim_names[0])
    t_dims = (target_gm.xy_dim_names[1], target_gm.xy_dim_names[0])
    H, W = target_gm.height, target_gm.width
    items, results = [], []
    for var_name, var in source_ds.items():
        if var.dims[-2:] == yx_dims:
            assert len(var.dims) in (2, 3), f"Data variable {var_name} has {len(var.dims)} dimensions."
            fill_value = _get_fill_value(fill_values, var_name, var)
            interp_method = _get_interp_method_str(interp_methods, var_name, var)
            if interp_method not in INTERP_CODES:
                raise NotImplementedError(
                    f"interp_methods must be one of 0, 1, 'nearest', 'bilinear', "
                    f"'triangular', was '{interp_method}'."
                )
            values = getattr(var, "source", None) or var.values  # io.LazyDataArray: streamed from its store
            n_b = 1 if values.ndim == 2 else values.shape[0]
            # numpy's promotion in reproject.py:315-328 makes bilinear results float64
            out_dtype = np.float64 if interp_method == "bilinear" else values.dtype
            out = _dev.pinned_empty((n_b, H, W), out_dtype)
            items.append((values, Target(str(var_name), interp_method, fill_value, out)))
            dims = t_dims if len(var.dims) == 2 else (var.dims[0],) + t_dims
            results.append((var_name, out[0] if values.ndim == 2 else out, dims, var.attrs))
        elif yx_dims[0] not in var.dims and yx_dims[1] not in var.dims:
            results.append((var_name, var, None, None))
    groups = group_by_buffer(items)
    if groups:
        devs = [None] if devices is None else list(devices)
        reproject_groups(groups, source_gm, target_gm, devs)
    for var_name, out, dims, attrs in results:
        target_ds[var_name] = out if dims is None else DataArray(out, dims=dims, attrs=attrs, name=var_name)
    return to_like(target_ds, user_ds)


def reproject_groups(groups, source_gm: GridMapping, target_gm: GridMapping, devices, band_edges=None,
                     windows: SourceWindows | None = None, chunk_bands: int = 4, stats: list | None = None) -> None:
    """Fill the targets of ``groups`` (``_pipeline.SourceGroup`` list) on ``devices``: device k
    computes target rows ``band_edges[k]:band_edges[k+1]`` (default: equal heights on multiples of 32
    rows) from the source rectangle its tiles can read.  One device = the whole image."""
    from . import multigpu
    from ._pipeline import GatherPipeline

    n = len(devices)
    first = _dev.require_cuda(devices[0])
    if windows is None:
        windows = get_source_windows(source_gm, target_gm, first)
    edges = list(band_edges) if band_edges is not None else multigpu.default_band_edges(target_gm.height, n)
    h, w = source_gm.height, source_gm.width

    def worker(k, dev, _exchange):
        rows = (int(edges[k]), int(edges[k + 1]))
        if rows[1] <= rows[0]:
            return
        dev = _dev.require_cuda(dev)
        with torch.cuda.device(dev):
            plan = ReprojectPlan(source_gm, target_gm, dev, rows=rows, windows=windows)
            fp = plan.footprint()
            if fp is None:  # the band lies entirely outside the source
                for grp in groups:
                    for tgt in grp.targets:
                        tgt.out_host[:, rows[0] - tgt.row0:rows[1] - tgt.row0, :] = \
                            np.asarray(tgt.fill).astype(tgt.out_dtype)
                return
            i_lo, j_lo, i_hi, j_hi = fp
            a = 32  # column range on 128-byte boundaries of 4-byte data
            i_lo, i_hi = (i_lo // a) * a, min(w, -(-i_hi // a) * a)
            pipe = GatherPipeline(dev, (h, w), target_gm.width, rows, src_window=(j_lo, j_hi),
                                  segments=[(j_lo, j_hi, i_lo, i_hi)], chunk_bands=chunk_bands)

            def process(src_view, tgt, out_view, b0):
                plan.run(src_view, tgt.method, tgt.fill, out=out_view, out_dtype=tgt.out_dtype, window_origin=(0, j_lo))

            pipe.run(groups, process)
            if stats is not None:
                stats.append({"device": str(dev), "rows": rows, "footprint": (i_lo, j_lo, i_hi, j_hi),
                              "h2d_bytes": pipe.h2d_bytes, "d2h_bytes": pipe.d2h_bytes})

    multigpu.run_on_devices([(_dev.require_cuda(d)) for d in devices], worker)


def _flip_y(ds: Dataset, y_dim: str) -> Dataset:
    """``ds.isel({y_dim: slice(None, None, -1)})``."""

    def flip(var: DataArray) -> DataArray:
        if y_dim not in var.dims:
            return var
        idx = tuple(slice(None, None, -1) if d == y_dim else slice(None) for d in var.dims)
        return DataArray(var.values[idx], dims=var.dims, attrs=var.attrs, name=var.name)

    return Dataset(data_vars={n: flip(v) for n, v in ds.items()},
                   coords={n: flip(v) for n, v in ds.coords.items()}, attrs=ds.attrs)


def _downscale_source_dataset(source_ds: Dataset, source_gm: GridMapping, target_gm: GridMapping, interp_methods,
                              agg_methods, recover_nans):
    """reproject.py:338-382: when the source is finer than the target (ratio < 0.95 in x or y) clip
    it to the target's footprint and resample it to the target's resolution first."""
    from .affine import affine_transform_dataset

    box = transform_bounds(target_gm.crs, source_gm.crs, [target_gm.xy_bbox])[0]
    xres_trans = (box[2] - box[0]) / target_gm.width
    yres_trans = (box[3] - box[1]) / target_gm.height
    x_scale = source_gm.x_res / xres_trans
    y_scale = source_gm.y_res / yres_trans
    if x_scale < SCALE_LIMIT or y_scale < SCALE_LIMIT:
        box = (box[0] - 2 * source_gm.x_res, box[1] - 2 * source_gm.y_res,
               box[2] + 2 * source_gm.x_res, box[3] + 2 * source_gm.y_res)
        source_ds = clip_dataset_by_bbox(source_ds, box, source_gm.xy_dim_names)
        source_gm = GridMapping.from_dataset(source_ds, crs=source_gm.crs)
        w, h = round(x_scale * source_gm.width), round(y_scale * source_gm.height)
        size = (w if w >= 2 else 2, h if h >= 2 else 2)
        downscale_target_gm = GridMapping.regular(size=size, xy_min=(source_gm.xy_bbox[0], source_gm.xy_bbox[1]),
                                                  xy_res=(float(xres_trans), float(yres_trans)), crs=source_gm.crs,
                                                  tile_size=source_gm.tile_size)
        source_ds = from_any(affine_transform_dataset(
            source_ds, downscale_target_gm, source_gm=source_gm,
            interp_methods=_prep_interp_methods_downscale(interp_methods), agg_methods=agg_methods,
            recover_nans=recover_nans))
        source_gm = GridMapping.from_dataset(source_ds, crs=source_gm.crs)
    return source_ds, source_gm

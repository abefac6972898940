"""rectify_dataset: irregular (2-D coordinate) source -> regular target grid.

Same entry point, arguments and error behaviour as the reference's
``xcube_resampling/rectify.py:54-179``.  What changes is underneath: instead of
one dask task per target tile the whole scene stays resident in HBM and three
kernels of ``libxrs.so`` do the work --

* K0 ``xrs_tile_src_bboxes``  per-tile source windows (``gridmapping/bboxes.py:28-106``)
* K1 ``xrs_rectify_ij``       source-index image (``rectify.py:312-576``)
* K2 ``xrs_gather_ij``        gather of all bands (``rectify.py:579-734``)

The device-level functions (``compute_target_source_ij``, ``gather_ij``) take
and return torch tensors used purely as device buffers.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np
import torch

from . import _dev
from ._lib import check, load
from .constants import DTYPE_CODES, INTERP_CODES, SCALE_LIMIT, UV_DELTA
from .dataset import DataArray, Dataset, from_any, to_like
from .gridmapping import GridMapping
from .utils import (
    _get_fill_value,
    _get_interp_method_str,
    _is_equal_crs,
    _prep_interp_methods_downscale,
    _select_variables,
    normalize_grid_mapping,
)


# ---------------------------------------------------------------------------
# device level
# ---------------------------------------------------------------------------
def _xy_border(target_gm: GridMapping) -> float:
    """rectify.py:329-340: empirical border around each tile box."""
    num_tiles_x = target_gm.width / target_gm.tile_width
    num_tiles_y = target_gm.height / target_gm.tile_height
    x_min, y_min, x_max, y_max = target_gm.xy_bbox
    return min(
        min(2 * num_tiles_x * target_gm.x_res, 2 * num_tiles_y * target_gm.y_res),
        min(0.5 * (x_max - x_min), 0.5 * (y_max - y_min)),
    )


def _separable_axes(xy_bboxes: np.ndarray, xy_border: float):
    """Split row-major tile boxes into per-column x and per-row y intervals (grown by the border)."""
    xy_bboxes = np.asarray(xy_bboxes, dtype=np.float64)
    n = xy_bboxes.shape[0]
    # number of tile columns = length of the first run with a constant y interval
    ntx = 1
    while ntx < n and xy_bboxes[ntx, 1] == xy_bboxes[0, 1] and xy_bboxes[ntx, 3] == xy_bboxes[0, 3]:
        ntx += 1
    if n % ntx:
        raise NotImplementedError("xy_bboxes must form a row-major grid of tiles")
    nty = n // ntx
    grid = xy_bboxes.reshape(nty, ntx, 4)
    if not (np.all(grid[:, :, 0] == grid[0:1, :, 0]) and np.all(grid[:, :, 2] == grid[0:1, :, 2])
            and np.all(grid[:, :, 1] == grid[:, 0:1, 1]) and np.all(grid[:, :, 3] == grid[:, 0:1, 3])):
        raise NotImplementedError("xy_bboxes must form a separable (regular) grid of tiles")
    # same doubles as bboxes.py:60-63
    x_lo, x_hi = grid[0, :, 0] - xy_border, grid[0, :, 2] + xy_border
    y_lo, y_hi = grid[:, 0, 1] - xy_border, grid[:, 0, 3] + xy_border
    return x_lo, x_hi, y_lo, y_hi, ntx, nty


class RectifyPlan:
    """Device-resident constants and scratch buffers of one (source shape, target grid) pair.

    Building the plan uploads the tile axis tables once and allocates the K0/K1 workspaces and
    the ij buffer; afterwards :meth:`windows` and :meth:`ij` only enqueue kernels (no allocation,
    no host<->device traffic, no synchronisation), which is what the steady state of a
    scene-after-scene service looks like.
    """

    def __init__(self, target_gm: GridMapping, device=None, rows: tuple[int, int] | None = None,
                 uv_delta: float = UV_DELTA, ij_border: int = 1):
        lib = load()
        self.lib = lib
        self.gm = target_gm
        self.device = _dev.require_cuda(device)
        self.uv_delta = float(uv_delta)
        self.ij_border = int(ij_border)
        H, W = target_gm.height, target_gm.width
        self.rows = (0, H) if rows is None else (int(rows[0]), int(rows[1]))
        x_lo, x_hi, y_lo, y_hi, self.ntx, self.nty = _separable_axes(target_gm.xy_bboxes, _xy_border(target_gm))
        self._axes = _dev.to_device(np.concatenate([x_lo, x_hi, y_lo, y_hi]), self.device)
        self.tile_boxes = _dev.empty((self.ntx * self.nty, 4), np.int64, self.device)
        self._ws0 = _dev.workspace(lib.xrs_tile_src_bboxes_workspace_bytes(self.ntx, self.nty), self.device)
        self._ws1 = None  # K1 workspace, sized on first use (depends on the source shape)
        self.ij_buf = _dev.empty((2, self.rows[1] - self.rows[0], W), np.float64, self.device)

    def windows(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """K0: per-reference-tile source windows, (n_tiles, 4) int64 on the device."""
        _check_coords(x, y)
        h, w = x.shape
        base, ntx, nty = self._axes.data_ptr(), self.ntx, self.nty
        check(self.lib.xrs_tile_src_bboxes(
            _dev.ptr(x), _dev.ptr(y), h, w, x.stride(0), base, base + 8 * ntx, ntx, base + 16 * ntx,
            base + 16 * ntx + 8 * nty, nty, self.ij_border, _dev.ptr(self.tile_boxes), _dev.ptr(self._ws0),
            _dev.stream_ptr(self.device)), "xrs_tile_src_bboxes")
        return self.tile_boxes

    def ij(self, x: torch.Tensor, y: torch.Tensor, tile_boxes: torch.Tensor | None = None) -> torch.Tensor:
        """K0 (unless ``tile_boxes`` is given) + K1: the source-index image of ``rows``."""
        _check_coords(x, y)
        if tile_boxes is None:
            tile_boxes = self.windows(x, y)
        gm = self.gm
        x_min, y_min, x_max, y_max = gm.xy_bbox
        h, w = x.shape
        need = self.lib.xrs_rectify_ij_workspace_bytes(h, w, self.rows[1] - self.rows[0], gm.width)
        if self._ws1 is None or self._ws1.numel() < need:
            self._ws1 = _dev.workspace(need, self.device)
        check(self.lib.xrs_rectify_ij(
            _dev.ptr(x), _dev.ptr(y), h, w, x.stride(0), _dev.ptr(tile_boxes), _dev.ptr(self.ij_buf), gm.height,
            gm.width, gm.tile_height, gm.tile_width, float(x_min), float(y_min), float(y_max), float(gm.x_res),
            float(gm.y_res), int(bool(gm.is_j_axis_up)), self.uv_delta, self.rows[0], self.rows[1],
            _dev.ptr(self._ws1), _dev.stream_ptr(self.device)), "xrs_rectify_ij")
        return self.ij_buf

    def rectify_gather(self, x: torch.Tensor, y: torch.Tensor, src: torch.Tensor, interp_method: str, fill_value,
                       out: torch.Tensor | None = None, tile_boxes: torch.Tensor | None = None,
                       window_origin: tuple[int, int] = (0, 0), full_size: tuple[int, int] | None = None):
        """Fused K1 + K2 for ONE variable (see :func:`_rectify_gather_dev`)."""
        return _rectify_gather_dev(self, x, y, src, interp_method, fill_value, out, tile_boxes, window_origin, full_size)


def _rectify_gather_dev(plan: "RectifyPlan", x: torch.Tensor, y: torch.Tensor, src: torch.Tensor, interp_method: str,
                        fill_value, out: torch.Tensor | None = None, tile_boxes: torch.Tensor | None = None,
                        window_origin: tuple[int, int] = (0, 0), full_size: tuple[int, int] | None = None):
    """K0 (unless ``tile_boxes`` is given) + K1 claim stage + fused resolve / gather
    (``xrs_rectify_gather``): the ij image is never written.  Same results as
    ``gather_ij(src, plan.ij(x, y), ...)``; for calls that gather a single variable."""
    lib = plan.lib
    _check_coords(x, y)
    if interp_method not in INTERP_CODES:
        raise NotImplementedError(
            f"interp_methods must be one of 0, 1, 'nearest', 'bilinear', "
            f"'triangular', was '{interp_method}'."
        )
    if tile_boxes is None:
        tile_boxes = plan.windows(x, y)
    gm = plan.gm
    x_min, y_min, x_max, y_max = gm.xy_bbox
    h, w = x.shape
    n_rows = plan.rows[1] - plan.rows[0]
    need = lib.xrs_rectify_ij_workspace_bytes(h, w, n_rows, gm.width)
    if plan._ws1 is None or plan._ws1.numel() < need:
        plan._ws1 = _dev.workspace(need, plan.device)
    squeeze = src.dim() == 2
    src3 = src.unsqueeze(0) if squeeze else src
    if src3.stride(2) != 1 or (src3.shape[0] > 1 and src3.stride(0) < src3.stride(1) * src3.shape[1]):
        src3 = src3.contiguous()
    np_dtype = np.dtype(str(src3.dtype).replace("torch.", ""))
    bands, win_h, win_w = src3.shape
    full_w, full_h = (win_w, win_h) if full_size is None else (int(full_size[0]), int(full_size[1]))
    if (full_w, full_h) != (w, h):
        raise ValueError("data variable and coordinates must describe the same source image")
    if out is None:
        out = torch.empty((bands, n_rows, gm.width), dtype=src3.dtype, device=src3.device)
    check(lib.xrs_rectify_gather(
        _dev.ptr(x), _dev.ptr(y), h, w, x.stride(0), _dev.ptr(tile_boxes), gm.height, gm.width, gm.tile_height,
        gm.tile_width, float(x_min), float(y_min), float(y_max), float(gm.x_res), float(gm.y_res),
        int(bool(gm.is_j_axis_up)), plan.uv_delta, plan.rows[0], plan.rows[1], _dev.ptr(plan._ws1),
        _dev.plane_ptr_array(src3), _dev.plane_ptr_array(out), bands, DTYPE_CODES[np_dtype], src3.stride(1),
        int(window_origin[0]), int(window_origin[1]), win_w, win_h, INTERP_CODES[interp_method], float(fill_value),
        _dev.stream_ptr(plan.device)), "xrs_rectify_gather")
    return out[0] if squeeze else out


def _check_coords(x: torch.Tensor, y: torch.Tensor):
    if x.dtype != torch.float64 or y.dtype != torch.float64:
        raise TypeError("source coordinates must be float64 device tensors")
    if x.dim() != 2 or x.shape != y.shape or x.stride() != y.stride() or x.stride(1) != 1:
        raise ValueError("x and y must be 2-D with the same shape and row-major layout")


def tile_source_windows_dev(x: torch.Tensor, y: torch.Tensor, xy_bboxes: np.ndarray, xy_border: float,
                            ij_border: int) -> torch.Tensor:
    """K0 on device coordinates for arbitrary separable boxes; (n_tiles, 4) int64 on the device."""
    lib = load()
    _check_coords(x, y)
    x_lo, x_hi, y_lo, y_hi, ntx, nty = _separable_axes(xy_bboxes, xy_border)
    dev = x.device
    axes = _dev.to_device(np.concatenate([x_lo, x_hi, y_lo, y_hi]), dev)
    out = _dev.empty((ntx * nty, 4), np.int64, dev)
    ws = _dev.workspace(lib.xrs_tile_src_bboxes_workspace_bytes(ntx, nty), dev)
    h, w = x.shape
    base = axes.data_ptr()
    check(lib.xrs_tile_src_bboxes(
        _dev.ptr(x), _dev.ptr(y), h, w, x.stride(0),
        base, base + 8 * ntx, ntx, base + 16 * ntx, base + 16 * ntx + 8 * nty, nty,
        int(ij_border), _dev.ptr(out), _dev.ptr(ws), _dev.stream_ptr(dev)), "xrs_tile_src_bboxes")
    return out


def tile_source_windows_for_boxes(source_gm: GridMapping, xy_bboxes: np.ndarray, xy_border: float,
                                  ij_border: int) -> np.ndarray:
    """``GridMapping.ij_bboxes_from_xy_bboxes`` (base.py:565-629) through K0."""
    xy = source_gm.xy_coords.values
    x = _dev.to_device(xy[0], dtype=np.float64)
    y = _dev.to_device(xy[1], dtype=np.float64)
    xy_bboxes = np.asarray(xy_bboxes, dtype=np.float64).reshape(-1, 4)
    try:
        return _dev.to_host(tile_source_windows_dev(x, y, xy_bboxes, xy_border, ij_border))
    except NotImplementedError:
        # arbitrary boxes (not the row-major tiles of one grid): one pass per box over the resident coordinates
        parts = [tile_source_windows_dev(x, y, xy_bboxes[k:k + 1], xy_border, ij_border) for k in range(len(xy_bboxes))]
        return _dev.to_host(torch.cat(parts, dim=0))


def compute_target_source_ij(x: torch.Tensor, y: torch.Tensor, target_gm: GridMapping,
                             uv_delta: float = UV_DELTA, tile_boxes: torch.Tensor | None = None,
                             rows: tuple[int, int] | None = None) -> torch.Tensor:
    """``_compute_target_source_ij`` (rectify.py:312-370) on device buffers.

    x, y: (h, w) float64 device tensors with the source coordinates in the target
    CRS.  Returns the (2, H, W) float64 source-index image on the device, or only
    its target rows ``rows=(begin, end)`` (one row band of a multi-GPU split).
    """
    return RectifyPlan(target_gm, x.device, rows=rows, uv_delta=uv_delta).ij(x, y, tile_boxes)


def gather_ij(src: torch.Tensor, ij: torch.Tensor, interp_method: str, fill_value,
              out: torch.Tensor | None = None, window_origin: tuple[int, int] = (0, 0),
              full_size: tuple[int, int] | None = None) -> torch.Tensor:
    """``_compute_var_image`` (rectify.py:579-734) on device buffers.

    src: (bands, h, w) or (h, w) device tensor; ij: (2, H, W) float64.  When only a
    window of the source is resident (row-band footprint) ``window_origin=(i0, j0)``
    is its position in the full image of size ``full_size=(width, height)``.
    """
    lib = load()
    if interp_method not in INTERP_CODES:
        raise NotImplementedError(
            f"interp_methods must be one of 0, 1, 'nearest', 'bilinear', "
            f"'triangular', was '{interp_method}'."
        )
    squeeze = src.dim() == 2
    src3 = src.unsqueeze(0) if squeeze else src
    if src3.stride(2) != 1 or (src3.shape[0] > 1 and src3.stride(0) < src3.stride(1) * src3.shape[1]):
        src3 = src3.contiguous()
    np_dtype = np.dtype(str(src3.dtype).replace("torch.", ""))
    bands, win_h, win_w = src3.shape
    w, h = (win_w, win_h) if full_size is None else (int(full_size[0]), int(full_size[1]))
    _, H, W = ij.shape
    if out is None:
        out = torch.empty((bands, H, W), dtype=src3.dtype, device=src3.device)
    src_planes = _dev.plane_ptr_array(src3)
    dst_planes = _dev.plane_ptr_array(out)
    fill = float(fill_value)
    check(lib.xrs_gather_ij(src_planes, dst_planes, bands, DTYPE_CODES[np_dtype], h, w, src3.stride(1),
                            int(window_origin[0]), int(window_origin[1]), win_w, win_h, _dev.ptr(ij), H, W,
                            INTERP_CODES[interp_method], fill, _dev.stream_ptr(src3.device)),
          "xrs_gather_ij")
    return out[0] if squeeze else out


def rectify_band_host(x: np.ndarray, y: np.ndarray, src_window: np.ndarray, window_origin: tuple[int, int],
                      full_size: tuple[int, int], target_gm: GridMapping, rows: tuple[int, int], interp_method: str,
                      fill_value, device=None) -> np.ndarray:
    """Rectify one target row band from host buffers to a host buffer (multi-GPU building block).

    x, y: full (h, w) source coordinates (needed to find the band's source windows);
    src_window: (bands, wh, ww) window of the data variable whose element (0, 0, 0) is pixel
    ``window_origin=(i0, j0)`` of the full ``full_size=(width, height)`` image and which covers
    :func:`xcube_resampling_b200.bands.rectify_band_footprint` of ``rows``.
    Returns (bands, rows[1]-rows[0], target width) in (pinned) host memory.
    """
    dev = _dev.require_cuda(device)
    x_dev = _dev.to_device(x, dev, dtype=np.float64)
    y_dev = _dev.to_device(y, dev, dtype=np.float64)
    plan = RectifyPlan(target_gm, dev, rows=rows)
    ij = plan.ij(x_dev, y_dev)
    return _gather_from_host(np.asarray(src_window), ij, interp_method, fill_value, window_origin, full_size)


# ---------------------------------------------------------------------------
# dataset level
# ---------------------------------------------------------------------------
def rectify_dataset(
    source_ds,
    target_gm: GridMapping | None = None,
    source_gm: GridMapping | None = None,
    variables: str | Iterable[str] | None = None,
    interp_methods=None,
    agg_methods=None,
    recover_nans=False,
    fill_values=None,
    tile_size: int | tuple[int, int] | None = None,
):
    """Rectify a dataset with 2-D (irregular) coordinates to a regular grid.

    Drop-in for ``xcube_resampling.rectify.rectify_dataset`` (rectify.py:54-179):
    same arguments, defaults and errors; always eager (numpy in, numpy out).
    """
    user_ds = source_ds
    source_ds = from_any(source_ds)
    if source_gm is None:
        source_gm = GridMapping.from_dataset(source_ds)
    source_ds = normalize_grid_mapping(source_ds, source_gm)

    if target_gm is None:
        target_gm = source_gm.to_regular(tile_size=tile_size)

    # source coordinates in the target CRS (rectify.py:126-129)
    src_x, src_y = source_gm.x_values, source_gm.y_values
    if src_x.ndim == 1:
        xy = source_gm.xy_coords.values
        src_x, src_y = xy[0], xy[1]
    x_dev = _dev.to_device(src_x, dtype=np.float64)
    y_dev = _dev.to_device(src_y, dtype=np.float64)
    if not _is_equal_crs(source_gm, target_gm):
        from .reproject import transform_points_dev

        x_dev, y_dev = transform_points_dev(x_dev, y_dev, source_gm.crs, target_gm.crs)
        source_gm = _gm_from_transformed(source_gm, target_gm, x_dev, y_dev)

    source_ds = _select_variables(source_ds, variables)

    # pre-downscale when the source is finer than the target (rectify.py:135-143)
    x_scale = source_gm.x_res / target_gm.x_res
    y_scale = source_gm.y_res / target_gm.y_res
    if x_scale < SCALE_LIMIT or y_scale < SCALE_LIMIT:
        source_ds, source_gm, x_dev, y_dev = _downscale_source(
            source_ds, source_gm, x_dev, y_dev, x_scale, y_scale,
            _prep_interp_methods_downscale(interp_methods), agg_methods, recover_nans)

    plan = RectifyPlan(target_gm, x_dev.device, uv_delta=UV_DELTA)
    ij = None  # computed on first need; a single small variable takes the fused K1 + K2 path instead

    # output coordinates (rectify.py:148-157)
    sx_name, sy_name = source_gm.xy_var_names
    coords = {n: v for n, v in source_ds.coords.items() if n not in (sx_name, sy_name)}
    tx_name, ty_name = target_gm.xy_var_names
    target_coords = target_gm.to_coords()
    coords[tx_name] = target_coords[tx_name]
    coords[ty_name] = target_coords[ty_name]
    coords["spatial_ref"] = DataArray(np.array(0), dims=(), attrs=target_gm.crs.to_cf())
    target_ds = Dataset(coords=coords, attrs=source_ds.attrs)

    yx_dims = (source_gm.xy_dim_names[1], source_gm.xy_dim_names[0])
    t_dims = (target_gm.xy_dim_names[1], target_gm.xy_dim_names[0])
    n_spatial = sum(1 for _n, v in source_ds.items() if v.dims[-2:] == yx_dims)
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
            if n_spatial == 1 and var.values.nbytes < _PIPELINE_MIN_BYTES:
                # one variable, small enough not to be streamed: ij is resolved in registers
                src = _dev.to_device_pitched(var.values, x_dev.device)
                out = _dev.to_host(plan.rectify_gather(x_dev, y_dev, src, interp_method, fill_value))
            else:
                if ij is None:
                    ij = plan.ij(x_dev, y_dev)
                out = _gather_from_host(var.values, ij, interp_method, fill_value)
            dims = t_dims if len(var.dims) == 2 else (var.dims[0],) + t_dims
            target_ds[var_name] = DataArray(out, dims=dims, attrs=var.attrs, name=var_name)
        elif yx_dims[0] not in var.dims and yx_dims[1] not in var.dims:
            target_ds[var_name] = var
    return to_like(target_ds, user_ds)


# a variable at least this large is streamed through the device in band chunks (upload, kernels
# and download overlapped); smaller ones take the plain upload / gather / download sequence
_PIPELINE_MIN_BYTES = 64 << 20


def _gather_from_host(values: np.ndarray, ij: torch.Tensor, interp_method: str, fill_value,
                      window_origin: tuple[int, int] = (0, 0), full_size: tuple[int, int] | None = None) -> np.ndarray:
    """K2 for one host variable ((y, x) or (bands, y, x)), result in (pinned) host memory."""
    if values.ndim == 3 and values.shape[0] > 1 and values.nbytes >= _PIPELINE_MIN_BYTES:
        pipe = _dev.BandPipeline(values, ij.shape[1:], values.dtype, ij.device)
        return pipe.run(lambda src, out: gather_ij(src, ij, interp_method, fill_value, out=out,
                                                   window_origin=window_origin, full_size=full_size))
    src = _dev.to_device_pitched(values, ij.device)
    return _dev.to_host(gather_ij(src, ij, interp_method, fill_value, window_origin=window_origin,
                                  full_size=full_size))


def _gm_from_transformed(source_gm: GridMapping, target_gm: GridMapping, x_dev, y_dev) -> GridMapping:
    """rectify.py:129: re-derive the source grid mapping from the transformed 2-D coordinates."""
    names = ("lon", "lat") if target_gm.crs.is_geographic else ("transformed_x", "transformed_y")
    dims = (source_gm.xy_dim_names[1], source_gm.xy_dim_names[0])
    xs, ys = _dev.to_host(x_dev), _dev.to_host(y_dev)
    return GridMapping.from_coords(DataArray(xs, dims=dims, name=names[0]), DataArray(ys, dims=dims, name=names[1]),
                                   target_gm.crs, tile_size=source_gm.tile_size)


def _downscale_source(source_ds, source_gm, x_dev, y_dev, x_scale, y_scale, interp_methods, agg_methods,
                      recover_nans):
    """rectify.py:234-260: affine pre-downscale of every yx variable including the 2-D coordinates."""
    from .affine import resample_dataset

    w, h = round(x_scale * source_gm.width), round(y_scale * source_gm.height)
    size = (w if w >= 2 else 2, h if h >= 2 else 2)
    yx_dims = (source_gm.xy_dim_names[1], source_gm.xy_dim_names[0])
    x_name, y_name = source_gm.xy_var_names
    ds = source_ds.assign_coords({
        x_name: DataArray(_dev.to_host(x_dev), dims=yx_dims, name=x_name),
        y_name: DataArray(_dev.to_host(y_dev), dims=yx_dims, name=y_name),
    })
    ds = resample_dataset(ds, ((1 / x_scale, 0, 0), (0, 1 / y_scale, 0)), yx_dims, size, source_gm.tile_size,
                          interp_methods, agg_methods, recover_nans)
    gm = GridMapping.from_coords(ds[x_name], ds[y_name], source_gm.crs)
    return ds, gm, _dev.to_device(gm.x_values, dtype=np.float64), _dev.to_device(gm.y_values, dtype=np.float64)

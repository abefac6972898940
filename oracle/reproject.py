"""CPU restatement of the reference's reprojection path (test infrastructure only).

Follows ``xcube_resampling/reproject.py`` of the reference:

* :func:`source_windows`        ``_get_scr_bboxes_indices``      reproject.py:385-469
* :func:`transform_gridpoints`  ``_transform_gridpoints``        reproject.py:472-496
* :func:`sample_window`         ``_reproject_block``             reproject.py:268-335
* :func:`reproject`             ``_reproject_data_array`` + ``_reorganize_data_array_slice``
                                                                 reproject.py:189-265, 499-530

The CRS transform itself is ``oracle.proj`` (PROJ is absent, see there).  ``sample_window`` is
pinned bit-for-bit against the reference's own ``_reproject_block`` (``tests/golden/reproject.npz``,
made by ``tests/golden/make_golden_reproject.py``); the whole chain is pinned against the expected
arrays of ``tests/test_reproject.py`` (see ``tests/test_oracle_golden.py``).

One deliberate difference: the reference narrows window indices to int16
(reproject.py:282-283, 286-289, 316-319); indices stay int64 here, identical whenever a tile's
source window is smaller than 32768 pixels per axis.
"""

import math

import numpy as np

from . import grid as ogrid
from . import proj as oproj


def source_windows(src_x0, src_y0, src_x_res, src_y_res, src_y_step, src_w, src_h, tgt: ogrid.RegularGrid,
                   tgt_proj: oproj.Proj, src_proj: oproj.Proj, bounds_fn=None):
    """Per target tile: source window start (unpadded source indices), common window size, the
    float32 coordinate of window element 0, and the padding the reference applies.

    src_x0, src_y0: centre of source pixel (0, 0); src_y_step: y_coords[1] - y_coords[0] (signed).
    """
    bounds_fn = bounds_fn or (lambda box: oproj.transform_bounds(tgt_proj, src_proj, *box))
    nty, ntx = tgt.n_tiles
    lo_i = np.empty((nty, ntx), dtype=np.int64)
    lo_j = np.empty_like(lo_i)
    hi_i = np.empty_like(lo_i)
    hi_j = np.empty_like(lo_i)
    for k, box in enumerate(ogrid.tile_xy_bboxes(tgt)):  # reproject.py:395-403
        ty, tx = divmod(k, ntx)
        bx0, by0, bx1, by1 = bounds_fn(tuple(float(v) for v in box))
        lo_i[ty, tx] = math.floor((bx0 - src_x0) / src_x_res)
        hi_i[ty, tx] = math.ceil((bx1 - src_x0) / src_x_res)
        lo_j[ty, tx] = math.floor((src_y0 - by1) / src_y_res)
        hi_j[ty, tx] = math.ceil((src_y0 - by0) / src_y_res)
    # reproject.py:407-423: every tile gets the largest extent + 1, centred on its own box
    ext_i, ext_j = hi_i - lo_i, hi_j - lo_j
    win_w, win_h = int(ext_i.max()) + 1, int(ext_j.max()) + 1
    i0 = lo_i - (win_w - ext_i) // 2
    j0 = lo_j - (win_h - ext_j) // 2
    # reproject.py:427-450: coordinate axes spanning all windows, float32 per-tile copies
    i_min, i_max = int(i0.min()), int((i0 + win_w).max())
    j_min, j_max = int(j0.min()), int((j0 + win_h).max())
    x_axis = np.arange(src_x0 + i_min * src_x_res, src_x0 + i_max * src_x_res, src_x_res)
    y_axis = np.arange(src_y0 + j_min * src_y_step, src_y0 + j_max * src_y_step, src_y_step)
    x0 = x_axis[i0 - i_min].astype(np.float32)
    y0 = y_axis[j0 - j_min].astype(np.float32)
    pad = ((-min(0, j_min), max(0, j_max - src_h)), (-min(0, i_min), max(0, i_max - src_w)))  # :455-465
    return dict(i0=i0, j0=j0, win_w=win_w, win_h=win_h, x0=x0, y0=y0, pad=pad)


def transform_gridpoints(tgt: ogrid.RegularGrid, tgt_proj: oproj.Proj, src_proj: oproj.Proj):
    """Target pixel centres in source CRS coordinates, two (H, W) float64 images."""
    xx, yy = np.meshgrid(ogrid.x_centres(tgt), ogrid.y_centres(tgt))
    return oproj.transform(tgt_proj, src_proj, xx, yy)


def sample_window(xx, yy, window, x0, y0, x_res, y_res, method):
    """One tile: sample ``window`` (bands, wh, ww) at source coordinates xx, yy (th, tw).

    x0, y0: float32 coordinate of window element (0, 0).  Returns what ``_reproject_block`` returns,
    including its dtype quirks (bilinear promotes to float64, triangular casts back)."""
    fx = (xx - np.float32(x0)) / x_res
    fy = (yy - np.float32(y0)) / -y_res
    if method == "nearest":
        return window[:, np.rint(fy).astype(np.int64), np.rint(fx).astype(np.int64)]
    if method not in ("bilinear", "triangular"):
        raise NotImplementedError(
            f"interp_methods must be one of 0, 1, 'nearest', 'bilinear', 'triangular', was '{method}'."
        )
    i_lo, i_hi = np.floor(fx).astype(np.int64), np.ceil(fx).astype(np.int64)
    j_lo, j_hi = np.floor(fy).astype(np.int64), np.ceil(fy).astype(np.int64)
    u = fx - i_lo
    v = fy - j_lo
    p00, p01 = window[:, j_lo, i_lo], window[:, j_lo, i_hi]
    p10, p11 = window[:, j_hi, i_lo], window[:, j_hi, i_hi]
    with np.errstate(over="ignore", invalid="ignore"):
        if method == "bilinear":
            top = p00 + u * (p01 - p00)
            bottom = p10 + u * (p11 - p10)
            return top + v * (bottom - top)
        near = p00 + u * (p01 - p00) + v * (p10 - p00)
        far = p11 + (1.0 - u) * (p10 - p11) + (1.0 - v) * (p01 - p11)
        return np.where((u + v < 1.0)[None], near, far).astype(window.dtype)


def reproject(src, src_x0, src_y0, src_x_res, src_y_res, src_y_step, tgt: ogrid.RegularGrid, tgt_proj, src_proj, method,
              fill, xx=None, yy=None):
    """Full path for one (bands, h, w) variable on a j-axis-down source; returns (bands, H, W)."""
    src = np.asarray(src)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[None]
    _, h, w = src.shape
    win = source_windows(src_x0, src_y0, src_x_res, src_y_res, src_y_step, w, h, tgt, tgt_proj, src_proj)
    if xx is None:
        xx, yy = transform_gridpoints(tgt, tgt_proj, src_proj)
    (pt, pb), (pl, pr) = win["pad"]
    padded = np.pad(src, ((0, 0), (pt, pb), (pl, pr)), mode="constant", constant_values=fill)
    out = None
    nty, ntx = tgt.n_tiles
    for ty in range(nty):
        for tx in range(ntx):
            r0, c0 = ty * tgt.tile_h, tx * tgt.tile_w
            r1, c1 = min(r0 + tgt.tile_h, tgt.height), min(c0 + tgt.tile_w, tgt.width)
            j0, i0 = int(win["j0"][ty, tx]) + pt, int(win["i0"][ty, tx]) + pl
            window = padded[:, j0:j0 + win["win_h"], i0:i0 + win["win_w"]]
            block = sample_window(xx[r0:r1, c0:c1], yy[r0:r1, c0:c1], window, win["x0"][ty, tx], win["y0"][ty, tx],
                                  src_x_res, src_y_res, method)
            if out is None:
                out = np.empty((src.shape[0], tgt.height, tgt.width), dtype=block.dtype)
            out[:, r0:r1, c0:c1] = block
    return out[0] if squeeze else out


def predownscale(src, x, y, tgt: ogrid.RegularGrid, tgt_proj, src_proj, interp=None, agg=None, recover_nan=False,
                 limit=0.95):
    """``_downscale_source_dataset`` (reproject.py:338-382) for one (bands, h, w) variable with 1-D
    source pixel-centre coordinates x (increasing) and y (decreasing).

    Returns (src, x, y, x_res, y_res) -- unchanged unless the source is finer than the target by
    more than ``limit`` in x or y; then the source is clipped to the transformed target box +-2 px
    (utils.py:77-124) and resampled to the intermediate grid (affine.py:52-137).
    """
    from . import resample as ores

    src = np.asarray(src)
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    x_res, y_res = ogrid.axis_resolution(x), ogrid.axis_resolution(y)  # coords.py:143-164
    box = oproj.transform_bounds(tgt_proj, src_proj, tgt.x_min, tgt.y_min, tgt.x_max, tgt.y_max)
    xres_t = (box[2] - box[0]) / tgt.width
    yres_t = (box[3] - box[1]) / tgt.height
    x_scale, y_scale = x_res / xres_t, y_res / yres_t
    if not (x_scale < limit or y_scale < limit):
        return src, x, y, x_res, y_res
    box = (box[0] - 2 * x_res, box[1] - 2 * y_res, box[2] + 2 * x_res, box[3] + 2 * y_res)
    keep_x = (x >= box[0]) & (x <= box[2])  # label-based, inclusive (xarray .sel with slices)
    keep_y = (y >= box[1]) & (y <= box[3])
    src, x, y = src[..., keep_y, :][..., keep_x], x[keep_x], y[keep_y]
    h, w = src.shape[-2:]
    # grid mapping of the clipped source (1-D coordinates -> regular, j axis down)
    # (coords.py:300-326: the box comes from the first / last coordinate +- half a pixel)
    f = ogrid.to_int_or_float
    source_grid = ogrid.RegularGrid(w, h, w, h, f(x[0] - x_res / 2), f(y[-1] - y_res / 2), f(x[-1] + x_res / 2),
                                    f(y[0] + y_res / 2), x_res, y_res, False)
    nw, nh = round(x_scale * w), round(y_scale * h)
    inter = ogrid.regular_grid((max(nw, 2), max(nh, 2)), (source_grid.x_min, source_grid.y_min), (xres_t, yres_t))
    if interp == "triangular":
        interp = "bilinear"
    out = ores.affine_transform(src, source_grid, inter, interp=interp, agg=agg, recover_nan=recover_nan)
    # the reference re-derives the grid mapping from the resampled dataset's coordinates (:380)
    xc, yc = ogrid.x_centres(inter), ogrid.y_centres(inter)
    return out, xc, yc, ogrid.axis_resolution(xc), ogrid.axis_resolution(yc)


def reproject_dataset_like(src, x, y, tgt: ogrid.RegularGrid, tgt_proj, src_proj, method=None, fill=None, agg=None):
    """``reproject_dataset`` (reproject.py:51-186) for one variable given 1-D source coordinates:
    flip a j-axis-up source, optional pre-downscale, windows, transform, sampling."""
    from . import resample as ores

    src = np.asarray(src)
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    if y[-1] > y[0]:  # reproject.py:115-118
        src, y = src[..., ::-1, :], y[::-1]
    interp = ores.default_interp(src.dtype) if method is None else method
    interp = {0: "nearest", 1: "bilinear"}.get(interp, interp)
    fill = ores.default_fill(src.dtype) if fill is None else fill
    src, x, y, x_res, y_res = predownscale(src, x, y, tgt, tgt_proj, src_proj, interp=interp, agg=agg)
    return reproject(src, float(x[0]), float(y[0]), x_res, y_res, float(y[1] - y[0]), tgt, tgt_proj, src_proj, interp,
                     fill)

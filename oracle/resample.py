"""Oracle restatement of the affine / coarsen path (test infrastructure only).

The reference does this arithmetic in third-party code that IS installed here:
``scipy.ndimage.affine_transform`` (what ``dask_image.ndinterp.affine_transform``
wraps, ``affine.py:353-362``) and numpy's (nan-)reducers (``coarsen.py:72-111``,
``constants.py:51-65``).  The oracle therefore calls those directly and only
restates the reference's own glue: ``affine.py:243-362`` (scale / divisor logic,
NaN recovery), ``coarsen.py:50-155`` (reducers) and dask's ``coarsen`` reshape
convention (``coarsen.py:34-47``).

dask-image evaluates the transform per OUTPUT chunk with a chunk-adjusted offset;
its source is not in the container, so the oracle evaluates the whole array in one
scipy call ("single chunk" semantics).  For tile sizes that split the output this
can differ from the reference in the last ulp of the source coordinate --
"parity unpinned" for that detail, stated in DESIGN.md.
"""

import math
import warnings

import numpy as np
from scipy import ndimage as ndi

from ._lib import lib, ptr

AGG_NAMES = ("center", "count", "first", "last", "max", "mean", "median", "mode", "min", "prod", "std", "sum", "var")


# ---------------------------------------------------------------------------
# coarsen.py reducers on a (.., h, f_j, w, f_i) block, axis = the two window axes
# ---------------------------------------------------------------------------
def _pick(block, axis, which):
    index = []
    for i in range(block.ndim):
        if i in axis:
            index.append({"first": 0, "last": -1, "center": block.shape[i] // 2}[which])
        else:
            index.append(slice(None))
    return block[tuple(index)]


def _reduce(reducer, nan_reducer, block, axis):
    """coarsen.py:93-111."""
    if np.issubdtype(block.dtype, np.floating):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            return nan_reducer(block, axis)
    a = reducer(block, axis)
    if np.issubdtype(a.dtype, np.floating):
        return np.rint(a).astype(block.dtype)
    return a


def mode(block, axis):
    """coarsen.py:114-155 (histogram in xrs_oracle.c:xrso_mode)."""
    ndim = len(axis)
    block = np.moveaxis(block, axis, range(-ndim, 0))
    flat = block.reshape(-1, int(np.prod(block.shape[-ndim:])))
    min_val = int(flat.min())
    max_val = int(flat.max())
    normalized = np.ascontiguousarray((flat - min_val).astype(np.int64))
    out = np.empty(flat.shape[0], dtype=np.int64)
    lib().xrso_mode(ptr(normalized), flat.shape[0], flat.shape[1], min_val, max_val - min_val + 1, ptr(out))
    return out.reshape(block.shape[:-ndim])


def reducer(name):
    """constants.py:51-65 (AGG_METHODS)."""
    if name in ("first", "last", "center"):
        return lambda b, axis: _pick(b, axis, name)
    table = {
        "mean": lambda b, axis: _reduce(np.mean, np.nanmean, b, axis),
        "median": lambda b, axis: _reduce(np.median, np.nanmedian, b, axis),
        "std": lambda b, axis: _reduce(np.std, np.nanstd, b, axis),
        "var": lambda b, axis: _reduce(np.var, np.nanvar, b, axis),
        "sum": np.nansum,
        "prod": np.nanprod,
        "max": np.nanmax,
        "min": np.nanmin,
        "count": np.count_nonzero,
        "mode": mode,
    }
    return table[name]


def coarsen(array, f_j, f_i, agg):
    """``da.coarsen(agg, array, {ndim-2: f_j, ndim-1: f_i})`` on one chunk (affine.py:308-310)."""
    array = np.asarray(array)
    *lead, h, w = array.shape
    assert h % f_j == 0 and w % f_i == 0
    block = array.reshape(*lead, h // f_j, f_j, w // f_i, f_i)
    axis = (block.ndim - 3, block.ndim - 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        return np.asarray(reducer(agg)(block, axis))


# ---------------------------------------------------------------------------
# affine.py:243-362
# ---------------------------------------------------------------------------
def upscale(array, matrix, output_shape, interp, recover_nan, fill_value):
    """affine.py:316-362 (_upscale)."""
    ((i_scale, _, i_off), (_, j_scale, j_off)) = matrix
    array = np.asarray(array)
    offset = (array.ndim - 2) * (0,) + (j_off, i_off)
    scale = (array.ndim - 2) * (1,) + (j_scale, i_scale)
    m = np.diag(scale)
    if interp > 1:
        raise ValueError(
            "interp_methods must be one of 0, 1, 'nearest', 'bilinear'. "
            "Higher order is not supported for 3D arrays in affine transforms, "
            "as it causes unintended blending across the non-spatial (e.g., time) "
            "dimension."
        )
    kw = dict(offset=offset, order=interp, output_shape=tuple(output_shape), mode="constant", cval=fill_value)
    if recover_nan and interp > 0:
        mask = np.isnan(array)
        if np.any(mask):
            filled = np.where(mask, 0.0, array)
            scaled_im = ndi.affine_transform(filled, m, **kw)
            scaled_norm = ndi.affine_transform(1.0 - mask, m, **kw)
            with np.errstate(invalid="ignore", divide="ignore"):
                return np.where(np.isclose(scaled_norm, 0.0), np.nan, scaled_im / scaled_norm)
    return ndi.affine_transform(array, m, **kw)


def resample_array(array, matrix, output_shape, interp, agg, recover_nan, fill_value):
    """affine.py:243-313 (_resample_array / _downscale)."""
    ((i_scale, m01, i_off), (m10, j_scale, j_off)) = matrix
    if (matrix[0][0] > 1 or matrix[1][0] > 1) and interp != 0:
        j_div = math.ceil(abs(j_scale))
        i_div = math.ceil(abs(i_scale))
        m = ((i_scale / i_div, m01, i_off), (m10, j_scale / j_div, j_off))
        big = tuple(output_shape[:-2]) + (output_shape[-2] * j_div, output_shape[-1] * i_div)
        fine = upscale(array, m, big, interp, recover_nan, fill_value)
        return coarsen(fine, j_div, i_div, agg)
    return upscale(array, matrix, output_shape, interp, recover_nan, fill_value)


def default_interp(dtype):
    """utils.py:197-198."""
    return 0 if np.issubdtype(dtype, np.integer) else 1


def default_agg(dtype):
    """utils.py:259-260."""
    return "center" if np.issubdtype(dtype, np.integer) else "mean"


def default_fill(dtype):
    """utils.py:307-316."""
    if dtype == np.uint8:
        return 255
    if dtype == np.uint16:
        return 65535
    if np.issubdtype(dtype, np.integer):
        return -1
    return np.nan


def affine_transform(array, source_grid, target_grid, interp=None, agg=None, recover_nan=False, fill_value=None):
    """affine.py:119-129 for one variable: matrix from the grids, then _resample_array."""
    from .grid import ij_transform_to

    array = np.asarray(array)
    interp = default_interp(array.dtype) if interp is None else {"nearest": 0, "bilinear": 1}.get(interp, interp)
    agg = default_agg(array.dtype) if agg is None else agg
    fill_value = default_fill(array.dtype) if fill_value is None else fill_value
    matrix = ij_transform_to(target_grid, source_grid)
    out_shape = array.shape[:-2] + (target_grid.height, target_grid.width)
    return resample_array(array, matrix, out_shape, interp, agg, recover_nan, fill_value)

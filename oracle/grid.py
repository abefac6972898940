"""Oracle-side regular grid description (test infrastructure only).

Restates just the arithmetic of the reference's regular ``GridMapping`` that
feeds the hot path: ``gridmapping/regular.py:87-129`` (normalisation),
``gridmapping/base.py:503-533`` (tile boxes), ``gridmapping/regular.py:44-63``
(pixel-centre coordinates), ``gridmapping/base.py:436-478`` +
``gridmapping/helpers.py:51-56`` (affine algebra of the ``affine`` package).
"""

import dataclasses
import math

import numpy as np


def to_int_or_float(x):
    """gridmapping/helpers.py:39-48."""
    if isinstance(x, (int, np.integer)):
        return int(x)
    xf = float(x)
    xi = round(xf)
    return xi if math.isclose(xi, xf, rel_tol=1e-5) else xf


@dataclasses.dataclass(frozen=True)
class RegularGrid:
    width: int
    height: int
    tile_w: int
    tile_h: int
    x_min: float
    y_min: float
    x_max: float
    y_max: float
    x_res: float
    y_res: float
    is_j_axis_up: bool = False

    @property
    def n_tiles(self):
        return (-(-self.height // self.tile_h), -(-self.width // self.tile_w))


def regular_grid(size, xy_min, xy_res, tile_size=None, is_j_axis_up=False) -> RegularGrid:
    """gridmapping/regular.py:87-129 (new_regular_grid_mapping)."""
    width, height = (size, size) if isinstance(size, int) else (int(size[0]), int(size[1]))
    if isinstance(xy_res, (int, float)):
        xy_res = (xy_res, xy_res)
    x_res, y_res = to_int_or_float(xy_res[0]), to_int_or_float(xy_res[1])
    x_min, y_min = to_int_or_float(xy_min[0]), to_int_or_float(xy_min[1])
    x_max = to_int_or_float(x_min + x_res * width)
    y_max = to_int_or_float(y_min + y_res * height)
    if tile_size is None:
        tw, th = width, height
    elif isinstance(tile_size, int):
        tw, th = tile_size, tile_size
    else:
        tw, th = int(tile_size[0]), int(tile_size[1])
    return RegularGrid(width, height, tw, th, x_min, y_min, x_max, y_max, x_res, y_res, bool(is_j_axis_up))


def tile_ij_bboxes(g: RegularGrid) -> np.ndarray:
    """gridmapping/base.py:503-519; row-major tiles, (i0, j0, i1, j1)."""
    nty, ntx = g.n_tiles
    out = np.empty((nty * ntx, 4), dtype=np.int64)
    k = 0
    for ty in range(nty):
        for tx in range(ntx):
            out[k] = (
                tx * g.tile_w,
                ty * g.tile_h,
                min((tx + 1) * g.tile_w, g.width),
                min((ty + 1) * g.tile_h, g.height),
            )
            k += 1
    return out


def tile_xy_bboxes(g: RegularGrid) -> np.ndarray:
    """gridmapping/base.py:521-533."""
    ij = tile_ij_bboxes(g)
    if g.is_j_axis_up:
        off = np.array([g.x_min, g.y_min, g.x_min, g.y_min])
        scale = np.array([g.x_res, g.y_res, g.x_res, g.y_res])
        xy = off + scale * ij
    else:
        off = np.array([g.x_min, g.y_max, g.x_min, g.y_max])
        scale = np.array([g.x_res, -g.y_res, g.x_res, -g.y_res])
        xy = off + scale * ij
        xy[:, [1, 3]] = xy[:, [3, 1]]
    return xy


def _tiled_linspace(start, stop, num, chunk):
    """dask.array.linspace(start, stop, num, chunks=chunk) restated: dask builds each
    chunk as np.linspace(blockstart, blockstop, bs) with step=(stop-start)/(num-1),
    blockstart accumulated as blockstart + step*bs (dask/array/creation.py, from
    memory -- source not in container); gridmapping/regular.py:44-63 is the caller."""
    num = int(num)
    step = (stop - start) / (num - 1) if num > 1 else 0.0
    out = np.empty(num, dtype=np.float64)
    blockstart = start
    pos = 0
    while pos < num:
        bs = min(chunk, num - pos)
        bs_space = bs - 1
        blockstop = blockstart + bs_space * step
        out[pos : pos + bs] = np.linspace(blockstart, blockstop, bs)
        blockstart = blockstart + step * bs
        pos += bs
    return out


def x_centres(g: RegularGrid) -> np.ndarray:
    """gridmapping/regular.py:44-52."""
    return _tiled_linspace(g.x_min + g.x_res / 2, g.x_max - g.x_res / 2, g.width, g.tile_w)


def y_centres(g: RegularGrid) -> np.ndarray:
    """gridmapping/regular.py:54-63."""
    y1, y2 = g.y_min + g.y_res / 2, g.y_max - g.y_res / 2
    if not g.is_j_axis_up:
        y1, y2 = y2, y1
    return _tiled_linspace(y1, y2, g.height, g.tile_h)


# --- the `affine` package's 2x3 algebra (affine>=2.2, from memory; pinned by the
# exact tuples in tests/gridmapping/test_base.py:174-252 of the reference) ---------
def affine_mul(m1, m2):
    (sa, sb, sc), (sd, se, sf) = m1
    (oa, ob, oc), (od, oe, of) = m2
    return (
        (sa * oa + sb * od, sa * ob + sb * oe, sa * oc + sb * of + sc),
        (sd * oa + se * od, sd * ob + se * oe, sd * oc + se * of + sf),
    )


def affine_inv(m):
    (sa, sb, sc), (sd, se, sf) = m
    idet = 1.0 / (sa * se - sb * sd)
    ra, rb, rd, re = se * idet, -sb * idet, -sd * idet, sa * idet
    return ((ra, rb, -sc * ra - sf * rb), (rd, re, -sc * rd - sf * re))


def ij_to_xy(g: RegularGrid):
    """gridmapping/base.py:436-451."""
    if g.is_j_axis_up:
        return ((g.x_res, 0.0, g.x_min), (0.0, g.y_res, g.y_min))
    return ((g.x_res, 0.0, g.x_min), (0.0, -g.y_res, g.y_max))


def xy_to_ij(g: RegularGrid):
    """gridmapping/base.py:453-459."""
    return affine_inv(ij_to_xy(g))


def ij_transform_to(g: RegularGrid, other: RegularGrid):
    """gridmapping/base.py:461-478: image coords of *g* -> image coords of *other*."""
    return affine_mul(xy_to_ij(other), ij_to_xy(g))


# --- resolution a 1-D coordinate axis is given by the reference -----------------------------
def round_to_fraction(value: float, digits: int = 2, resolution: float = 1.0):
    """gridmapping/helpers.py:192-239: keep ``digits`` significant digits and round the remainder
    to a multiple of ``resolution`` (0.1, 0.2, 0.25, 0.5 or 1) of the last kept digit; exact
    rational arithmetic, returned as ``Fraction``."""
    from fractions import Fraction

    if value == 0:
        return Fraction(0)
    sign = -1 if value < 0 else 1
    value = abs(value)
    # the step expressed as an integer number of units one (or, for quarters, two) decimal
    # places below the last kept digit -- the same table as helpers.py:192-198
    unit, shift = {10: (1, 0), 20: (2, 0), 25: (25, 1), 50: (5, 0), 100: (1, -1)}[round(100 * resolution)]
    exponent = math.floor(math.log10(value)) - digits - shift
    magnitude = Fraction(10) ** exponent
    return sign * (unit * round(value / magnitude / unit)) * magnitude


def axis_resolution(coords, tolerance=1e-5) -> float:
    """gridmapping/coords.py:143-164: resolution the reference derives from a 1-D axis."""
    d = np.abs(np.diff(np.asarray(coords, dtype=np.float64)))
    d = np.where(d == 0, np.nan, d)
    r = d[0]
    if np.allclose(d, r, atol=tolerance):
        r = round_to_fraction(float(r), 5, 0.25)
    else:
        r = round_to_fraction(float(np.nanmedian(d)), 2, 0.5)
    return to_int_or_float(r)

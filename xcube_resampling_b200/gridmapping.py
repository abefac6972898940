"""GridMapping: image grid + CRS, the argument type of every entry point.

A from-scratch host-side restatement of the subset of the reference's
``GridMapping`` that the resampling path reads:
``gridmapping/base.py:59-913``, ``regular.py:38-166``, ``coords.py:49-337``,
``helpers.py:39-255``.  Coordinates are plain numpy on the host; device copies
are made by the entry points (or stay there: :meth:`GridMapping.from_device_coords`).
CF-convention discovery lives in ``cfconv.py``.
"""

from __future__ import annotations

import copy
import math
import threading
from fractions import Fraction
from typing import Any

import numpy as np

from .crs import CRS, CRS_CRS84, CRS_WGS84, normalize_crs  # noqa: F401
from .dataset import DataArray, Dataset, from_any

DEFAULT_TOLERANCE = 1.0e-5  # gridmapping/base.py:56
_ER = 6371000  # gridmapping/coords.py:46

AffineTransformMatrix = tuple


# ---------------------------------------------------------------------------
# helpers (gridmapping/helpers.py)
# ---------------------------------------------------------------------------
def _to_int_or_float(x):
    """helpers.py:39-48."""
    if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
        return int(x)
    xf = float(x)
    xi = round(xf)
    return xi if math.isclose(xi, xf, rel_tol=1e-5) else xf


def _normalize_int_pair(value, name=None, default="__undefined__"):
    """helpers.py:65-78."""
    if isinstance(value, (int, np.integer)):
        return int(value), int(value)
    if value is not None:
        x, y = value
        return int(x), int(y)
    if default != "__undefined__":
        return default
    raise ValueError(f"{name} must be an int or a sequence of two ints")


def _normalize_number_pair(value, name=None, default="__undefined__"):
    """helpers.py:81-95."""
    if isinstance(value, (float, int, np.integer, np.floating)):
        return _to_int_or_float(value), _to_int_or_float(value)
    if value is not None:
        x, y = value
        return _to_int_or_float(x), _to_int_or_float(y)
    if default != "__undefined__":
        return default
    raise ValueError(f"{name} must be a number or a sequence of two numbers")


def _assert_valid_xy_names(value, name=None):
    if not isinstance(value, tuple):
        raise TypeError(f"{name or 'value'} must be an instance of {tuple}, was {type(value)}")
    if not (len(value) == 2 and all(value) and value[0] != value[1]):
        raise ValueError(f"invalid {name or 'value'}")


_RESOLUTIONS = {10: (1, 0), 20: (2, 0), 25: (25, 1), 50: (5, 0), 100: (1, -1)}


def round_to_fraction(value: float, digits: int = 2, resolution: float = 1) -> Fraction:
    """helpers.py:203-239: round at *digits* significant digits in steps of *resolution*."""
    if digits < 1:
        raise ValueError("digits must be a positive integer")
    key = round(100 * resolution)
    if key not in _RESOLUTIONS or not math.isclose(100 * resolution, key):
        raise ValueError(f"resolution must be one of {sorted(k / 100 for k in _RESOLUTIONS)}")
    if value == 0:
        return Fraction(0, 1)
    sign = 1
    if value < 0:
        sign, value = -1, -value
    res, res_digits = _RESOLUTIONS[key]
    exponent = math.floor(math.log10(value)) - digits - res_digits
    magnitude = Fraction(10**exponent, 1) if exponent >= 0 else Fraction(1, 10**-exponent)
    scaled = value / magnitude
    discrete = res * round(scaled / res)
    return (sign * discrete) * magnitude


def scale_xy_res_and_size(xy_res, size, xy_scale):
    """helpers.py:242-255."""
    x_res, y_res = xy_res
    x_scale, y_scale = xy_scale
    w, h = size
    w, h = round(x_scale * w), round(y_scale * h)
    return (x_res / x_scale, y_res / y_scale), (w if w >= 2 else 2, h if h >= 2 else 2)


def to_lon_360(lon):
    lon = np.asarray(lon)
    return np.where(lon >= 0.0, lon, lon + 360.0)


def from_lon_360(lon):
    lon = np.asarray(lon)
    return np.where(lon <= 180.0, lon, lon - 360.0)


# the `affine` package's 2x3 algebra (helpers.py:51-56 wraps it): product and inverse
def _affine_mul(m1, m2):
    (sa, sb, sc), (sd, se, sf) = m1
    (oa, ob, oc), (od, oe, of) = m2
    return (
        (sa * oa + sb * od, sa * ob + sb * oe, sa * oc + sb * of + sc),
        (sd * oa + se * od, sd * ob + se * oe, sd * oc + se * of + sf),
    )


def _affine_inv(m):
    (sa, sb, sc), (sd, se, sf) = m
    idet = 1.0 / (sa * se - sb * sd)
    ra, rb, rd, re = se * idet, -sb * idet, -sd * idet, sa * idet
    return ((ra, rb, -sc * ra - sf * rb), (rd, re, -sc * rd - sf * re))


def _tiled_linspace(start, stop, num, chunk) -> np.ndarray:
    """``dask.array.linspace(start, stop, num, chunks=chunk)`` evaluated eagerly.

    regular.py:44-63 builds pixel-centre coordinates that way; dask computes each
    chunk with ``np.linspace`` from a running block start, which differs from a
    single ``np.linspace`` in the last ulp.  Kept so that coordinates (and
    therefore exact half-pixel ties) match the reference.
    """
    num = int(num)
    step = (stop - start) / (num - 1) if num > 1 else 0.0
    out = np.empty(num, dtype=np.float64)
    block_start = start
    pos = 0
    while pos < num:
        n = min(int(chunk), num - pos)
        block_stop = block_start + (n - 1) * step
        out[pos:pos + n] = np.linspace(block_start, block_stop, n)
        block_start = block_start + step * n
        pos += n
    return out


# ---------------------------------------------------------------------------
# GridMapping
# ---------------------------------------------------------------------------
class GridMapping:
    """Image geometry: size, tiling, bounding box, resolution, CRS, orientation.

    Create with :meth:`regular`, :meth:`from_coords` or :meth:`from_dataset`;
    derive with :meth:`derive`, :meth:`scale`, :meth:`to_regular`.  Thread-safe.
    """

    def __init__(self, /, size, tile_size, xy_bbox, xy_res, crs, xy_var_names, xy_dim_names, is_regular=None,
                 is_lon_360=None, is_j_axis_up=None, x_coords=None, y_coords=None):
        width, height = _normalize_int_pair(size, name="size")
        if not (width > 1 and height > 1):
            raise ValueError("invalid size")
        tile_width, tile_height = _normalize_int_pair(tile_size, default=(width, height))
        if not (tile_width > 1 and tile_height > 1):
            raise ValueError("invalid tile_size")
        if not xy_bbox:
            raise ValueError("xy_bbox must be given")
        if not xy_res:
            raise ValueError("xy_res must be given")
        _assert_valid_xy_names(xy_var_names, name="xy_var_names")
        _assert_valid_xy_names(xy_dim_names, name="xy_dim_names")
        if not isinstance(crs, CRS):
            raise TypeError(f"crs must be an instance of {CRS}, was {type(crs)}")
        for nm, c in (("x_coords", x_coords), ("y_coords", y_coords)):
            if c is not None and np.ndim(c) not in (1, 2):
                raise ValueError(f"{nm}.ndim must be 1 or 2, was {np.ndim(c)}")
        x_min, y_min, x_max, y_max = xy_bbox
        x_res, y_res = _normalize_number_pair(xy_res, name="xy_res")
        if not (x_res > 0 and y_res > 0):
            raise ValueError("invalid xy_res")
        self._lock = threading.RLock()
        self._size = width, height
        self._tile_size = tile_width, tile_height
        self._xy_bbox = x_min, y_min, x_max, y_max
        self._xy_res = x_res, y_res
        self._crs = crs
        self._xy_var_names = xy_var_names
        self._xy_dim_names = xy_dim_names
        self._is_regular = is_regular
        self._is_lon_360 = is_lon_360
        self._is_j_axis_up = is_j_axis_up
        self._x_coords = None if x_coords is None else np.asarray(x_coords)
        self._y_coords = None if y_coords is None else np.asarray(y_coords)
        self._coords_lazy = x_coords is None and y_coords is None
        self._x_dev = self._y_dev = None  # 2-D coordinates resident on a GPU (from_device_coords)

    # -- derived instances --------------------------------------------------
    def derive(self, /, xy_var_names=None, xy_dim_names=None, tile_size=None, is_j_axis_up=None) -> "GridMapping":
        """base.py:145-205."""
        other = copy.copy(self)
        other._lock = threading.RLock()
        if xy_var_names is not None:
            _assert_valid_xy_names(xy_var_names, name="xy_var_names")
            other._xy_var_names = xy_var_names
        if xy_dim_names is not None:
            _assert_valid_xy_names(xy_dim_names, name="xy_dim_names")
            other._xy_dim_names = xy_dim_names
        if tile_size is not None:
            tw, th = _normalize_int_pair(tile_size, name="tile_size")
            if not (tw > 1 and th > 1):
                raise ValueError("invalid tile_size")
            if other.tile_size != (tw, th):
                other._tile_size = tw, th
                if other._coords_lazy:
                    # generated coordinates follow the tile-sized chunking (regular.py:44-63)
                    other._x_coords = other._y_coords = None
        if is_j_axis_up is not None and is_j_axis_up != other._is_j_axis_up:
            other._is_j_axis_up = is_j_axis_up
            if other._y_coords is not None:
                other._y_coords = other._y_coords[::-1]
            if other._x_coords is not None and other._x_coords.ndim == 2:
                other._x_coords = other._x_coords[::-1]
        return other

    def scale(self, xy_scale, tile_size=None) -> "GridMapping":
        """base.py:207-246."""
        self._assert_regular()
        x_scale, y_scale = _normalize_number_pair(xy_scale)
        new_xy_res, new_size = scale_xy_res_and_size(self.xy_res, self.size, (x_scale, y_scale))
        tw, th = _normalize_int_pair(tile_size, name="tile_size") if tile_size is not None else self.tile_size
        tw, th = min(new_size[0], tw), min(new_size[1], th)
        return GridMapping.regular(new_size, (self.x_min, self.y_min), new_xy_res, self.crs, tile_size=(tw, th),
                                   is_j_axis_up=self.is_j_axis_up).derive(
            xy_dim_names=self.xy_dim_names, xy_var_names=self.xy_var_names)

    # -- plain properties -----------------------------------------------------
    size = property(lambda self: self._size)
    width = property(lambda self: self._size[0])
    height = property(lambda self: self._size[1])
    tile_size = property(lambda self: self._tile_size)
    tile_width = property(lambda self: self._tile_size[0])
    tile_height = property(lambda self: self._tile_size[1])
    is_tiled = property(lambda self: self._size != self._tile_size)
    xy_var_names = property(lambda self: self._xy_var_names)
    xy_dim_names = property(lambda self: self._xy_dim_names)
    xy_bbox = property(lambda self: self._xy_bbox)
    x_min = property(lambda self: self._xy_bbox[0])
    y_min = property(lambda self: self._xy_bbox[1])
    x_max = property(lambda self: self._xy_bbox[2])
    y_max = property(lambda self: self._xy_bbox[3])
    xy_res = property(lambda self: self._xy_res)
    x_res = property(lambda self: self._xy_res[0])
    y_res = property(lambda self: self._xy_res[1])
    crs = property(lambda self: self._crs)
    spatial_unit_name = property(lambda self: self._crs.unit_name)  # base.py:402-404 (axis_info[0].unit_name)
    is_lon_360 = property(lambda self: self._is_lon_360)
    is_regular = property(lambda self: self._is_regular)
    is_j_axis_up = property(lambda self: self._is_j_axis_up)
    ij_bbox = property(lambda self: (0, 0, self.width, self.height))

    @property
    def xy_coords_chunks(self):
        return 2, self.tile_height, self.tile_width

    # -- coordinates ----------------------------------------------------------
    def _computed(self, attr, fn):
        value = getattr(self, attr)
        if value is not None:
            return value
        with self._lock:
            value = getattr(self, attr)
            if value is None:
                value = fn()
                setattr(self, attr, value)
            return value

    def _new_x(self) -> np.ndarray:
        """regular.py:44-52."""
        if self._x_dev is not None:  # device-resident 2-D coordinates: fetched only when asked for
            from . import _dev

            return _dev.to_host(self._x_dev)
        self._assert_regular()
        return _tiled_linspace(self.x_min + self.x_res / 2, self.x_max - self.x_res / 2, self.width, self.tile_width)

    def _new_y(self) -> np.ndarray:
        """regular.py:54-63."""
        if self._y_dev is not None:
            from . import _dev

            return _dev.to_host(self._y_dev)
        self._assert_regular()
        y1, y2 = self.y_min + self.y_res / 2, self.y_max - self.y_res / 2
        if not self.is_j_axis_up:
            y1, y2 = y2, y1
        return _tiled_linspace(y1, y2, self.height, self.tile_height)

    @property
    def x_values(self) -> np.ndarray:
        """x coordinates as numpy: (width,) or (height, width)."""
        return self._computed("_x_coords", self._new_x)

    @property
    def y_values(self) -> np.ndarray:
        return self._computed("_y_coords", self._new_y)

    @property
    def x_coords(self) -> DataArray:
        v = self.x_values
        dims = (self.xy_dim_names[0],) if v.ndim == 1 else (self.xy_dim_names[1], self.xy_dim_names[0])
        return DataArray(v, dims=dims, name=self.xy_var_names[0])

    @property
    def y_coords(self) -> DataArray:
        v = self.y_values
        dims = (self.xy_dim_names[1],) if v.ndim == 1 else (self.xy_dim_names[1], self.xy_dim_names[0])
        return DataArray(v, dims=dims, name=self.xy_var_names[1])

    @property
    def xy_coords(self) -> DataArray:
        """(2, height, width) coordinates in CRS units (base.py:309-316)."""
        x, y = self.x_values, self.y_values
        if x.ndim == 1:
            yy, xx = np.broadcast_arrays(y[:, None], x[None, :])
        else:
            xx, yy = x, y
        return DataArray(np.stack([xx, yy]), dims=("coord", self.xy_dim_names[1], self.xy_dim_names[0]),
                         name="xy_coords")

    # -- affine transforms (base.py:436-496) ----------------------------------
    @property
    def ij_to_xy_transform(self) -> AffineTransformMatrix:
        self._assert_regular()
        if self.is_j_axis_up:
            return ((self.x_res, 0.0, self.x_min), (0.0, self.y_res, self.y_min))
        return ((self.x_res, 0.0, self.x_min), (0.0, -self.y_res, self.y_max))

    @property
    def xy_to_ij_transform(self) -> AffineTransformMatrix:
        self._assert_regular()
        return _affine_inv(self.ij_to_xy_transform)

    def ij_transform_to(self, other: "GridMapping") -> AffineTransformMatrix:
        self._assert_regular()
        self.assert_regular(other, name="other")
        return _affine_mul(other.xy_to_ij_transform, self.ij_to_xy_transform)

    def ij_transform_from(self, other: "GridMapping") -> AffineTransformMatrix:
        self._assert_regular()
        self.assert_regular(other, name="other")
        return _affine_inv(self.ij_transform_to(other))

    # -- tiles (base.py:498-533) ----------------------------------------------
    @property
    def ij_bboxes(self) -> np.ndarray:
        nty = -(-self.height // self.tile_height)
        ntx = -(-self.width // self.tile_width)
        out = np.empty((nty * ntx, 4), dtype=np.int64)
        k = 0
        for ty in range(nty):
            for tx in range(ntx):
                out[k] = (tx * self.tile_width, ty * self.tile_height,
                          min((tx + 1) * self.tile_width, self.width), min((ty + 1) * self.tile_height, self.height))
                k += 1
        return out

    @property
    def xy_bboxes(self) -> np.ndarray:
        if self.is_j_axis_up:
            off = np.array([self.x_min, self.y_min, self.x_min, self.y_min])
            scale = np.array([self.x_res, self.y_res, self.x_res, self.y_res])
            xy = off + scale * self.ij_bboxes
        else:
            off = np.array([self.x_min, self.y_max, self.x_min, self.y_max])
            scale = np.array([self.x_res, -self.y_res, self.x_res, -self.y_res])
            xy = off + scale * self.ij_bboxes
            xy[:, [1, 3]] = xy[:, [3, 1]]
        return xy

    def ij_bbox_from_xy_bbox(self, xy_bbox, xy_border: float = 0.0, ij_border: int = 0):
        """base.py:535-563."""
        boxes = self.ij_bboxes_from_xy_bboxes(np.array([xy_bbox], dtype=np.float64), xy_border=xy_border,
                                              ij_border=ij_border)
        return tuple(map(int, boxes[0]))

    def ij_bboxes_from_xy_bboxes(self, xy_bboxes, xy_border: float = 0.0, ij_border: int = 0, ij_bboxes=None):
        """base.py:565-629 on the device (K0); boxes must form a separable grid."""
        from .rectify import tile_source_windows_for_boxes

        out = tile_source_windows_for_boxes(self, np.asarray(xy_bboxes, dtype=np.float64), xy_border, ij_border)
        if ij_bboxes is not None:
            ij_bboxes[:, :] = out
            return ij_bboxes
        return out

    # -- coordinate variables (coords.py:340-472) -----------------------------
    def to_coords(self, xy_var_names=None, xy_dim_names=None, exclude_bounds: bool = False,
                  reuse_coords: bool = False) -> dict:
        self._assert_regular()
        if xy_var_names:
            _assert_valid_xy_names(xy_var_names, name="xy_var_names")
        if xy_dim_names:
            _assert_valid_xy_names(xy_dim_names, name="xy_dim_names")
        x_name, y_name = xy_var_names or self.xy_var_names
        x_dim, y_dim = xy_dim_names or self.xy_dim_names
        if reuse_coords and self._x_coords is not None and self._y_coords is not None and \
                self._x_coords.ndim == 1 and self._y_coords.ndim == 1:
            return {x_name: DataArray(self._x_coords, dims=x_dim), y_name: DataArray(self._y_coords, dims=y_dim)}
        w, h = self.size
        x1, y1, x2, y2 = self.xy_bbox
        x_res, y_res = self.xy_res
        xh, yh = x_res / 2, y_res / 2
        x_data = np.linspace(x1 + xh, x2 - xh, w, dtype=np.float64)
        if self.is_lon_360:
            x_data = from_lon_360(x_data)
        if self.is_j_axis_up:
            y_data = np.linspace(y1 + yh, y2 - yh, h, dtype=np.float64)
        else:
            y_data = np.linspace(y2 - yh, y1 + yh, h, dtype=np.float64)
        if self.crs.is_geographic:
            x_attrs = dict(long_name="longitude coordinate", standard_name="longitude", units="degrees_east")
            y_attrs = dict(long_name="latitude coordinate", standard_name="latitude", units="degrees_north")
        else:
            x_attrs = dict(long_name="x coordinate of projection", standard_name="projection_x_coordinate")
            y_attrs = dict(long_name="y coordinate of projection", standard_name="projection_y_coordinate")
        coords = {x_name: DataArray(x_data, dims=x_dim, attrs=x_attrs), y_name: DataArray(y_data, dims=y_dim, attrs=y_attrs)}
        if not exclude_bounds:
            xb0, xb1 = np.linspace(x1, x2 - x_res, w), np.linspace(x1 + x_res, x2, w)
            if self.is_lon_360:
                xb0, xb1 = from_lon_360(xb0), from_lon_360(xb1)
            if self.is_j_axis_up:
                yb0, yb1 = np.linspace(y1, y2 - y_res, h), np.linspace(y1 + y_res, y2, h)
            else:
                yb0, yb1 = np.linspace(y2, y1 + y_res, h), np.linspace(y2 - y_res, y1, h)
            coords[x_name].attrs.update(bounds=f"{x_name}_bnds")
            coords[y_name].attrs.update(bounds=f"{y_name}_bnds")
            coords[f"{x_name}_bnds"] = DataArray(np.stack([xb0, xb1], axis=1), dims=(x_dim, "bnds"))
            coords[f"{y_name}_bnds"] = DataArray(np.stack([yb0, yb1], axis=1), dims=(y_dim, "bnds"))
        return coords

    # -- factories ------------------------------------------------------------
    @classmethod
    def regular(cls, size, xy_min, xy_res, crs, *, tile_size=None, is_j_axis_up: bool = False) -> "GridMapping":
        """base.py:704-738 -> regular.py:78-129."""
        width, height = _normalize_int_pair(size, name="size")
        if not (width > 1 and height > 1):
            raise ValueError("invalid size")
        x_min, y_min = _normalize_number_pair(xy_min, name="xy_min")
        x_res, y_res = _normalize_number_pair(xy_res, name="xy_res")
        if not (x_res > 0 and y_res > 0):
            raise ValueError("invalid xy_res")
        crs = normalize_crs(crs)
        x_min, y_min = _to_int_or_float(x_min), _to_int_or_float(y_min)
        x_max = _to_int_or_float(x_min + x_res * width)
        y_max = _to_int_or_float(y_min + y_res * height)
        if crs.is_geographic:
            if y_min < -90:
                raise ValueError("invalid y_min")
            if y_max > 90:
                raise ValueError("invalid size, y_min combination")
        names = ("lon", "lat") if crs.is_geographic else ("x", "y")
        return cls(crs=crs, size=(width, height), tile_size=tile_size or (width, height),
                   xy_bbox=(x_min, y_min, x_max, y_max), xy_res=(x_res, y_res), xy_var_names=names,
                   xy_dim_names=names, is_regular=True, is_lon_360=(x_max > 180) and crs.is_geographic,
                   is_j_axis_up=is_j_axis_up)

    def to_regular(self, tile_size=None, is_j_axis_up: bool = False) -> "GridMapping":
        """base.py:740-758 -> regular.py:132-166."""
        if self.is_regular:
            if tile_size is not None or is_j_axis_up != self.is_j_axis_up:
                return self.derive(tile_size=tile_size, is_j_axis_up=is_j_axis_up)
            return self
        x_min, y_min, x_max, y_max = self.xy_bbox
        x_res, y_res = self.xy_res
        xy_res = min(x_res, y_res) or max(x_res, y_res)
        width = round((x_max - x_min + xy_res) / xy_res)
        height = round((y_max - y_min + xy_res) / xy_res)
        width = width if width >= 2 else 2
        height = height if height >= 2 else 2
        if tile_size is None:
            tile_size = self.tile_size
        return GridMapping.regular(size=(width, height), xy_min=(x_min, y_min), xy_res=xy_res, crs=self.crs,
                                   tile_size=tile_size, is_j_axis_up=is_j_axis_up)

    def transform(self, crs, *, xy_res=None, tile_size=None, xy_var_names=None,
                  tolerance: float = DEFAULT_TOLERANCE) -> "GridMapping":
        """base.py:884-913 -> transform.py:57-125: this grid mapping with its coordinates expressed in
        another CRS.  The point transform runs on the device (xrs_transform_points), the bounding box
        of a given ``xy_res`` comes from edges densified with 101 points (transform.py:89)."""
        from .reproject import transform_bounds, transform_points  # device code; imported late (cycle)

        target_crs = normalize_crs(crs)
        if xy_var_names:
            _assert_valid_xy_names(xy_var_names, name="xy_var_names")
        if self.crs == target_crs:
            if tile_size is not None or xy_var_names is not None:
                return self.derive(tile_size=tile_size, xy_var_names=xy_var_names)
            return self
        xy = self.xy_coords.values
        x2, y2 = transform_points(np.ascontiguousarray(xy[0], dtype=np.float64),
                                  np.ascontiguousarray(xy[1], dtype=np.float64), self.crs, target_crs)
        xy_bbox = None
        if xy_res is not None:
            box = transform_bounds(self.crs, target_crs, np.asarray([self.xy_bbox], dtype=np.float64),
                                   densify_pts=101)[0]
            x_res, y_res = _normalize_number_pair(xy_res)
            xy_bbox = (float(box[0]) - x_res / 2, float(box[1]) - y_res / 2,
                       float(box[2]) + x_res / 2, float(box[3]) + y_res / 2)
        names = tuple(xy_var_names) if xy_var_names else ("transformed_x", "transformed_y")
        dims = (self.xy_dim_names[1], self.xy_dim_names[0])
        return GridMapping.from_coords(DataArray(x2, dims=dims, name=names[0]), DataArray(y2, dims=dims, name=names[1]),
                                       target_crs, xy_res=xy_res, xy_bbox=xy_bbox,
                                       tile_size=tile_size if tile_size is not None else self.tile_size,
                                       tolerance=tolerance)

    @classmethod
    def from_coords(cls, x_coords, y_coords, crs, *, xy_res=None, xy_bbox=None, tile_size=None,
                    tolerance: float = DEFAULT_TOLERANCE, xy_var_names=None, xy_dim_names=None) -> "GridMapping":
        """base.py:803-837 -> coords.py:99-337 (numpy, eager)."""
        crs = normalize_crs(crs)
        x_name = getattr(x_coords, "name", None)
        y_name = getattr(y_coords, "name", None)
        x_dims = getattr(x_coords, "dims", None)
        y_dims = getattr(y_coords, "dims", None)
        x = np.asarray(getattr(x_coords, "values", x_coords))
        y = np.asarray(getattr(y_coords, "values", y_coords))
        if x.ndim not in (1, 2):
            raise ValueError("x_coords and y_coords must be either 1D or 2D arrays")
        if not isinstance(tolerance, float):
            raise TypeError(f"tolerance must be an instance of {float}, was {type(tolerance)}")
        if not tolerance > 0.0:
            raise ValueError("tolerance must be greater zero")
        if xy_var_names is None:
            xy_var_names = (str(x_name), str(y_name)) if x_name and y_name else (
                ("lon", "lat") if crs.is_geographic else ("x", "y"))
        tile_size = _normalize_int_pair(tile_size, default=None)
        is_lon_360 = bool(np.any(x > 180)) if crs.is_geographic else None

        if x.ndim == 1:
            if not (x.size >= 2 and y.size >= 2):
                raise ValueError("sizes of x_coords and y_coords 1D arrays must be >= 2")
            size = x.size, y.size
            x_dim = x_dims[0] if x_dims else xy_var_names[0]
            y_dim = y_dims[0] if y_dims else xy_var_names[1]
            x_diff, y_diff = _abs_no_zero(np.diff(x)), _abs_no_zero(np.diff(y))
            if not is_lon_360 and crs.is_geographic and np.nanmax(x_diff) > 180:
                x = to_lon_360(x)
                x_diff = _abs_no_zero(np.diff(x))
                is_lon_360 = True
            is_regular = None
            if xy_res is not None:
                x_res, y_res = _normalize_number_pair(xy_res)
            else:
                x_res, y_res = x_diff[0], y_diff[0]
                is_regular = bool(np.allclose(x_diff, x_res, atol=tolerance) and np.allclose(y_diff, y_res, atol=tolerance))
                if is_regular:
                    x_res = round_to_fraction(float(x_res), 5, 0.25)
                    y_res = round_to_fraction(float(y_res), 5, 0.25)
                else:
                    x_res = round_to_fraction(float(np.nanmedian(x_diff)), 2, 0.5)
                    y_res = round_to_fraction(float(np.nanmedian(y_diff)), 2, 0.5)
            is_j_axis_up = bool(y[0] < y[-1])
        else:
            if x.shape != y.shape:
                raise ValueError("shapes of x_coords and y_coords 2D arrays must be equal")
            height, width = x.shape
            size = width, height
            if x_dims:
                y_dim, x_dim = x_dims
            else:
                y_dim, x_dim = "y", "x"
            x_x_diff, x_y_diff = _abs_no_nan(np.diff(x[0, :])), _abs_no_nan(np.diff(x[:, 0]))
            y_x_diff, y_y_diff = _abs_no_nan(np.diff(y[0, :])), _abs_no_nan(np.diff(y[:, 0]))
            if not is_lon_360 and crs.is_geographic and (np.max(x_x_diff) > 180 or np.max(x_y_diff) > 180):
                x = to_lon_360(x)
                x_x_diff, x_y_diff = _abs_no_nan(np.diff(x[0, :])), _abs_no_nan(np.diff(x[:, 0]))
                is_lon_360 = True
            if xy_res is not None:
                x_res, y_res = _normalize_number_pair(xy_res)
            else:
                x_res, y_res = x_x_diff[0], y_y_diff[0]
            is_regular = bool(np.allclose(x_x_diff, x_res, atol=tolerance) and np.allclose(y_y_diff, y_res, atol=tolerance)
                              and np.allclose(x_y_diff, 0, atol=tolerance) and np.allclose(y_x_diff, 0, atol=tolerance))
            if not is_regular and xy_res is None:
                x_res = y_res = _estimate_resolution_2d(x, y, crs.is_geographic)
            is_j_axis_up = bool(np.all(y[0, :] < y[-1, :]))

        if not (x_res > 0 and y_res > 0):
            raise RuntimeError("internal error: x_res and y_res could not be determined")
        x_res, y_res = _to_int_or_float(x_res), _to_int_or_float(y_res)
        if xy_bbox is None:
            xh, yh = x_res / 2, y_res / 2
            # xarray's .min() / .max() skip NaN (coords.py:272-281): NaN-padded swath edges are legal input
            x_min = _to_int_or_float(float(np.nanmin(x[..., 0])) - xh)
            x_max = _to_int_or_float(float(np.nanmax(x[..., -1])) + xh)
            if is_j_axis_up:
                y_min = _to_int_or_float(float(np.nanmin(y[0, ...])) - yh)
                y_max = _to_int_or_float(float(np.nanmax(y[-1, ...])) + yh)
            else:
                y_min = _to_int_or_float(float(np.nanmin(y[-1, ...])) - yh)
                y_max = _to_int_or_float(float(np.nanmax(y[0, ...])) + yh)
            xy_bbox = (x_min, y_min, x_max, y_max)
        if xy_dim_names is None:
            xy_dim_names = (str(x_dim), str(y_dim))
        return cls(x_coords=x, y_coords=y, crs=crs, size=size, tile_size=tile_size, xy_bbox=xy_bbox,
                   xy_res=(x_res, y_res), xy_var_names=tuple(xy_var_names), xy_dim_names=tuple(xy_dim_names),
                   is_regular=is_regular, is_lon_360=is_lon_360, is_j_axis_up=is_j_axis_up)

    @property
    def device_coords(self):
        """(x, y) float64 device tensors when the 2-D coordinates live on a GPU, else ``None``."""
        return None if self._x_dev is None else (self._x_dev, self._y_dev)

    @classmethod
    def from_device_coords(cls, x_dev, y_dev, crs, *, xy_res=None, xy_bbox=None, tile_size=None,
                           tolerance: float = DEFAULT_TOLERANCE, xy_var_names=None,
                           xy_dim_names=None) -> "GridMapping":
        """:meth:`from_coords` (coords.py:99-337) for 2-D float64 coordinate images that are resident on
        a GPU and stay there.  What the derivation needs from the whole images -- ``any(x > 180)`` and
        the extreme cell areas behind the resolution estimate -- is reduced by one device pass
        (``xrs_coords_stats``); the regularity test, bounding box and axis direction read the first /
        last rows and columns only, which are the only values copied to the host.  ``x_values`` /
        ``y_values`` fetch the images on demand."""
        import torch

        from . import _dev
        from ._lib import check, load

        lib = load()
        crs = normalize_crs(crs)
        if x_dev.dim() != 2 or x_dev.shape != y_dev.shape or x_dev.dtype != torch.float64 or y_dev.dtype != torch.float64:
            raise ValueError("x_dev and y_dev must be 2-D float64 device tensors of equal shape")
        if x_dev.stride(1) != 1 or y_dev.stride() != x_dev.stride():
            x_dev, y_dev = x_dev.contiguous(), y_dev.contiguous()
        if not isinstance(tolerance, float):
            raise TypeError(f"tolerance must be an instance of {float}, was {type(tolerance)}")
        if not tolerance > 0.0:
            raise ValueError("tolerance must be greater zero")
        height, width = x_dev.shape
        if xy_var_names is None:
            xy_var_names = ("lon", "lat") if crs.is_geographic else ("x", "y")
        tile_size = _normalize_int_pair(tile_size, default=None)

        def stats():
            out = torch.empty(4, dtype=torch.int64, device=x_dev.device)
            check(lib.xrs_coords_stats(_dev.ptr(x_dev), _dev.ptr(y_dev), height, width, x_dev.stride(0),
                                       int(bool(crs.is_geographic)), _dev.ptr(out), _dev.stream_ptr(x_dev.device)),
                  "xrs_coords_stats")
            raw = _dev.to_host(out).view(np.uint64)
            areas = raw[1:3].copy().view(np.float64)
            return bool(raw[0]), (float(areas[0]) if raw[1] != np.uint64(0xFFFFFFFFFFFFFFFF) else math.nan,
                                  float(areas[1]) if raw[2] != 0 else math.nan)

        def edges():
            parts = [x_dev[0], x_dev[-1], y_dev[0], y_dev[-1], x_dev[:, 0], x_dev[:, -1], y_dev[:, 0], y_dev[:, -1]]
            flat = _dev.to_host(torch.cat([p.reshape(-1) for p in parts]))
            out, pos = [], 0
            for n in (width,) * 4 + (height,) * 4:
                out.append(flat[pos:pos + n])
                pos += n
            return out

        any_gt_180, areas = stats()
        x_r0, x_r1, y_r0, y_r1, x_c0, x_c1, y_c0, y_c1 = edges()
        is_lon_360 = any_gt_180 if crs.is_geographic else None
        x_x_diff, x_y_diff = _abs_no_nan(np.diff(x_r0)), _abs_no_nan(np.diff(x_c0))
        y_x_diff, y_y_diff = _abs_no_nan(np.diff(y_r0)), _abs_no_nan(np.diff(y_c0))
        if not is_lon_360 and crs.is_geographic and (np.max(x_x_diff) > 180 or np.max(x_y_diff) > 180):
            x_dev = x_dev.clone()
            check(lib.xrs_lon_360(_dev.ptr(x_dev), height, width, x_dev.stride(0), _dev.stream_ptr(x_dev.device)),
                  "xrs_lon_360")
            _, areas = stats()
            x_r0, x_r1, y_r0, y_r1, x_c0, x_c1, y_c0, y_c1 = edges()
            x_x_diff, x_y_diff = _abs_no_nan(np.diff(x_r0)), _abs_no_nan(np.diff(x_c0))
            is_lon_360 = True
        if xy_res is not None:
            x_res, y_res = _normalize_number_pair(xy_res)
        else:
            x_res, y_res = x_x_diff[0], y_y_diff[0]
        is_regular = bool(np.allclose(x_x_diff, x_res, atol=tolerance) and np.allclose(y_y_diff, y_res, atol=tolerance)
                          and np.allclose(x_y_diff, 0, atol=tolerance) and np.allclose(y_x_diff, 0, atol=tolerance))
        if not is_regular and xy_res is None:
            # coords.py:253-264 from the extreme areas
            xy = 0.7 * math.sqrt(areas[0]) + 0.3 * math.sqrt(areas[1])
            if crs.is_geographic:
                xy = math.degrees(xy / _ER)
            x_res = y_res = float(round_to_fraction(xy, digits=1, resolution=0.5))
        is_j_axis_up = bool(np.all(y_r0 < y_r1))
        if not (x_res > 0 and y_res > 0):
            raise RuntimeError("internal error: x_res and y_res could not be determined")
        x_res, y_res = _to_int_or_float(x_res), _to_int_or_float(y_res)
        if xy_bbox is None:
            xh, yh = x_res / 2, y_res / 2
            y_first, y_last = (y_r0, y_r1) if is_j_axis_up else (y_r1, y_r0)
            xy_bbox = (_to_int_or_float(float(np.nanmin(x_c0)) - xh), _to_int_or_float(float(np.nanmin(y_first)) - yh),
                       _to_int_or_float(float(np.nanmax(x_c1)) + xh), _to_int_or_float(float(np.nanmax(y_last)) + yh))
        if xy_dim_names is None:
            xy_dim_names = ("x", "y")
        gm = cls(crs=crs, size=(width, height), tile_size=tile_size, xy_bbox=xy_bbox, xy_res=(x_res, y_res),
                 xy_var_names=tuple(xy_var_names), xy_dim_names=tuple(xy_dim_names), is_regular=is_regular,
                 is_lon_360=is_lon_360, is_j_axis_up=is_j_axis_up)
        gm._x_dev, gm._y_dev = x_dev, y_dev
        return gm

    @classmethod
    def from_dataset(cls, dataset: Any, *, crs=None, tile_size=None, prefer_is_regular: bool = True, prefer_crs=None,
                     emit_warnings: bool = False, tolerance: float = DEFAULT_TOLERANCE) -> "GridMapping":
        """base.py:760-801 -> dataset.py:34-112: CF grid-mapping discovery (``cfconv.py``), one grid
        mapping per CRS found, then the preference rules when there are several."""
        from .cfconv import get_dataset_grid_mapping_proxies

        ds = from_any(dataset)
        crs = normalize_crs(crs) if crs is not None else None
        prefer_crs = normalize_crs(prefer_crs) if prefer_crs is not None else crs
        proxies = get_dataset_grid_mapping_proxies(ds, emit_warnings=emit_warnings, missing_projected_crs=crs,
                                                   missing_rotated_latitude_longitude_crs=crs,
                                                   missing_latitude_longitude_crs=crs).values()
        gms = [cls.from_coords(gmp.coords.x, gmp.coords.y, gmp.crs, tile_size=tile_size or gmp.tile_size,
                               tolerance=tolerance) for gmp in proxies]
        if len(gms) > 1:
            def both_geographic(gm):
                return gm.crs.is_geographic and prefer_crs.is_geographic

            rules = []
            if prefer_crs is not None and prefer_is_regular is not None:
                rules += [lambda gm: gm.crs == prefer_crs and bool(gm.is_regular) == prefer_is_regular,
                          lambda gm: both_geographic(gm) and bool(gm.is_regular) == prefer_is_regular]
            if prefer_crs is not None:
                rules += [lambda gm: gm.crs == prefer_crs, both_geographic]
            if prefer_is_regular is not None:
                rules += [lambda gm: bool(gm.is_regular) == prefer_is_regular]
            for rule in rules:
                for gm in gms:
                    if rule(gm):
                        return gm
        if gms:
            return gms[0]
        raise ValueError("cannot find any grid mapping in dataset")

    # -- comparisons / assertions --------------------------------------------
    def is_close(self, other: "GridMapping", tolerance: float = DEFAULT_TOLERANCE) -> bool:
        """base.py:839-876."""
        if self is other:
            return True
        if (self.is_j_axis_up == other.is_j_axis_up and self.is_lon_360 == other.is_lon_360
                and self.is_regular == other.is_regular and self.size == other.size
                and self.tile_size == other.tile_size and self.crs == other.crs):
            sxr, syr = self.xy_res
            oxr, oyr = other.xy_res
            if math.isclose(sxr, oxr, abs_tol=tolerance) and math.isclose(syr, oyr, abs_tol=tolerance):
                return all(math.isclose(a, b, abs_tol=tolerance) for a, b in zip(self.xy_bbox, other.xy_bbox))
        return False

    @classmethod
    def assert_regular(cls, value: Any, name: str = None):
        if not isinstance(value, GridMapping):
            raise TypeError(f"{name or 'value'} must be an instance of {GridMapping}, was {type(value)}")
        if not value.is_regular:
            raise ValueError(f"{name or 'value'} must be a regular grid mapping")

    def _assert_regular(self):
        if not self.is_regular:
            raise NotImplementedError("Operation not implemented for non-regular grid mappings")

    def _repr_markdown_(self) -> str:
        """Notebook representation, base.py:890-913: one bullet per property; unknown flags say so and the
        resolution of a non-regular mapping is marked as an estimate."""
        def flag(v):
            return "_unknown_" if v is None else v

        rows = [("is_regular", flag(self.is_regular)), ("is_j_axis_up", flag(self.is_j_axis_up)),
                ("is_lon_360", flag(self.is_lon_360)), ("crs", self.crs),
                ("xy_res", repr(self.xy_res) + ("" if self.is_regular else "  _estimated_")),
                ("xy_bbox", self.xy_bbox), ("ij_bbox", self.ij_bbox), ("xy_dim_names", self.xy_dim_names),
                ("xy_var_names", self.xy_var_names), ("size", self.size), ("tile_size", self.tile_size)]
        return "\n".join([f"class: **{type(self).__name__}**"] + [f"* {k}: {v}" for k, v in rows])

    def __repr__(self):
        return (f"GridMapping(size={self.size}, tile_size={self.tile_size}, xy_bbox={self.xy_bbox}, "
                f"xy_res={self.xy_res}, crs={self.crs.name!r}, is_regular={self.is_regular}, "
                f"is_j_axis_up={self.is_j_axis_up}, is_lon_360={self.is_lon_360})")


def _abs_no_zero(a):
    a = np.fabs(np.asarray(a, dtype=np.float64))
    return np.where(np.isclose(a, 0), np.nan, a)


def _abs_no_nan(a):
    a = np.fabs(np.asarray(a, dtype=np.float64))
    return np.where(np.logical_or(np.isnan(a), np.isclose(a, 0)), 0, a)


def _estimate_resolution_2d(x: np.ndarray, y: np.ndarray, is_geographic: bool) -> float:
    """coords.py:226-264: resolution of an irregular 2-D grid from cell areas."""
    x_x = _abs_no_nan(np.diff(x, axis=1))
    x_y = _abs_no_nan(np.diff(x, axis=0))
    y_x = _abs_no_nan(np.diff(y, axis=1))
    y_y = _abs_no_nan(np.diff(y, axis=0))
    x_x = np.concatenate([x_x, x_x[:, -1:]], axis=1)
    y_x = np.concatenate([y_x, y_x[:, -1:]], axis=1)
    x_y = np.concatenate([x_y, x_y[-1:, :]], axis=0)
    y_y = np.concatenate([y_y, y_y[-1:, :]], axis=0)
    x_abs = np.sqrt(np.square(x_x) + np.square(x_y))
    y_abs = np.sqrt(np.square(y_x) + np.square(y_y))
    if is_geographic:
        x_r, y_r = np.radians(x_abs), np.radians(y_abs)
        x_abs = _ER * np.cos(x_r) * y_r
        y_abs = _ER * y_r
    areas = (x_abs * y_abs).flatten()
    areas = np.where(areas > 0, areas, np.nan)
    res_min = math.sqrt(areas[np.nanargmin(areas)])
    res_max = math.sqrt(areas[np.nanargmax(areas)])
    xy_res = 0.7 * res_min + 0.3 * res_max
    if is_geographic:
        xy_res = math.degrees(xy_res / _ER)
    return float(round_to_fraction(xy_res, digits=1, resolution=0.5))

"""GridMapping.from_coords against the expectations of the reference's
tests/gridmapping/test_coords.py (sizes, resolutions, bounding boxes, regularity, axis direction,
longitude convention).  These are what rectify_dataset derives its target grid from when the caller
passes none (rectify.py:112-118)."""

import numpy as np
import pytest

from xcube_resampling_b200.crs import CRS
from xcube_resampling_b200.dataset import DataArray
from xcube_resampling_b200.gridmapping import GridMapping

GEO = CRS.from_epsg(4326)
Y_DOWN = np.linspace(4.5, -4.5, 10)


def _gm_1d(x, y=Y_DOWN, **kw):
    return GridMapping.from_coords(DataArray(np.asarray(x, dtype=np.float64), dims="lon"),
                                   DataArray(np.asarray(y, dtype=np.float64), dims="lat"), GEO, **kw)


def _props(gm):
    return gm.size, gm.tile_size, gm.xy_res, gm.xy_bbox, gm.is_regular, gm.is_j_axis_up, gm.is_lon_360


def test_1d_axis_directions():
    """test_coords.py:38-74."""
    assert _props(_gm_1d(np.linspace(1.5, 8.5, 8))) == ((8, 10), (8, 10), (1, 1), (1, -5, 9, 5), True, False, False)
    assert _props(_gm_1d(np.linspace(1.5, 8.5, 8), np.linspace(-4.5, 4.5, 10))) == \
        ((8, 10), (8, 10), (1, 1), (1, -5, 9, 5), True, True, False)


def test_1d_lon_360_and_antimeridian():
    """test_coords.py:76-106: 0..360 longitudes, and -180..180 longitudes that jump at the antimeridian."""
    want = ((8, 10), (8, 10), (1, 1), (177, -5, 185, 5), True, False, True)
    lon = np.linspace(177.5, 184.5, 8)
    assert _props(_gm_1d(lon)) == want
    assert _props(_gm_1d(np.where(lon > 180, lon - 360, lon))) == want
    assert _props(_gm_1d(lon, tile_size=(5, 3))) == (want[0], (5, 3)) + want[2:]


def test_1d_slightly_irregular_x():
    """test_coords.py:138-153."""
    gm = _gm_1d([1.5, 2.5, 3.5, 4.5, 5.49, 6.5, 7.5, 8.5])
    assert _props(gm) == ((8, 10), (8, 10), (1, 1), (1, -5, 9, 5), False, False, False)


def test_1d_xy_coords_and_names():
    """test_coords.py:155-168."""
    gm = _gm_1d(np.linspace(1.5, 8.5, 8))
    xy = gm.xy_coords
    assert xy.dims == ("coord", "lat", "lon") and xy.shape == (2, 10, 8)
    assert gm.xy_var_names == ("lon", "lat") and gm.xy_dim_names == ("lon", "lat")


X_2D = [[10.0, 10.1, 10.2, 10.3], [10.1, 10.2, 10.3, 10.4], [10.2, 10.3, 10.4, 10.5]]
Y_2D = [[52.0, 52.2, 52.4, 52.6], [52.2, 52.4, 52.6, 52.8], [52.4, 52.6, 52.8, 53.0]]


def _gm_2d(x, y, dims=("lat", "lon"), **kw):
    return GridMapping.from_coords(DataArray(np.asarray(x, dtype=np.float64), dims=dims),
                                   DataArray(np.asarray(y, dtype=np.float64), dims=dims), GEO, **kw)


def test_2d_sheared_grid():
    """test_coords.py:192-224: area-based resolution estimate (coords.py:226-264) and the bbox from it."""
    gm = _gm_2d(X_2D, Y_2D)
    assert _props(gm) == ((4, 3), (4, 3), (0.3, 0.3), (9.85, 51.85, 10.65, 53.15), False, True, False)


def test_2d_regular():
    """test_coords.py:252-283."""
    gm = _gm_2d([[10.2, 10.3, 10.4, 10.5]] * 3, [[52.4] * 4, [52.6] * 4, [52.8] * 4])
    assert (gm.size, gm.tile_size, gm.is_regular, gm.is_j_axis_up, gm.is_lon_360) == ((4, 3), (4, 3), True, True, False)
    assert gm.x_res == pytest.approx(0.1) and gm.y_res == pytest.approx(0.2)
    assert (gm.x_min, gm.y_min, gm.x_max, gm.y_max) == pytest.approx((10.15, 52.3, 10.55, 52.9))


def test_2d_antimeridian():
    """test_coords.py:285-313."""
    gm = _gm_2d([[+177.5, +178.5, +179.5, -179.5], [+178.5, +179.5, -179.5, -178.5], [+179.5, -179.5, -178.5, -177.5]],
                [[52.4] * 4, [52.6] * 4, [52.8] * 4])
    assert (gm.size, gm.tile_size, gm.is_regular, gm.is_j_axis_up, gm.is_lon_360) == ((4, 3), (4, 3), False, True, True)
    assert gm.x_res == pytest.approx(0.2) and gm.y_res == pytest.approx(0.2)
    assert gm.xy_bbox == (177.4, 52.3, 182.6, 52.9)


def test_2d_to_regular():
    """test_coords.py:315-328: the 2x2 swath of the rectify tests -> a 4x4 grid of 4 degrees."""
    gm = _gm_2d([[1.0, 6.0], [0.0, 2.0]], [[56.0, 53.0], [52.0, 50.0]], dims=("y", "x")).to_regular()
    want = GridMapping.regular(size=(4, 4), tile_size=(2, 2), xy_min=(-2, 48), xy_res=4.0, crs=GEO)
    assert (gm.size, gm.tile_size, gm.xy_res, gm.xy_bbox, gm.crs) == \
        (want.size, want.tile_size, want.xy_res, want.xy_bbox, want.crs)


def test_to_coords_keeps_the_dtype_of_reused_coordinates():
    """test_coords.py:170-188."""
    gm = GridMapping.regular(size=(10, 6), xy_min=(-2600.0, 1200.0), xy_res=10.0, crs="EPSG:32633")
    cv = gm.to_coords(reuse_coords=False)
    assert cv["x"].values.dtype == np.float64 and cv["y"].values.dtype == np.float64
    gm2 = GridMapping.from_coords(DataArray(cv["x"].values.astype(np.float32), dims="x"),
                                  DataArray(cv["y"].values.astype(np.float32), dims="y"), gm.crs)
    cv2 = gm2.to_coords(xy_var_names=("a", "b"), xy_dim_names=("u", "v"), reuse_coords=True)
    assert cv2["a"].values.dtype == np.float32 and cv2["b"].values.dtype == np.float32


def test_from_dataset_non_regular_cube():
    """tests/gridmapping/test_dataset.py:40-66: 2-D float32 lon/lat variables of a swath."""
    from xcube_resampling_b200.dataset import Dataset

    lon = np.array([[8, 9.3, 10.6, 11.9], [8, 9.2, 10.4, 11.6], [8, 9.1, 10.2, 11.3]], dtype=np.float32)
    lat = np.array([[56, 56.1, 56.2, 56.3], [55, 55.2, 55.4, 55.6], [54, 54.3, 54.6, 54.9]], dtype=np.float32)
    rad = np.random.default_rng(0).random((3, 4))
    ds = Dataset(dict(lon=DataArray(lon, dims=("y", "x")), lat=DataArray(lat, dims=("y", "x")),
                      rad=DataArray(rad, dims=("y", "x"))))
    gm = GridMapping.from_dataset(ds)
    assert (gm.size, gm.tile_size, gm.crs) == ((4, 3), (4, 3), GEO)
    assert (gm.is_regular, gm.is_lon_360, gm.is_j_axis_up) == (False, False, False)
    assert gm.xy_coords.shape == (2, 3, 4) and gm.xy_coords.dims == ("coord", "y", "x")
    assert gm.xy_res == (0.8, 0.8)


def test_from_dataset_explicit_crs():
    """tests/gridmapping/test_dataset.py:68-82."""
    from xcube_resampling_b200.dataset import Dataset

    ds = Dataset(data_vars={"var": (("lat", "lon"), np.random.default_rng(1).random((2, 2)))},
                 coords={"lon": ("lon", np.array([0.0, 1.0])), "lat": ("lat", np.array([0.0, 1.0]))})
    gm = GridMapping.from_dataset(ds, crs="EPSG:4326")
    assert gm.is_regular and gm.crs == GEO


def test_from_coords_skips_nan_in_the_edge_rows_and_columns():
    """xarray's .min() / .max() skip NaN (coords.py:272-281): NaN-padded swath edges must not break
    the bounding box (np.min / np.max raised 'cannot convert float NaN to integer')."""
    from xcube_resampling_b200.synthetic import swath

    lon, lat = swath(40, 30, theta=15.0, seed=1)
    clean = GridMapping.from_coords(lon, lat, "EPSG:4326", xy_res=0.0027, xy_dim_names=("x", "y"))
    for (j, i) in ((0, 3), (29, 0), (10, 0), (5, 39), (29, 39)):
        x, y = lon.copy(), lat.copy()
        x[j, i] = np.nan
        y[j, i] = np.nan
        gm = GridMapping.from_coords(x, y, "EPSG:4326", xy_res=0.0027, xy_dim_names=("x", "y"))
        assert all(np.isfinite(v) for v in gm.xy_bbox)
        assert gm.size == clean.size
        for a, b in zip(gm.xy_bbox, clean.xy_bbox):
            assert abs(a - b) < 0.01

"""GridMapping.from_coords against the expectations of the reference's
tests/gridmapping/test_coords.py (sizes, resolutions, bounding boxes, regularity, axis direction,
longitude convention).  These are what rectify_dataset derives its target grid from when the caller
passes none (rectify.py:112-118)."""

import numpy as np
import pytest

import xcube_resampling_b200 as xrs
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


def _s2plus_dataset():
    """The two-grid-mapping Sentinel-2 sample of the reference's tests/sampledata.py:211-290: regular UTM x / y
    (CF transverse_mercator variable) AND irregular 2-D lon / lat of the same 5x5 pixels."""
    x = xrs.DataArray(310005.0 + 10.0 * np.arange(5), dims=["x"],
                      attrs=dict(units="m", standard_name="projection_x_coordinate"))
    y = xrs.DataArray(5689995.0 - 10.0 * np.arange(5), dims=["y"],
                      attrs=dict(units="m", standard_name="projection_y_coordinate"))
    lon = xrs.DataArray(np.array([[0.272763, 0.272906, 0.273050, 0.273193, 0.273336],
                                  [0.272768, 0.272911, 0.273055, 0.273198, 0.273342],
                                  [0.272773, 0.272917, 0.273060, 0.273204, 0.273347],
                                  [0.272779, 0.272922, 0.273066, 0.273209, 0.273352],
                                  [0.272784, 0.272927, 0.273071, 0.273214, 0.273358]]), dims=["y", "x"],
                        attrs=dict(units="degrees_east", standard_name="longitude"))
    lat = xrs.DataArray(np.array([[51.329464, 51.329464, 51.329468, 51.32947, 51.329475],
                                  [51.329372, 51.329376, 51.32938, 51.329384, 51.329388],
                                  [51.329285, 51.329285, 51.32929, 51.329292, 51.329296],
                                  [51.329193, 51.329197, 51.32920, 51.329205, 51.329205],
                                  [51.329100, 51.329105, 51.32911, 51.329113, 51.329117]]), dims=["y", "x"],
                        attrs=dict(units="degrees_north", standard_name="latitude"))
    rrs = xrs.DataArray(np.full((5, 5), 0.014), dims=["y", "x"],
                        attrs=dict(units="sr-1", grid_mapping="transverse_mercator"))
    tm = xrs.DataArray(np.array([0xFFFFFFFF], dtype=np.uint32), dims=["dim_0"], attrs=dict(
        grid_mapping_name="transverse_mercator", scale_factor_at_central_meridian=0.9996,
        longitude_of_central_meridian=3.0, latitude_of_projection_origin=0.0, false_easting=500000.0,
        false_northing=0.0, semi_major_axis=6378137.0, inverse_flattening=298.257223563))
    return xrs.Dataset(dict(rrs_443=rrs, rrs_665=rrs, transverse_mercator=tm), coords=dict(x=x, y=y, lon=lon, lat=lat))


@pytest.mark.parametrize("prefer, projected, regular", [
    ({}, True, True), (dict(prefer_is_regular=True), True, True), (dict(prefer_is_regular=False), False, False),
    (dict(prefer_crs=GEO), False, False), (dict(prefer_crs=GEO, prefer_is_regular=True), False, False)])
def test_from_dataset_picks_among_two_grid_mappings(prefer, projected, regular):
    # tests/gridmapping/test_dataset.py:111-141
    gm = xrs.GridMapping.from_dataset(_s2plus_dataset(), tolerance=1e-6, **prefer)
    assert gm.crs.is_projected is projected and gm.is_regular is regular and gm.size == (5, 5)
    if projected:
        assert gm.xy_res == (10, 10) and gm.xy_bbox == (310000, 5689950, 310050, 5690000)
        assert gm.crs == xrs.CRS.from_epsg(32631)  # found from the CF parameters alone


def test_from_dataset_without_any_grid_mapping():
    # tests/gridmapping/test_dataset.py:143-146
    with pytest.raises(ValueError) as e:
        xrs.GridMapping.from_dataset(xrs.Dataset())
    assert str(e.value) == "cannot find any grid mapping in dataset"

"""The reference's tests/test_spatial.py and the CRS-changing rectify cases
(tests/test_rectify.py:461-500) through the B200 entry points: dispatch rules, rectify with a
forward coordinate transform on the device (rectify.py:182-231) and the rectify pre-downscale
(rectify.py:234-260)."""

import logging

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
nan = np.nan


@pytest.fixture(scope="module")
def xrs():
    import torch

    assert torch.cuda.is_available()
    import xcube_resampling_b200 as pkg

    return pkg


def _irregular_4x4(xrs):
    """tests/sampledata.py:175-211."""
    lon = np.array([[1.0, 2.0, 3.0, 4.0], [0.0, 1.0, 2.0, 3.0], [-1.0, 0.0, 1.0, 2.0], [-2.0, -1.0, 0.0, 1.0]])
    lat = np.array([[56.0, 55.0, 54.0, 53.0], [55.0, 54.0, 53.0, 52.0], [54.0, 53.0, 52.0, 51.0],
                    [53.0, 52.0, 51.0, 50.0]])
    rad = np.arange(1.0, 17.0).reshape(4, 4)
    return xrs.Dataset(data_vars=dict(rad=(("y", "x"), rad)),
                       coords=dict(lon=(("y", "x"), lon), lat=(("y", "x"), lat)))


def _irregular_2x2(xrs):
    """tests/sampledata.py:29-39."""
    return xrs.Dataset(data_vars=dict(rad=(("y", "x"), np.array([[1.0, 2.0], [3.0, 4.0]]))),
                       coords=dict(lon=(("y", "x"), np.array([[1.0, 6.0], [0.0, 2.0]])),
                                   lat=(("y", "x"), np.array([[56.0, 53.0], [52.0, 50.0]]))))


def _regular_8x6(xrs):
    """tests/sampledata.py:60-83."""
    res = 0.1
    refl = np.array([[0, 1, 0, 2, 0, 3, 0, 4], [2, 0, 3, 0, 4, 0, 1, 0], [0, 4, 0, nan, 0, 2, 0, 3],
                     [1, 0, 2, 0, 3, 0, 4, 0], [0, 3, 0, 4, 0, 1, 0, 2], [4, 0, 1, 0, 2, 0, 3, 0]], dtype=np.float64)
    return xrs.Dataset(data_vars=dict(refl=(("lat", "lon"), refl)),
                       coords=dict(lon=("lon", 50.0 + res * np.arange(0, 8) + 0.5 * res),
                                   lat=("lat", 10.6 - res * np.arange(0, 6) - 0.5 * res)))


def _utm_5x5(xrs):
    """tests/sampledata.py:95-109."""
    return xrs.Dataset(
        data_vars=dict(band_1=xrs.DataArray(np.arange(25).reshape(5, 5), dims=("y", "x"),
                                            attrs=dict(grid_mapping="spatial_ref"))),
        coords=dict(x=("x", np.arange(565300.0, 565800.0, 100.0)), y=("y", np.arange(5934300.0, 5933800.0, -100.0)),
                    spatial_ref=xrs.DataArray(np.array(0), dims=(), attrs=xrs.CRS.from_epsg(32632).to_cf())))


def test_ref_rectify_different_crs(xrs):
    """tests/test_rectify.py:461-479: geographic 2-D coordinates, LAEA target."""
    tgt = xrs.GridMapping.regular(size=(3, 3), xy_min=(3600000, 3200000), xy_res=100000, crs="epsg:3035")
    out = xrs.rectify_dataset(_irregular_4x4(xrs), target_gm=tgt, interp_methods=0)
    np.testing.assert_almost_equal(out.x.values, [3650000.0, 3750000.0, 3850000.0])
    np.testing.assert_almost_equal(out.y.values, [3450000.0, 3350000.0, 3250000.0])
    np.testing.assert_almost_equal(out.rad.values, [[10.0, 6.0, 3.0], [10.0, 7.0, 3.0], [11.0, 11.0, 8.0]])


def test_ref_spatial_affine(xrs):
    """tests/test_spatial.py:25-49."""
    ds = _regular_8x6(xrs)
    gm = xrs.GridMapping.from_dataset(ds)
    tgt = xrs.GridMapping.regular((3, 3), (50.0, 10.0), 0.1, gm.crs)
    out = xrs.resample_in_space(ds, tgt, interp_methods=1)
    assert out.refl.shape == (3, 3)
    np.testing.assert_almost_equal(out.refl.values, [[1, 0, 2], [0, 3, 0], [4, 0, 1]])


def test_ref_spatial_rectify_and_downscale(xrs):
    """tests/test_spatial.py:51-77."""
    tgt = xrs.GridMapping.regular(size=(2, 2), xy_min=(-1, 51), xy_res=2, crs=xrs.CRS_WGS84)
    out = xrs.resample_in_space(_irregular_4x4(xrs), target_gm=tgt, interp_methods=0)
    np.testing.assert_almost_equal(out.rad.values, [[5, 2], [14, 8]])
    out = xrs.resample_in_space(_irregular_4x4(xrs), target_gm=tgt, interp_methods=1)
    np.testing.assert_almost_equal(out.rad.values, [[7.5, 4.5], [12.5, 9.5]])


def test_ref_spatial_rectify_and_upscale(xrs):
    """tests/test_spatial.py:79-96."""
    tgt = xrs.GridMapping.regular(size=(4, 4), xy_min=(-1, 49), xy_res=2, crs=xrs.CRS_WGS84)
    out = xrs.resample_in_space(_irregular_2x2(xrs), target_gm=tgt, interp_methods=0)
    np.testing.assert_almost_equal(out.rad.values, [[nan, nan, nan, nan], [nan, 1.0, 2.0, nan], [3.0, 3.0, 2.0, nan],
                                                    [nan, 4.0, nan, nan]])


def test_ref_spatial_reproject(xrs):
    """tests/test_spatial.py:98-177."""
    cases = [
        (dict(size=(5, 5), xy_min=(4320080, 3382480), xy_res=80, crs="epsg:3035"),
         [[1, 1, 2, 3, 4], [6, 6, 7, 8, 9], [11, 12, 12, 13, 14], [16, 17, 17, 18, 19], [21, 17, 17, 18, 19]]),
        (dict(size=(5, 5), xy_min=(4320080, 3382480), xy_res=20, crs="epsg:3035"),
         [[15, 16, 16, 16, 16], [15, 16, 16, 16, 16], [15, 16, 16, 16, 16], [20, 21, 21, 21, 21], [20, 21, 21, 21, 21]]),
        (dict(size=(5, 5), xy_min=(9.9886, 53.5499), xy_res=0.0006, crs=xrs.CRS_WGS84),
         [[7, 8, 8, 8, 9], [12, 13, 13, 13, 14], [12, 13, 13, 13, 14], [17, 18, 18, 18, 19], [22, 23, 23, 23, 24]]),
        (dict(size=(5, 5), xy_min=(9.9886, 53.5499), xy_res=0.0003, crs=xrs.CRS_WGS84),
         [[12, 12, 12, 13, 13], [17, 17, 17, 18, 18], [17, 17, 17, 18, 18], [22, 17, 17, 18, 18], [22, 22, 22, 23, 23]]),
    ]
    for kw, expected in cases:
        out = xrs.resample_in_space(_utm_5x5(xrs), target_gm=xrs.GridMapping.regular(**kw), interp_methods=0)
        np.testing.assert_array_equal(out.band_1.values, expected)


def test_ref_spatial_logs_and_passthrough(xrs, caplog):
    """tests/test_spatial.py:179-193."""
    ds = _utm_5x5(xrs)
    with caplog.at_level(logging.WARNING, logger="xcube.resampling"):
        out = xrs.resample_in_space(ds)
    assert out is ds
    assert any("If source grid mapping is regular `target_gm` must be given. Source dataset is returned." in m
               for m in caplog.messages)
    out = xrs.resample_in_space(ds, target_gm=xrs.GridMapping.from_dataset(ds))
    assert out is ds


def test_ref_gridmapping_transform(xrs):
    """tests/gridmapping/test_transform.py:35-64: CRS84 3x3 grid -> UTM 32N, 7 decimals."""
    gm = xrs.GridMapping.regular(size=(3, 3), xy_min=(10, 53), xy_res=0.1, crs="CRS84")
    gm_t = gm.transform(crs="EPSG:32632")
    assert gm_t.crs == xrs.CRS.from_epsg(32632)
    assert gm_t.is_regular is False
    assert gm_t.xy_var_names == ("transformed_x", "transformed_y")
    assert gm_t.xy_dim_names == ("lon", "lat")
    xy = gm_t.xy_coords.values
    np.testing.assert_almost_equal(xy[0], np.array([
        [570057.076286, 576728.9360228, 583400.7295284],
        [570220.3304187, 576907.7404859, 583595.0849538],
        [570383.3684844, 577086.3083212, 583789.1831954]]))
    np.testing.assert_almost_equal(xy[1], np.array([
        [5900595.928991, 5900698.5746648, 5900810.5532744],
        [5889471.9033896, 5889574.6540572, 5889686.7472201],
        [5878348.0594403, 5878450.9138481, 5878563.1201969]]))


def test_ref_gridmapping_transform_names_and_noop(xrs):
    """tests/gridmapping/test_transform.py:66-110."""
    gm = xrs.GridMapping.regular(size=(3, 3), xy_min=(10, 53), xy_res=0.1, crs="CRS84")
    gm_t = gm.transform(crs="EPSG:32632", xy_var_names=("x", "y"))
    assert gm_t.xy_var_names == ("x", "y")
    assert gm_t.xy_dim_names == ("lon", "lat")
    assert gm.transform(gm.crs) is gm
    assert gm.transform(crs=gm.crs, xy_var_names=("x", "y")).xy_var_names == ("x", "y")
    # a projected regular grid to both geographic CRSs, and an explicit resolution (bbox from densified edges)
    utm = xrs.GridMapping.regular(size=(20, 10), xy_min=(500000, 5900000), xy_res=100, crs="EPSG:32632")
    for crs in ("CRS84", "EPSG:4326"):
        assert utm.transform(crs).crs.is_geographic
    geo = utm.transform("EPSG:4326", xy_res=0.001)
    assert geo.xy_res == (0.001, 0.001)
    x0, y0, x1, y1 = geo.xy_bbox
    xy = geo.xy_coords.values
    assert x0 < xy[0].min() and xy[0].max() < x1 and y0 < xy[1].min() and xy[1].max() < y1


def test_ref_gridmapping_transform_to_regular(xrs):
    """tests/gridmapping/test_base.py:348-404 with the device point transform."""
    gm = xrs.GridMapping.regular((400, 200), (20, 56), 0.01, "EPSG:4326", tile_size=(200, 200))
    t = gm.transform("EPSG:32633", xy_res=1000)
    assert (t.size, t.tile_size, t.xy_res, t.is_j_axis_up) == ((400, 200), (200, 200), (1000, 1000), False)
    r = t.to_regular()
    assert (r.size, r.tile_size, r.xy_res, r.is_j_axis_up) == ((267, 249), (200, 200), (1000, 1000), False)
    assert r.xy_var_names == ("x", "y") and r.xy_dim_names == ("x", "y")
    r = xrs.GridMapping.regular((1000, 1000), (9.6, 47.6), 0.0002, "EPSG:4326").transform("EPSG:32633").to_regular()
    assert (r.size, r.tile_size, r.is_j_axis_up, r.is_lon_360) == ((827, 1163), (1000, 1000), False, False)


def test_ref_ij_bbox_from_xy_bbox(xrs):
    """tests/gridmapping/test_base.py:456-512: source-index boxes of xy boxes on a global 0.5 deg grid (K0)."""
    gm = xrs.GridMapping.regular((720, 360), (-180.0, -90.0), 0.5, "EPSG:4326", tile_size=(360, 180))
    assert gm.ij_bbox_from_xy_bbox((-180, -90, 180, 90)) == (0, 0, 720, 360)
    assert gm.ij_bbox_from_xy_bbox((-180, -90, 0, 0)) == (0, 180, 360, 360)
    assert gm.ij_bbox_from_xy_bbox((0, 0, 180, 90)) == (360, 0, 720, 180)
    assert gm.ij_bbox_from_xy_bbox((-180, -90, 0, 0), ij_border=1) == (0, 179, 361, 360)
    assert gm.ij_bbox_from_xy_bbox((0, 0, 180, 90), ij_border=1) == (359, 0, 720, 181)
    assert gm.ij_bbox_from_xy_bbox((-190, -100, -170, -80), ij_border=1) == (0, 339, 21, 360)
    assert gm.ij_bbox_from_xy_bbox((-190, -100, -180, -90), ij_border=1) == (-1, -1, -1, -1)
    # a batch of arbitrary boxes (not the tiles of one grid): one K0 pass per box
    boxes = np.array([[-180, -90, 180, 90], [-180, -90, 0, 0], [0, 0, 180, 90], [-180, -90, 0, 0], [0, 0, 180, 90],
                      [-190, -100, -170, -80], [-190, -100, -180, -90]], dtype=np.float32)
    want = np.array([[0, 0, 720, 360], [0, 180, 360, 360], [360, 0, 720, 180], [0, 180, 360, 360], [360, 0, 720, 180],
                     [0, 340, 20, 360], [-1, -1, -1, -1]], dtype=np.int64)
    got = gm.ij_bboxes_from_xy_bboxes(boxes)
    assert got.dtype == np.int64 and np.array_equal(got, want)

"""GridMapping host type: expectations of the reference's tests/gridmapping/test_regular.py,
test_base.py (transforms, tile boxes, derive / scale, is_close) and test_coords.py (1-D / 2-D
coordinate derivation), for the fields and methods the resampling path reads (CPU only)."""

import numpy as np
import pytest

import xcube_resampling_b200 as xrs

GM = xrs.GridMapping
WGS84 = xrs.CRS_WGS84
UTM = "EPSG:32632"  # any projected CRS (the reference uses EPSG:5243)


def _apply(m, p):
    (a, b, c), (d, e, f) = m
    return a * p[0] + b * p[1] + c, d * p[0] + e * p[1] + f


# ---- tests/gridmapping/test_regular.py ---------------------------------------------------------
def test_default_props_and_validation():
    gm = GM.regular((1000, 1000), (10, 53), 0.01, WGS84)
    assert gm.size == (1000, 1000) and gm.tile_size == (1000, 1000)
    assert gm.x_min == 10 and gm.y_min == 53 and gm.xy_res == (0.01, 0.01)
    assert gm.is_regular is True and gm.is_j_axis_up is False
    assert gm.xy_bbox == (10, 53, 20, 63) and gm.is_lon_360 is False
    with pytest.raises(ValueError, match="invalid y_min"):
        GM.regular((1000, 1000), (10, -90.5), 0.01, WGS84)
    with pytest.raises(ValueError, match="invalid size, y_min combination"):
        GM.regular((1000, 1000), (10, 53), 0.1, WGS84)


def test_xy_bbox_anti_meridian():
    gm = GM.regular((2000, 1000), (174.0, -30.0), 0.005, WGS84)
    assert gm.xy_bbox == (174.0, -30.0, 184.0, -25.0) and gm.is_lon_360 is True


def test_derive_and_names():
    gm = GM.regular((1000, 1000), (10, 53), 0.01, WGS84)
    d = gm.derive(tile_size=500, is_j_axis_up=True)
    assert d is not gm and d.size == (1000, 1000) and d.tile_size == (500, 500) and d.is_j_axis_up is True
    assert gm.xy_var_names == ("lon", "lat") and gm.xy_dim_names == ("lon", "lat")
    p = GM.regular((1000, 1000), (10, 53), 0.01, UTM).derive(tile_size=500)
    assert p.xy_var_names == ("x", "y") and p.xy_dim_names == ("x", "y")


def test_xy_coords():
    gm = GM.regular((8, 4), (10, 53), 0.1, WGS84).derive(tile_size=(4, 2))
    xy = gm.xy_coords
    assert xy.dims == ("coord", "lat", "lon") and xy.shape == (2, 4, 8)
    np.testing.assert_almost_equal(xy.values[0], np.tile(10.05 + 0.1 * np.arange(8), (4, 1)))
    np.testing.assert_almost_equal(xy.values[1], np.tile((53.35 - 0.1 * np.arange(4))[:, None], (1, 8)))


def test_ij_and_xy_bboxes():
    gm = GM.regular(size=(2000, 1000), xy_min=(10.0, 20.0), xy_res=0.1, crs=UTM)
    np.testing.assert_array_equal(gm.ij_bboxes, [[0, 0, 2000, 1000]])
    np.testing.assert_almost_equal(gm.xy_bboxes, [[10.0, 20.0, 210.0, 120.0]])
    t = gm.derive(tile_size=500)
    np.testing.assert_array_equal(t.ij_bboxes, [
        [0, 0, 500, 500], [500, 0, 1000, 500], [1000, 0, 1500, 500], [1500, 0, 2000, 500],
        [0, 500, 500, 1000], [500, 500, 1000, 1000], [1000, 500, 1500, 1000], [1500, 500, 2000, 1000]])
    np.testing.assert_almost_equal(t.xy_bboxes, [
        [10.0, 70, 60, 120.0], [60.0, 70, 110, 120.0], [110.0, 70, 160, 120.0], [160.0, 70, 210, 120.0],
        [10.0, 20, 60, 70.0], [60.0, 20, 110, 70.0], [110.0, 20, 160, 70.0], [160.0, 20, 210, 70.0]])
    up = gm.derive(tile_size=500, is_j_axis_up=True)
    np.testing.assert_almost_equal(up.xy_bboxes, [
        [10.0, 20.0, 60.0, 70.0], [60.0, 20.0, 110.0, 70.0], [110.0, 20.0, 160.0, 70.0], [160.0, 20.0, 210.0, 70.0],
        [10.0, 70.0, 60.0, 120.0], [60.0, 70.0, 110.0, 120.0], [110.0, 70.0, 160.0, 120.0],
        [160.0, 70.0, 210.0, 120.0]])


def test_to_coords():
    gm = GM.regular(size=(10, 6), xy_min=(-2600.0, 1200.0), xy_res=10.0, crs=UTM)
    cv = gm.to_coords(xy_var_names=("x", "y"))
    assert cv["x"].shape == (10,) and cv["y"].shape == (6,)
    np.testing.assert_almost_equal(cv["x"].values[[0, -1]], [-2595.0, -2505.0])
    np.testing.assert_almost_equal(cv["y"].values[[0, -1]], [1255.0, 1205.0])
    np.testing.assert_almost_equal(cv["x_bnds"].values[[0, -1]], [[-2600.0, -2590.0], [-2510.0, -2500.0]])
    np.testing.assert_almost_equal(cv["y_bnds"].values[[0, -1]], [[1260.0, 1250.0], [1210.0, 1200.0]])
    up = gm.derive(is_j_axis_up=True).to_coords(xy_var_names=("x", "y"))
    np.testing.assert_almost_equal(up["y"].values[[0, -1]], [1205.0, 1255.0])
    np.testing.assert_almost_equal(up["y_bnds"].values[[0, -1]], [[1200.0, 1210.0], [1250.0, 1260.0]])
    am = GM.regular(size=(10, 10), xy_min=(172.0, 53.0), xy_res=2.0, crs=WGS84).to_coords(xy_var_names=("lon", "lat"))
    np.testing.assert_almost_equal(am["lon"].values[[0, -1]], [173.0, -169.0])
    np.testing.assert_almost_equal(am["lon_bnds"].values[[0, -1]], [[172.0, 174.0], [-170.0, -168.0]])


def test_to_regular():
    gm = GM.regular((1000, 1000), (10, 53), 0.01, WGS84)
    for kw, tile, up in (({}, (1000, 1000), False), (dict(tile_size=500), (500, 500), False),
                         (dict(is_j_axis_up=True), (1000, 1000), True)):
        r = gm.to_regular(**kw)
        assert r.size == (1000, 1000) and r.tile_size == tile and r.crs == WGS84 and r.xy_res == (0.01, 0.01)
        assert r.is_j_axis_up is up


# ---- tests/gridmapping/test_base.py:174-252 -------------------------------------------------------
def test_ij_to_xy_and_back_transforms():
    gm = GM.regular((1200, 1200), (0, 0), 1, UTM)
    assert gm.ij_to_xy_transform == ((1, 0, 0), (0.0, -1, 1200))
    assert _apply(gm.ij_to_xy_transform, (1024, 1200 - 1024)) == pytest.approx((1024, 1024))
    assert gm.xy_to_ij_transform == ((1, 0, 0), (0.0, -1, 1200))
    gm = GM.regular((1440, 720), (-180, -90), 0.25, WGS84)
    assert gm.ij_to_xy_transform == ((0.25, 0.0, -180.0), (0.0, -0.25, 90.0))
    assert gm.xy_to_ij_transform == ((4.0, 0.0, 720.0), (0.0, -4.0, 360.0))
    assert _apply(gm.xy_to_ij_transform, (180, 90)) == pytest.approx((1440, 0))
    up = GM.regular((1440, 720), (-180, -90), 0.25, WGS84, is_j_axis_up=True)
    assert up.ij_to_xy_transform == ((0.25, 0.0, -180.0), (0.0, 0.25, -90.0))
    assert up.xy_to_ij_transform == ((4.0, 0.0, 720.0), (0.0, 4.0, 360.0))


def test_ij_transform_to_and_from():
    gm1 = GM.regular((1440, 720), (-180, -90), 0.25, WGS84, is_j_axis_up=True)
    gm2 = GM.regular((1000, 1000), (10, 50), 0.025, WGS84, is_j_axis_up=True)
    assert gm1.ij_transform_to(gm2) == ((10.0, 0.0, -7600.0), (0.0, 10.0, -5600.0))
    assert gm2.ij_transform_from(gm1) == ((10.0, 0.0, -7600.0), (0.0, 10.0, -5600.0))
    assert gm2.ij_transform_to(gm1) == ((0.1, 0.0, 760.0), (0.0, 0.1, 560.0))
    assert gm1.ij_transform_from(gm2) == ((0.1, 0.0, 760.0), (0.0, 0.1, 560.0))


def test_scale_and_is_close():
    gm = GM.regular((720, 360), (-180, -90), 0.5, WGS84, tile_size=(360, 180))
    s = gm.scale((0.25, 0.5))
    assert s.size == (180, 180) and s.xy_res == (2.0, 1.0) and s.xy_bbox == gm.xy_bbox
    assert gm.is_close(GM.regular((720, 360), (-180, -90), 0.5, WGS84, tile_size=(360, 180)))
    assert gm.is_close(GM.regular((720, 360), (-180 + 1e-7, -90), 0.5, WGS84, tile_size=(360, 180)))
    assert not gm.is_close(GM.regular((720, 360), (-180, -90), 0.5, WGS84))  # tile size is compared too
    assert not gm.is_close(GM.regular((720, 360), (-170, -90), 0.5, WGS84, tile_size=(360, 180)))
    assert not gm.is_close(GM.regular((720, 360), (-180, -90), 0.5, WGS84, tile_size=(360, 180), is_j_axis_up=True))
    with pytest.raises(ValueError):
        GM.assert_regular(GM.from_coords(np.array([[1.0, 6.0], [0.0, 2.0]]), np.array([[56.0, 53.0], [52.0, 50.0]]),
                                         WGS84), name="target_gm")


# ---- tests/gridmapping/test_coords.py ------------------------------------------------------------
def test_from_coords_1d_regular_and_j_axis():
    x = np.linspace(10.05, 10.95, 10)
    y = np.linspace(53.95, 53.05, 10)
    gm = GM.from_coords(xrs.DataArray(x, dims="lon", name="lon"), xrs.DataArray(y, dims="lat", name="lat"), WGS84)
    assert gm.is_regular and not gm.is_j_axis_up and gm.size == (10, 10)
    assert gm.xy_res == pytest.approx((0.1, 0.1))
    assert gm.xy_bbox == pytest.approx((10.0, 53.0, 11.0, 54.0))
    up = GM.from_coords(xrs.DataArray(x, dims="lon", name="lon"), xrs.DataArray(y[::-1], dims="lat", name="lat"), WGS84)
    assert up.is_j_axis_up and up.xy_bbox == pytest.approx((10.0, 53.0, 11.0, 54.0))


def test_from_coords_2d_irregular():
    lon = np.array([[1.0, 6.0], [0.0, 2.0]])
    lat = np.array([[56.0, 53.0], [52.0, 50.0]])
    gm = GM.from_coords(xrs.DataArray(lon, dims=("y", "x"), name="lon"), xrs.DataArray(lat, dims=("y", "x"), name="lat"),
                        WGS84)
    assert gm.is_regular is False and gm.size == (2, 2) and gm.xy_dim_names == ("x", "y")
    assert gm.xy_var_names == ("lon", "lat")
    assert gm.x_res == gm.y_res and gm.x_res > 0
    reg = gm.to_regular()
    assert reg.is_regular and reg.crs == WGS84 and reg.xy_res == gm.xy_res


def test_from_coords_antimeridian_2d():
    lon = np.array([[+179.0, -176.0], [+178.0, +180.0]])
    lat = np.array([[56.0, 53.0], [52.0, 50.0]])
    gm = GM.from_coords(lon, lat, WGS84)
    assert gm.is_lon_360 is True
    assert gm.xy_bbox[2] > 180.0


def test_from_coords_validation():
    with pytest.raises(ValueError):
        GM.from_coords(np.zeros((2, 2, 2)), np.zeros((2, 2, 2)), WGS84)
    with pytest.raises(ValueError):
        GM.from_coords(np.zeros((2, 3)), np.zeros((3, 2)), WGS84)
    with pytest.raises(ValueError):
        GM.from_coords(np.array([1.0]), np.array([1.0]), WGS84)


# ---- tests/gridmapping/test_base.py:90-172, 514-end ---------------------------------------------------
_BASE = dict(size=(720, 360), tile_size=(360, 180), xy_bbox=(-180.0, -90.0, 180.0, 90.0), xy_res=(0.5, 0.5),
             crs=WGS84, xy_var_names=("x", "y"), xy_dim_names=("x", "y"), is_regular=True, is_lon_360=False,
             is_j_axis_up=False)


def test_constructor_properties_and_tile_boxes():
    gm = GM(**_BASE)
    assert (gm.size, gm.width, gm.height) == ((720, 360), 720, 360)
    assert (gm.is_tiled, gm.tile_size, gm.tile_width, gm.tile_height) == (True, (360, 180), 360, 180)
    assert gm.ij_bbox == (0, 0, 720, 360) and gm.xy_bbox == (-180.0, -90.0, 180.0, 90.0)
    assert (gm.x_min, gm.y_min, gm.x_max, gm.y_max) == (-180.0, -90.0, 180.0, 90.0)
    assert (gm.xy_res, gm.x_res, gm.y_res) == ((0.5, 0.5), 0.5, 0.5)
    assert gm.crs == WGS84 and gm.spatial_unit_name == "degree"
    assert GM.regular((10, 10), (0, 0), 10, UTM).spatial_unit_name == "metre"
    np.testing.assert_array_equal(gm.ij_bboxes, [[0, 0, 360, 180], [360, 0, 720, 180], [0, 180, 360, 360],
                                                 [360, 180, 720, 360]])
    np.testing.assert_array_equal(gm.xy_bboxes, [[-180.0, 0.0, 0.0, 90.0], [0.0, 0.0, 180.0, 90.0],
                                                 [-180.0, -90.0, 0.0, 0.0], [0.0, -90.0, 180.0, 0.0]])


@pytest.mark.parametrize("override, message", [
    (dict(size=(360, 1)), "invalid size"),
    (dict(size=(360,)), "not enough values to unpack (expected 2, got 1)"),
    (dict(size=None), "size must be an int or a sequence of two ints"),
    (dict(tile_size=0), "invalid tile_size"),
    (dict(xy_res=-0.1), "invalid xy_res")])
def test_constructor_rejects(override, message):
    with pytest.raises(ValueError) as e:
        GM(**dict(_BASE, **override))
    assert str(e.value) == message


def test_constructor_scalars_and_untiled():
    gm = GM(**dict(_BASE, size=360, tile_size=180, xy_res=0.1))
    assert (gm.size, gm.tile_size, gm.xy_res) == ((360, 360), (180, 180), (0.1, 0.1))
    gm = GM(**dict(_BASE, tile_size=None))
    assert gm.tile_size == (720, 360) and gm.is_tiled is False


def test_non_regular_guards():
    gm = GM(**dict(_BASE, is_regular=False))
    with pytest.raises(ValueError, match="must be a regular grid mapping"):
        GM.assert_regular(gm)
    with pytest.raises(NotImplementedError, match="Operation not implemented for non-regular grid mappings"):
        gm._assert_regular()


def test_repr_markdown():
    md = GM(**_BASE)._repr_markdown_()
    for line in ("class: **GridMapping**", "* is_regular: True", "* is_j_axis_up: False", "* is_lon_360: False",
                 "* crs: EPSG:4326", "* xy_res: (0.5, 0.5)", "* xy_bbox: (-180.0, -90.0, 180.0, 90.0)",
                 "* ij_bbox: (0, 0, 720, 360)", "* xy_dim_names: ('x', 'y')", "* xy_var_names: ('x', 'y')",
                 "* size: (720, 360)", "* tile_size: (360, 180)"):
        assert line in md.split("\n")
    md = GM(**dict(_BASE, is_regular=None, is_j_axis_up=None, is_lon_360=None))._repr_markdown_()
    assert "* is_regular: _unknown_" in md and "* is_j_axis_up: _unknown_" in md and "* is_lon_360: _unknown_" in md
    assert "* xy_res: (0.5, 0.5)  _estimated_" in md
    assert str(xrs.CRS_CRS84) == "OGC:CRS84" and str(xrs.CRS.from_epsg(32633)) == "EPSG:32633"

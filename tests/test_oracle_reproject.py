"""Pin the reprojection oracle (CPU only): ``oracle.proj`` against the reference's known-answer
vectors and round trips, ``oracle.reproject.sample_window`` bit-for-bit against the reference's own
``_reproject_block`` (tests/golden/reproject.npz) and the whole chain against the expected arrays of
the reference's ``tests/test_reproject.py``."""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import proj as oproj
from oracle import reproject as orep

from .helpers import assert_same, load_golden

UTM32 = oproj.from_epsg(32632)
LAEA = oproj.from_epsg(3035)
WGS84 = oproj.from_epsg(4326)
WEBMERC = oproj.from_epsg(3857)


# ---------------------------------------------------------------------------
# projections
# ---------------------------------------------------------------------------
def test_ref_crs84_to_utm32_known_answer():
    """tests/gridmapping/test_transform.py:46-65 (7 decimals)."""
    g = ogrid.regular_grid((3, 3), (10, 53), 0.1)
    xx, yy = np.meshgrid(ogrid.x_centres(g), ogrid.y_centres(g))
    x, y = oproj.transform(WGS84, UTM32, xx, yy)
    np.testing.assert_almost_equal(x, np.array([
        [570057.076286, 576728.9360228, 583400.7295284],
        [570220.3304187, 576907.7404859, 583595.0849538],
        [570383.3684844, 577086.3083212, 583789.1831954]]), decimal=7)
    np.testing.assert_almost_equal(y, np.array([
        [5900595.928991, 5900698.5746648, 5900810.5532744],
        [5889471.9033896, 5889574.6540572, 5889686.7472201],
        [5878348.0594403, 5878450.9138481, 5878563.1201969]]), decimal=7)


def test_laea_epsg_guidance_note_example():
    """EPSG Guidance Note 7-2, Lambert Azimuthal Equal Area worked example (ETRS89-LAEA):
    lat 50 N, lon 5 E -> E 3962799.45, N 2999718.85."""
    x, y = oproj.laea_forward(LAEA, 5.0, 50.0)
    assert abs(float(x) - 3962799.45) < 0.01 and abs(float(y) - 2999718.85) < 0.01
    lon, lat = oproj.laea_inverse(LAEA, 3962799.45, 2999718.85)
    assert abs(float(lon) - 5.0) < 1e-7 and abs(float(lat) - 50.0) < 1e-7


@pytest.mark.parametrize("p", [UTM32, LAEA, WEBMERC, oproj.from_epsg(32733)])
def test_projection_round_trip(p):
    rng = np.random.default_rng(3)
    lon0 = p.lon0 if p.kind != oproj.WEBMERC else 0.0
    lon = lon0 + rng.uniform(-12, 12, 2000)
    lat = rng.uniform(-75, 75, 2000) if p.kind != oproj.LAEA else rng.uniform(25, 75, 2000)
    x, y = oproj.forward(p, lon, lat)
    lon2, lat2 = oproj.inverse(p, x, y)
    assert np.abs(lon2 - lon).max() < 1e-11 and np.abs(lat2 - lat).max() < 1e-11


def test_webmerc_known_values():
    x, y = oproj.forward(WEBMERC, np.array([180.0, 0.0]), np.array([0.0, 85.0511287798066]))
    np.testing.assert_allclose(x, [20037508.342789244, 0.0], atol=1e-6)
    np.testing.assert_allclose(y, [0.0, 20037508.342789244], atol=1e-5)


# EPSG:3857 has no vector in the reference's own tests; these are the published ones:
# IOGP Guidance Note 7-2, section 3.5.1 "Popular Visualisation Pseudo-Mercator" worked example
# (forward and reverse), and the square world extent +-pi * 6378137 m at latitude
# atan(sinh(pi)) = 85.0511287798066 deg that EPSG:3857 tilings are defined on.
GN7_2_LON, GN7_2_LAT = -(100 + 20 / 60), 24 + 22 / 60 + 54.433 / 3600
GN7_2_E, GN7_2_N = -11169055.58, 2800000.00
GN7_2_REV_N, GN7_2_REV_LAT = 2810000.00, 24 + 27 / 60 + 48.889 / 3600
WORLD = 20037508.342789244


def test_webmerc_guidance_note_7_2_example():
    x, y = oproj.transform(oproj.from_epsg(4326), WEBMERC, np.array([GN7_2_LON]), np.array([GN7_2_LAT]))
    assert abs(x[0] - GN7_2_E) < 0.005 and abs(y[0] - GN7_2_N) < 0.005  # published to the centimetre
    lon, lat = oproj.transform(WEBMERC, oproj.from_epsg(4326), np.array([GN7_2_E]), np.array([GN7_2_REV_N]))
    assert abs(lon[0] - GN7_2_LON) < 5e-8 and abs(lat[0] - GN7_2_REV_LAT) < 5e-8  # published to 0.001 arc second
    x, y = oproj.transform(oproj.from_epsg(4326), WEBMERC, np.array([180.0, -180.0]),
                           np.array([85.0511287798066, -85.0511287798066]))
    np.testing.assert_allclose(x, [WORLD, -WORLD], rtol=0, atol=1e-8)
    np.testing.assert_allclose(y, [WORLD, -WORLD], rtol=0, atol=2e-6)


def test_transform_bounds_webmerc_tile_at_the_antimeridian():
    """The easternmost tile of config C5 (36000^2 over the world extent, tile 4500): its box ends on
    x = +WORLD, i.e. on the antimeridian.  The densified boundary must come back as a contiguous
    longitude interval ending at 180 (not wrapped to -180), latitudes from the closed form."""
    t = 2 * WORLD / 8
    box = oproj.transform_bounds(WEBMERC, oproj.from_epsg(4326), WORLD - t, 0.0, WORLD, t)
    assert abs(box[0] - 135.0) < 1e-9 and abs(box[2] - 180.0) < 1e-9
    assert box[1] == 0.0 and abs(box[3] - np.degrees(np.arctan(np.sinh(t / 6378137.0)))) < 1e-12
    west = oproj.transform_bounds(WEBMERC, oproj.from_epsg(4326), -WORLD, -t, -WORLD + t, 0.0)
    assert abs(west[0] + 180.0) < 1e-9 and abs(west[2] + 135.0) < 1e-9


# IOGP Guidance Note 7-2, Transverse Mercator worked example (OSGB 1936 / British National Grid): Airy 1830
# ellipsoid, a NON-ZERO latitude of origin, scale factor and false origin -- every parameter of the
# family that the UTM known-answer vector of the reference (test_transform.py:46-65) leaves at its default.
GN7_2_TM = dict(a=6377563.396, inv_f=299.32496, lon0=-2.0, lat0=49.0, k0=0.9996012717, fe=400000.0, fn=-100000.0)
GN7_2_TM_LON, GN7_2_TM_LAT = 0.5, 50.5
GN7_2_TM_E, GN7_2_TM_N = 577274.99, 69740.50


def test_tmerc_guidance_note_7_2_example():
    p = oproj.Proj(oproj.TMERC, **GN7_2_TM)
    e, n = oproj.tmerc_forward(p, np.array([GN7_2_TM_LON]), np.array([GN7_2_TM_LAT]))
    # published to the centimetre from the USGS series; the Krueger series (PROJ's etmerc) lands within 1 cm
    assert abs(e[0] - GN7_2_TM_E) < 0.011 and abs(n[0] - GN7_2_TM_N) < 0.011
    lon, lat = oproj.tmerc_inverse(p, np.array([GN7_2_TM_E]), np.array([GN7_2_TM_N]))
    assert abs(lon[0] - GN7_2_TM_LON) < 2e-7 and abs(lat[0] - GN7_2_TM_LAT) < 2e-7  # 1 cm on the ground
    # and the round trip is exact to the nanodegree
    lon2, lat2 = oproj.tmerc_inverse(p, e, n)
    assert abs(lon2[0] - GN7_2_TM_LON) < 1e-11 and abs(lat2[0] - GN7_2_TM_LAT) < 1e-11


# Snyder, "Map Projections -- A Working Manual" (USGS Professional Paper 1395, 1987), numerical examples on
# the Clarke 1866 ellipsoid, published to 0.1 m: Transverse Mercator (phi0 = 0, lambda0 = 75 W, k0 = 0.9996;
# 40 30' N, 73 30' W) and Lambert Azimuthal Equal-Area, ellipsoidal oblique form (phi1 = 40 N, lambda0 = 100 W;
# 30 N, 110 W).
CLARKE_1866 = dict(a=6378206.4, inv_f=294.978698214)


def test_snyder_worked_examples_on_the_clarke_1866_ellipsoid():
    tm = oproj.Proj(oproj.TMERC, lon0=-75.0, lat0=0.0, k0=0.9996, **CLARKE_1866)
    x, y = oproj.tmerc_forward(tm, np.array([-73.5]), np.array([40.5]))
    assert abs(x[0] - 127106.5) < 0.06 and abs(y[0] - 4484124.4) < 0.06
    lon, lat = oproj.tmerc_inverse(tm, x, y)
    assert abs(lon[0] + 73.5) < 1e-11 and abs(lat[0] - 40.5) < 1e-11
    laea = oproj.Proj(oproj.LAEA, lon0=-100.0, lat0=40.0, **CLARKE_1866)
    x, y = oproj.laea_forward(laea, np.array([-110.0]), np.array([30.0]))
    assert abs(x[0] + 965932.1) < 0.06 and abs(y[0] + 1056814.9) < 0.06
    lon, lat = oproj.laea_inverse(laea, x, y)
    assert abs(lon[0] + 110.0) < 1e-10 and abs(lat[0] - 30.0) < 1e-10


# ---------------------------------------------------------------------------
# _reproject_block
# ---------------------------------------------------------------------------
def _cases():
    z = load_golden("reproject.npz")
    return z, [str(c) for c in z["cases"]]


@pytest.mark.parametrize("case", _cases()[1])
@pytest.mark.parametrize("method", ["nearest", "bilinear", "triangular"])
def test_sample_window_matches_reference_block(case, method):
    z, _ = _cases()
    xres, yres = z[f"{case}/res"]
    for dn in ("f32", "f64", "u8", "i16", "i32", "i64"):
        got = orep.sample_window(z[f"{case}/xx"], z[f"{case}/yy"], z[f"{case}/src_{dn}"], z[f"{case}/x_coord"][0],
                                 z[f"{case}/y_coord"][0], xres, yres, method)
        assert_same(got, z[f"{case}/out_{dn}_{method}"], f"{case}/{dn}/{method}")


def test_sample_window_invalid_method():
    with pytest.raises(NotImplementedError, match="interp_methods must be one of 0, 1, 'nearest', 'bilinear', 'triangular'"):
        orep.sample_window(np.zeros((2, 2)), np.zeros((2, 2)), np.zeros((1, 3, 3)), 0.0, 0.0, 1.0, 1.0, "cubic")


# ---------------------------------------------------------------------------
# expectations of the reference's tests/test_reproject.py
# ---------------------------------------------------------------------------
BAND_5X5 = np.arange(25).reshape(5, 5)  # tests/sampledata.py:95-109 (int64 -> nearest, fill -1)


def _reproject_5x5(size, xy_min, res, tgt_proj, j_up=False):
    g = ogrid.regular_grid(size, xy_min, res, is_j_axis_up=j_up)
    x = np.arange(565300.0, 565800.0, 100.0)
    y = np.arange(5934300.0, 5933800.0, -100.0)
    return orep.reproject_dataset_like(BAND_5X5, x, y, g, tgt_proj, UTM32)


def test_ref_reproject_target_gm():
    """tests/test_reproject.py:21-39."""
    np.testing.assert_array_equal(_reproject_5x5((5, 5), (4320080, 3382480), 80, LAEA), [
        [1, 1, 2, 3, 4], [6, 6, 7, 8, 9], [11, 12, 12, 13, 14], [16, 17, 17, 18, 19], [21, 17, 17, 18, 19]])


def test_ref_reproject_target_gm_j_axis_up():
    """tests/test_reproject.py:78-99."""
    np.testing.assert_array_equal(_reproject_5x5((5, 5), (4320080, 3382480), 80, LAEA, j_up=True), [
        [21, 17, 17, 18, 19], [16, 17, 17, 18, 19], [11, 12, 12, 13, 14], [6, 6, 7, 8, 9], [1, 1, 2, 3, 4]])


def test_ref_reproject_finer_and_coarser():
    """tests/test_reproject.py:122-160."""
    np.testing.assert_array_equal(_reproject_5x5((5, 5), (4320080, 3382480), 20, LAEA), [
        [15, 16, 16, 16, 16], [15, 16, 16, 16, 16], [15, 16, 16, 16, 16], [20, 21, 21, 21, 21],
        [20, 21, 21, 21, 21]])
    np.testing.assert_array_equal(_reproject_5x5((3, 3), (4320050, 3382500), 120, LAEA), [
        [0, 1, 2], [5, 6, 7], [15, 16, 17]])


def test_ref_reproject_geographic_targets():
    """tests/test_reproject.py:162-201."""
    np.testing.assert_array_equal(_reproject_5x5((5, 5), (9.9886, 53.5499), 0.0006, WGS84), [
        [7, 8, 8, 8, 9], [12, 13, 13, 13, 14], [12, 13, 13, 13, 14], [17, 18, 18, 18, 19], [22, 23, 23, 23, 24]])
    np.testing.assert_array_equal(_reproject_5x5((5, 5), (9.9886, 53.5499), 0.0003, WGS84), [
        [12, 12, 12, 13, 13], [17, 17, 17, 18, 18], [17, 17, 17, 18, 18], [22, 17, 17, 18, 18],
        [22, 22, 22, 23, 23]])


def test_ref_reproject_complex_dask_array():
    """tests/test_reproject.py:203-245: LAEA source (j-axis up -> flipped, reproject.py:115-118),
    geographic target tiled 5x5, triangular and bilinear; values to 4 decimals."""
    nt, nx, ny = 2, 100, 100
    x = np.linspace(3900000, 4500000, nx)
    y = np.linspace(2600000, 3200000, ny)
    data = np.arange(10 * nx * ny, dtype=np.float32).reshape(10, nx, ny)[:nt]
    g = ogrid.regular_grid((10, 10), (6.0, 48.0), 0.2, tile_size=(5, 5))
    for method, first in (("triangular", 6353.582), ("bilinear", 6353.5823)):
        out = orep.reproject_dataset_like(data, x, y, g, WGS84, LAEA, method=method)
        assert abs(float(out[0, 0, 0]) - first) < 5e-4, (method, out[0, 0, 0])
        assert abs(float(out[0, -1, -1]) - 3007.1228) < 5e-4, (method, out[0, -1, -1])

"""GPU parity of the rectify path: libxrs.so (through the C ABI / ctypes) against
the CPU oracle and the committed reference-kernel goldens.

ij image: bit-exact expected (tolerance of the north star: 1e-6 px);
nearest / bilinear / triangular gathers: bit-exact (same fp64 arithmetic, no FMA).
"""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import rectify as orect

from .helpers import assert_same, covering_grid_args, grid_from_golden, load_golden, swath

pytestmark = pytest.mark.gpu
nan = np.nan


@pytest.fixture(scope="module")
def xrs():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import xcube_resampling_b200 as pkg
    from xcube_resampling_b200 import _dev, rectify

    pkg.dev = _dev
    pkg.rect = rectify
    return pkg


def _gm(xrs, g: ogrid.RegularGrid, crs="EPSG:4326"):
    return xrs.GridMapping.regular((g.width, g.height), (g.x_min, g.y_min), (g.x_res, g.y_res), crs,
                                   tile_size=(g.tile_w, g.tile_h), is_j_axis_up=g.is_j_axis_up)


def _device_rectify(xrs, x, y, gm):
    xd = xrs.dev.to_device(x, dtype=np.float64)
    yd = xrs.dev.to_device(y, dtype=np.float64)
    windows = xrs.rect.tile_source_windows_dev(xd, yd, gm.xy_bboxes, xrs.rect._xy_border(gm), 1)
    ij = xrs.rect.compute_target_source_ij(xd, yd, gm, tile_boxes=windows)
    return xrs.dev.to_host(windows), ij


def _golden_cases():
    z = load_golden("rectify.npz")
    return [str(c) for c in z["cases"]]


@pytest.mark.parametrize("case", _golden_cases())
def test_golden_windows_ij_and_gather(xrs, case):
    z = load_golden("rectify.npz")
    g = grid_from_golden(z[f"{case}/grid"])
    gm = _gm(xrs, g)
    # the golden grid was built from raw floats; make sure the GridMapping kept them
    assert (gm.x_min, gm.y_min, gm.y_max, gm.x_res, gm.y_res) == (g.x_min, g.y_min, g.y_max, g.x_res, g.y_res)
    windows, ij = _device_rectify(xrs, z[f"{case}/x"], z[f"{case}/y"], gm)
    assert_same(windows, z[f"{case}/windows"], "K0 windows")
    ij_host = xrs.dev.to_host(ij)
    ref_ij = z[f"{case}/ij"]
    assert np.array_equal(np.isnan(ij_host), np.isnan(ref_ij)), "NaN mask of the ij image differs"
    assert np.nanmax(np.abs(ij_host - ref_ij), initial=0.0) <= 1e-6  # north-star tolerance
    assert_same(ij_host, ref_ij, "K1 ij (bit-exact)")
    for vname, fill in (("f32", nan), ("u8", 255), ("i16", -1), ("f64", nan)):
        src = xrs.dev.to_device(z[f"{case}/src_{vname}"])
        for method in ("nearest", "bilinear", "triangular"):
            out = xrs.dev.to_host(xrs.rect.gather_ij(src, ij, method, fill))
            assert_same(out, z[f"{case}/out_{vname}_{method}"], f"K2 {vname}/{method}")


@pytest.mark.parametrize("shape,theta,tile", [
    ((700, 520), 12.0, 512),       # several reference tiles, windows span many CTA tiles
    ((1030, 333), -40.0, None),    # single reference tile, odd pitch (bulk-copy alignment shifts)
    ((257, 900), 77.0, (100, 64)),  # strongly rotated, small ragged tiles
    ((400, 300), 0.0, 2048),       # axis-aligned swath
])
def test_seeded_swaths_against_oracle(xrs, shape, theta, tile):
    w, h = shape
    x, y = swath(w, h, theta=theta, seed=w + h)
    res = 0.0027
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile)  # tile may exceed the image (as in the reference)
    gm = _gm(xrs, g)
    windows, ij = _device_rectify(xrs, x, y, gm)
    assert_same(windows, orect.source_windows(x, y, g), "K0 windows")
    ref_ij = orect.rectify_ij(x, y, g)
    ij_host = xrs.dev.to_host(ij)
    assert np.array_equal(np.isnan(ij_host), np.isnan(ref_ij))
    assert_same(ij_host, ref_ij, "K1 ij")
    rng = np.random.default_rng(1)
    src = rng.random((5, h, w)).astype(np.float32)
    src[1, rng.random((h, w)) < 0.01] = nan
    sd = xrs.dev.to_device(src)
    for method in ("nearest", "bilinear"):
        out = xrs.dev.to_host(xrs.rect.gather_ij(sd, ij, method, nan))
        assert_same(out, orect.gather(src, ij_host, method, nan), f"K2 {method}")


def test_fine_target_large_quads(xrs):
    """Target 6x finer than the source: every quad covers ~36 pixels."""
    x, y = swath(60, 50, theta=20.0, seed=3)
    res = 0.0027 / 6.3
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=200)
    gm = _gm(xrs, g)
    _, ij = _device_rectify(xrs, x, y, gm)
    assert_same(xrs.dev.to_host(ij), orect.rectify_ij(x, y, g), "ij")


def test_coarse_target_many_quads_per_pixel(xrs):
    """Target 5x coarser: many quads compete for each pixel -> first-writer rule decides."""
    x, y = swath(500, 400, theta=-15.0, seed=5)
    res = 0.0027 * 5.0
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=64)
    g = ogrid.RegularGrid(g.width, g.height, min(g.tile_w, g.width), min(g.tile_h, g.height), g.x_min, g.y_min,
                          g.x_max, g.y_max, g.x_res, g.y_res, g.is_j_axis_up)
    gm = _gm(xrs, g)
    _, ij = _device_rectify(xrs, x, y, gm)
    assert_same(xrs.dev.to_host(ij), orect.rectify_ij(x, y, g), "ij")


def test_wide_window_forces_chunking(xrs):
    """A source much wider than one CTA window chunk (K1_MAX_COLS) mapped onto few pixels."""
    w, h = 3000, 40
    x, y = swath(w, h, theta=0.0, seed=9)
    res = 0.0027 * 40.0
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res)
    gm = _gm(xrs, g)
    _, ij = _device_rectify(xrs, x, y, gm)
    assert_same(xrs.dev.to_host(ij), orect.rectify_ij(x, y, g), "ij")


# ---------------------------------------------------------------------------
# entry point, mirroring the reference's tests/test_rectify.py
# ---------------------------------------------------------------------------
LON_2X2 = np.array([[1.0, 6.0], [0.0, 2.0]])
LAT_2X2 = np.array([[56.0, 53.0], [52.0, 50.0]])
RAD_2X2 = np.array([[1.0, 2.0], [3.0, 4.0]])


def _ds_2x2(xrs, rad=RAD_2X2, lon=LON_2X2):
    return xrs.Dataset(data_vars=dict(rad=(("y", "x"), rad)),
                       coords=dict(lon=(("y", "x"), lon), lat=(("y", "x"), LAT_2X2)))


def test_rectify_dataset_2x2_to_default(xrs):
    # tests/test_rectify.py:42-61
    gm = xrs.GridMapping.regular(size=(4, 4), xy_min=(-1, 49), xy_res=2, crs=xrs.CRS_WGS84)
    out = xrs.rectify_dataset(_ds_2x2(xrs), target_gm=gm, interp_methods=0)
    np.testing.assert_almost_equal(out["rad"].values, np.array([
        [nan, nan, nan, nan], [nan, 1.0, 2.0, nan], [3.0, 3.0, 2.0, nan], [nan, 4.0, nan, nan]]))
    assert out["rad"].dims == ("lat", "lon")
    assert "spatial_ref" in out.coords


def test_rectify_dataset_2x2_to_regular(xrs):
    # tests/test_rectify.py:63-78: target derived with to_regular()
    out = xrs.rectify_dataset(_ds_2x2(xrs), interp_methods=0)
    np.testing.assert_almost_equal(out["rad"].values, np.array([
        [nan, nan, nan, nan], [nan, nan, nan, nan], [nan, 2.0, nan, nan], [nan, nan, nan, nan]]))


def test_rectify_dataset_3d_and_passthrough(xrs):
    # tests/test_rectify.py:80-110
    rad = np.stack([RAD_2X2, RAD_2X2])
    ds = xrs.Dataset(data_vars=dict(rad=(("time", "y", "x"), rad), time_series=(("time",), np.array([1, 2]))),
                     coords=dict(lon=(("y", "x"), LON_2X2), lat=(("y", "x"), LAT_2X2), time=np.array([0, 1])))
    gm = xrs.GridMapping.regular(size=(4, 4), xy_min=(-1, 49), xy_res=2, crs=xrs.CRS_WGS84)
    out = xrs.rectify_dataset(ds, target_gm=gm, interp_methods=0)
    exp = np.array([[nan, nan, nan, nan], [nan, 1.0, 2.0, nan], [3.0, 3.0, 2.0, nan], [nan, 4.0, nan, nan]])
    np.testing.assert_almost_equal(out["rad"].values, np.stack([exp, exp]))
    assert out["rad"].dims == ("time", "lat", "lon")
    np.testing.assert_array_equal(out["time_series"].values, [1, 2])


@pytest.mark.parametrize("tile_size", [None, 7, 5, (3, 13), (13, 3)])
@pytest.mark.parametrize("j_up", [False, True])
def test_rectify_dataset_2x2_to_13x13(xrs, tile_size, j_up):
    # tests/test_rectify.py:261-387
    from .test_oracle_golden import EXPECTED_RAD_13X13

    gm = xrs.GridMapping.regular(size=(13, 13), xy_min=(-0.25, 49.75), xy_res=0.5, crs=xrs.CRS_WGS84,
                                 tile_size=tile_size, is_j_axis_up=j_up)
    out = xrs.rectify_dataset(_ds_2x2(xrs), target_gm=gm, interp_methods=0)
    np.testing.assert_almost_equal(out["lon"].values, np.arange(0, 6.1, 0.5))
    exp_lat = np.arange(50, 56.1, 0.5) if j_up else np.arange(56, 49.9, -0.5)
    np.testing.assert_almost_equal(out["lat"].values, exp_lat)
    np.testing.assert_almost_equal(out["rad"].values, EXPECTED_RAD_13X13[::-1] if j_up else EXPECTED_RAD_13X13)


def test_rectify_dataset_antimeridian(xrs):
    # tests/test_rectify.py:389-424
    from .test_oracle_golden import EXPECTED_RAD_13X13

    lon = np.array([[+179.0, -176.0], [+178.0, +180.0]])
    gm = xrs.GridMapping.regular(size=(13, 13), xy_min=(177.75, 49.75), xy_res=0.5, crs=xrs.CRS_WGS84)
    assert gm.is_lon_360 is True
    out = xrs.rectify_dataset(_ds_2x2(xrs, lon=lon), target_gm=gm, interp_methods=0)
    np.testing.assert_almost_equal(out["lon"].values, np.array(
        [178.0, 178.5, 179.0, 179.5, 180.0, -179.5, -179.0, -178.5, -178.0, -177.5, -177.0, -176.5, -176.0]))
    np.testing.assert_almost_equal(out["rad"].values, EXPECTED_RAD_13X13)


@pytest.mark.parametrize("method,decimal,expected", [
    ("triangular", 3, [
        [nan, 1.000, nan, nan, nan, nan, nan],
        [nan, 1.478, 1.391, nan, nan, nan, nan],
        [nan, 1.957, 1.870, 1.784, 1.697, nan, nan],
        [nan, 2.435, 2.348, 2.261, 2.174, 2.087, 2.000],
        [3.000, 3.000, 3.000, 3.000, 3.000, nan, nan],
        [nan, 4.000, 4.000, 4.000, nan, nan, nan],
        [nan, nan, 5.000, nan, nan, nan, nan]]),
    ("bilinear", 3, [
        [nan, 1.000, nan, nan, nan, nan, nan],
        [nan, 1.488, 1.410, nan, nan, nan, nan],
        [nan, 1.994, 1.949, 1.858, 1.722, nan, nan],
        [nan, 2.520, 2.506, 2.448, 2.344, 2.195, 2.000],
        [3.000, 3.112, 3.163, 3.153, 3.082, nan, nan],
        [nan, 4.000, 4.041, 4.020, nan, nan, nan],
        [nan, nan, 5.000, nan, nan, nan, nan]]),
])
def test_rectify_dataset_7x7_interpolation(xrs, method, decimal, expected):
    # tests/test_rectify.py:146-218
    rad = RAD_2X2 + np.array([[0.0, 0.0], [0.0, 1.0]])
    gm = xrs.GridMapping.regular(size=(7, 7), xy_min=(-0.5, 49.5), xy_res=1.0, crs=xrs.CRS_WGS84)
    out = xrs.rectify_dataset(_ds_2x2(xrs, rad=rad), target_gm=gm, interp_methods=method)
    np.testing.assert_almost_equal(out["rad"].values, np.array(expected), decimal=decimal)


def test_rectify_dataset_invalid_method(xrs):
    # tests/test_rectify.py:220-228
    gm = xrs.GridMapping.regular(size=(7, 7), xy_min=(-0.5, 49.5), xy_res=1.0, crs=xrs.CRS_WGS84)
    with pytest.raises(NotImplementedError):
        xrs.rectify_dataset(_ds_2x2(xrs), target_gm=gm, interp_methods="cubic")


@pytest.mark.parametrize("xy_min", [(10.0, 50.0), (-10.0, 50.0), (0.0, 58.0), (0.0, 42.0)])
def test_rectify_dataset_no_overlap(xrs, xy_min):
    # tests/test_rectify.py:426-459
    gm = xrs.GridMapping.regular(size=(13, 13), xy_min=xy_min, xy_res=0.5, crs=xrs.CRS_WGS84)
    out = xrs.rectify_dataset(_ds_2x2(xrs), target_gm=gm, interp_methods=0)
    assert np.isnan(out["rad"].values).all()


def test_c_abi_rejects_bad_arguments(xrs):
    from xcube_resampling_b200 import _lib

    lib = _lib.load()
    assert lib.xrs_gather_ij(None, None, 1, 0, 4, 4, 4, 0, 0, 4, 4, None, 4, 4, 0, 0.0, None) != 0
    assert b"null" in lib.xrs_last_error()


def test_band_pipeline_equals_plain_path(xrs, monkeypatch):
    """The band-chunk pipeline (upload / gather / download overlapped, shared ij image) gives the same
    bytes as the fused single-variable path (ij resolved in registers) -- and both match the oracle."""
    from xcube_resampling_b200 import rectify as xrect

    w, h = 300, 220
    x, y = swath(w, h, theta=20.0, seed=4)
    res = 0.0027
    size, xy_min = covering_grid_args(x, y, res)
    rng = np.random.default_rng(4)
    bands = rng.random((11, h, w)).astype(np.float32)  # 11 = 2 full chunks of 4 + a ragged one
    ds = xrs.Dataset(data_vars=dict(bands=(("band", "y", "x"), bands)),
                     coords=dict(lon=(("y", "x"), x), lat=(("y", "x"), y)))
    source_gm = xrs.GridMapping.from_coords(x, y, "EPSG:4326", xy_res=res, xy_dim_names=("x", "y"))
    target_gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=128)
    for method in ("nearest", "bilinear"):
        plain = xrs.rectify_dataset(ds, target_gm=target_gm, source_gm=source_gm, interp_methods=method)["bands"].values
        monkeypatch.setattr(xrect, "_PIPELINE_MIN_BYTES", 0)
        piped = xrs.rectify_dataset(ds, target_gm=target_gm, source_gm=source_gm, interp_methods=method)["bands"].values
        monkeypatch.undo()
        assert_same(piped, plain, method)
        g = ogrid.regular_grid(size, xy_min, res, tile_size=128)
        assert_same(piped, orect.gather(bands, orect.rectify_ij(x, y, g), method, nan), f"{method} vs oracle")


# ---------------------------------------------------------------------------
# adversarial geometry for the pixel-space scatter (edge functions + margins)
# ---------------------------------------------------------------------------
def _ij_equal(xrs, x, y, g, crs="EPSG:4326"):
    gm = _gm(xrs, g, crs)
    windows, ij = _device_rectify(xrs, x, y, gm)
    assert_same(windows, orect.source_windows(x, y, g), "K0 windows")
    assert_same(xrs.dev.to_host(ij), orect.rectify_ij(x, y, g), "K1 ij")


@pytest.mark.parametrize("tile", [None, 16, (24, 10)])
@pytest.mark.parametrize("ratio", [1.0, 2.0, 0.5, 3.0])
def test_axis_aligned_source_on_pixel_centres(xrs, tile, ratio):
    """Regular source whose vertices sit exactly on target pixel centres / corners: every triangle
    edge passes through pixel centres (u, v exactly 0 or 1) -- the margin / exact-arithmetic path."""
    res = 0.25
    w, h = 41, 33
    x = np.broadcast_to(10.0 + ratio * res * (np.arange(w) + 0.5), (h, w)).copy()
    y = np.broadcast_to((50.0 - ratio * res * (np.arange(h) + 0.5))[:, None], (h, w)).copy()
    size = (int(w * ratio) + 4, int(h * ratio) + 4)
    g = ogrid.regular_grid(size, (10.0 - 2 * res, 50.0 - (size[1] - 2) * res), res, tile_size=tile)
    _ij_equal(xrs, x, y, g)


def test_holes_duplicates_and_folds(xrs):
    """NaN holes, duplicated rows / columns (zero-area quads) and a fold (negative determinants)."""
    x, y = swath(180, 150, theta=25.0, seed=11)
    x[20:24, 30:50] = nan
    y[20:24, 30:50] = nan
    x[70, :] = x[69, :]
    y[70, :] = y[69, :]          # duplicated row
    x[:, 100] = x[:, 99]
    y[:, 100] = y[:, 99]         # duplicated column
    x[110:130] = x[110:130][::-1].copy()
    y[110:130] = y[110:130][::-1].copy()  # folded strip
    x[5, 5] = np.inf
    res = 0.0027
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=64)
    _ij_equal(xrs, x, y, g)


def test_projected_coordinate_magnitudes(xrs):
    """UTM-like metres: coordinates ~5e5 / 6e6 with 300 m pixels (rounding of the pixel-space
    transform is largest relative to the quad size here)."""
    lon, lat = swath(300, 260, theta=-8.0, seed=13)
    x = 500000.0 + (lon - 10.0) * 78000.0
    y = 5000000.0 + (lat - 45.0) * 111000.0
    res = 300.0
    size, xy_min = covering_grid_args(x, y, res)
    for tile in (None, 100):
        g = ogrid.regular_grid(size, xy_min, res, tile_size=tile)
        _ij_equal(xrs, x, y, g, "EPSG:32632")


def test_huge_quads_take_the_generic_path(xrs):
    """Quads spanning > 64 target pixels (K1_MAX_EXTENT) and a loose uv tolerance."""
    x, y = swath(12, 10, theta=33.0, seed=17)
    res = 0.0027 / 90.0
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=256)
    _ij_equal(xrs, x, y, g)


@pytest.mark.parametrize("dtype", [np.float32, np.uint8, np.int16, np.float64])
@pytest.mark.parametrize("tile", [None, 96])
def test_fused_rectify_gather_equals_two_step_path(xrs, dtype, tile):
    """xrs_rectify_gather (claims -> resolve in registers -> gather) == xrs_rectify_ij + xrs_gather_ij."""
    w, h = 333, 260
    x, y = swath(w, h, theta=-22.0, seed=21)
    x[40:44, 100:120] = nan
    res = 0.0027
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile)
    gm = _gm(xrs, g)
    rng = np.random.default_rng(2)
    src = (rng.random((5, h, w)) * 200).astype(dtype)
    fill = nan if np.issubdtype(dtype, np.floating) else 255 if dtype == np.uint8 else -1
    xd, yd = xrs.dev.to_device(x), xrs.dev.to_device(y)
    for pitched in (True, False):  # TMA-staged and direct kernels
        sd = xrs.dev.to_device_pitched(src) if pitched else xrs.dev.to_device(src)
        for rows in (None, (64, 160)):
            plan = xrs.rect.RectifyPlan(gm, xd.device, rows=rows)
            ij = plan.ij(xd, yd)
            for method in ("nearest", "bilinear", "triangular"):
                want = xrs.dev.to_host(xrs.rect.gather_ij(sd, ij, method, fill))
                got = xrs.dev.to_host(plan.rectify_gather(xd, yd, sd, method, fill))
                assert_same(got, want, f"{np.dtype(dtype).name} {method} rows={rows} pitched={pitched}")

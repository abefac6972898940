"""GPU parity of the reprojection path (xrs_transform_points, xrs_reproject through ctypes).

* projection math: device vs ``oracle.proj`` (both restate PROJ's published formulas with different
  evaluation schemes -- Clenshaw series on the device, Newton iterations in the oracle): agreement
  to 1e-7 m / 1e-12 deg, and the reference's CRS84 -> UTM 32N known-answer to 7 decimals;
* index / tap arithmetic: with two geographic CRSs the transform is the identity in both
  implementations, so every method and dtype must be BIT-EXACT against the oracle (which is pinned
  bit-for-bit to the reference's ``_reproject_block``);
* real projections: nearest may differ only where the fractional index sits on a rounding tie
  (mismatch fraction asserted < 1e-4 and reported), bilinear / triangular within 1e-6 relative
  (north-star tolerance; values are float64 here so the bound is far from tight);
* the reference's own ``tests/test_reproject.py`` expectations through ``reproject_dataset``.
"""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import proj as oproj
from oracle import reproject as orep

from .helpers import assert_same

pytestmark = pytest.mark.gpu
nan = np.nan

EPSG = {"utm32": 32632, "utm33s": 32733, "laea": 3035, "webmerc": 3857, "wgs84": 4326, "etrs_utm32": 25832}


@pytest.fixture(scope="module")
def xrs():
    import torch

    assert torch.cuda.is_available()
    import xcube_resampling_b200 as pkg
    from xcube_resampling_b200 import _dev, reproject

    pkg.dev = _dev
    pkg.rep = reproject
    return pkg


# ---------------------------------------------------------------------------
# projections
# ---------------------------------------------------------------------------
def _lonlat(kind, n=20000, seed=0):
    rng = np.random.default_rng(seed)
    p = oproj.from_epsg(EPSG[kind])
    lon0 = 0.0 if p.kind in (oproj.WEBMERC, oproj.GEOGRAPHIC) else p.lon0
    span = 170.0 if p.kind == oproj.WEBMERC else 20.0
    lon = lon0 + rng.uniform(-span, span, n)
    lat = rng.uniform(25, 75, n) if p.kind == oproj.LAEA else rng.uniform(-84, 84, n)
    return p, lon, lat


@pytest.mark.parametrize("kind", ["utm32", "utm33s", "laea", "webmerc", "etrs_utm32"])
def test_forward_inverse_match_oracle(xrs, kind):
    p, lon, lat = _lonlat(kind)
    crs = f"EPSG:{EPSG[kind]}"
    x, y = xrs.rep.transform_points(lon, lat, "EPSG:4326", crs)
    ex, ey = oproj.forward(p, lon, lat)
    assert np.abs(x - ex).max() < 1e-7 and np.abs(y - ey).max() < 1e-7, (np.abs(x - ex).max(), np.abs(y - ey).max())
    lon2, lat2 = xrs.rep.transform_points(ex, ey, crs, "EPSG:4326")
    elon, elat = oproj.inverse(p, ex, ey)
    assert np.abs(lon2 - elon).max() < 1e-12 and np.abs(lat2 - elat).max() < 1e-12
    assert np.abs(lon2 - lon).max() < 1e-11 and np.abs(lat2 - lat).max() < 1e-11  # round trip


@pytest.mark.parametrize("pair", [("utm32", "laea"), ("laea", "utm32"), ("webmerc", "utm32"), ("utm32", "webmerc"),
                                  ("utm32", "utm33s")])
def test_projected_to_projected_matches_oracle(xrs, pair):
    a, b = pair
    pa, lon, lat = _lonlat(a, seed=5)
    if "laea" in pair:
        lat = np.clip(lat, 30, 72)
    lon = np.clip(lon, 0, 18)
    x, y = oproj.forward(pa, lon, lat)
    pb = oproj.from_epsg(EPSG[b])
    ex, ey = oproj.transform(pa, pb, x, y)
    gx, gy = xrs.rep.transform_points(x, y, f"EPSG:{EPSG[a]}", f"EPSG:{EPSG[b]}")
    assert np.abs(gx - ex).max() < 1e-6 and np.abs(gy - ey).max() < 1e-6


def test_ref_crs84_to_utm32_known_answer(xrs):
    """tests/gridmapping/test_transform.py:46-65."""
    gm = xrs.GridMapping.regular(size=(3, 3), xy_min=(10, 53), xy_res=0.1, crs="OGC:CRS84")
    xy = gm.xy_coords.values
    x, y = xrs.rep.transform_points(xy[0], xy[1], "OGC:CRS84", "EPSG:32632")
    np.testing.assert_almost_equal(x, np.array([
        [570057.076286, 576728.9360228, 583400.7295284],
        [570220.3304187, 576907.7404859, 583595.0849538],
        [570383.3684844, 577086.3083212, 583789.1831954]]), decimal=7)
    np.testing.assert_almost_equal(y, np.array([
        [5900595.928991, 5900698.5746648, 5900810.5532744],
        [5889471.9033896, 5889574.6540572, 5889686.7472201],
        [5878348.0594403, 5878450.9138481, 5878563.1201969]]), decimal=7)


def test_untransformable_points_are_nan(xrs):
    x, y = xrs.rep.transform_points([0.0, 10.0, nan], [95.0, 50.0, 1.0], "EPSG:4326", "EPSG:32632")
    assert np.isnan(x[0]) and np.isnan(y[0]) and np.isfinite(x[1]) and np.isnan(x[2])
    # beyond the transverse Mercator domain
    x, y = xrs.rep.transform_points([5e8], [0.0], "EPSG:32632", "EPSG:4326")
    assert np.isnan(x[0]) and np.isnan(y[0])


def test_transform_bounds_matches_oracle(xrs):
    boxes = np.array([[399960.0, 5890200.0, 509760.0, 6000000.0], [500000.0, 0.0, 600000.0, 100000.0]])
    got = xrs.rep.transform_bounds("EPSG:32632", "EPSG:4326", boxes)
    for k in range(2):
        want = oproj.transform_bounds(oproj.from_epsg(32632), oproj.from_epsg(4326), *boxes[k])
        np.testing.assert_allclose(got[k], want, rtol=0, atol=1e-11)


def test_webmerc_published_known_answers(xrs):
    """EPSG:3857 has no vector in the reference's tests: IOGP Guidance Note 7-2 section 3.5.1 worked example
    (forward and reverse) and the world-extent corners, through the device transform."""
    lon, lat = -(100 + 20 / 60), 24 + 22 / 60 + 54.433 / 3600
    x, y = xrs.rep.transform_points([lon], [lat], "EPSG:4326", "EPSG:3857")
    assert abs(x[0] + 11169055.58) < 0.005 and abs(y[0] - 2800000.00) < 0.005
    lon2, lat2 = xrs.rep.transform_points([-11169055.58], [2810000.00], "EPSG:3857", "EPSG:4326")
    assert abs(lon2[0] - lon) < 5e-8 and abs(lat2[0] - (24 + 27 / 60 + 48.889 / 3600)) < 5e-8
    world = 20037508.342789244
    x, y = xrs.rep.transform_points([180.0, -180.0], [85.0511287798066, -85.0511287798066], "EPSG:4326", "EPSG:3857")
    np.testing.assert_allclose(x, [world, -world], rtol=0, atol=1e-8)
    np.testing.assert_allclose(y, [world, -world], rtol=0, atol=2e-6)


def test_transform_bounds_webmerc_tiles_at_the_antimeridian(xrs):
    """Config C5's easternmost / westernmost tiles touch x = +-world extent: contiguous longitude
    intervals ending at +-180, equal to the oracle's."""
    world = 20037508.342789244
    t = 2 * world / 8
    boxes = np.array([[world - t, 0.0, world, t], [-world, -t, -world + t, 0.0], [-t, -t, t, t]])
    got = xrs.rep.transform_bounds("EPSG:3857", "EPSG:4326", boxes)
    for k in range(3):
        want = oproj.transform_bounds(oproj.from_epsg(3857), oproj.from_epsg(4326), *boxes[k])
        np.testing.assert_allclose(got[k], want, rtol=0, atol=1e-10)
    assert abs(got[0][2] - 180.0) < 1e-9 and abs(got[0][0] - 135.0) < 1e-9
    assert abs(got[1][0] + 180.0) < 1e-9 and abs(got[1][2] + 135.0) < 1e-9


# ---------------------------------------------------------------------------
# kernel arithmetic, identity transform: bit-exact
# ---------------------------------------------------------------------------
def _source(dtype, bands, h, w, seed=0):
    rng = np.random.default_rng(seed)
    if np.issubdtype(dtype, np.floating):
        a = rng.normal(size=(bands, h, w)).astype(dtype)
        a[0, h // 3, w // 4] = nan
        return a
    info = np.iinfo(dtype)
    return rng.integers(max(info.min, -2**40), min(info.max, 2**40), size=(bands, h, w), dtype=np.int64).astype(dtype)


def _gm_pair(xrs, tile, j_up=False, shift=(0.013, -0.021), size=(70, 53), res=0.0113):
    """Geographic source (0.01 deg) and a rotated-free but shifted / rescaled geographic target that
    pokes out of the source on two sides (exercises the constant padding)."""
    src = xrs.GridMapping.regular((90, 64), (10.0, 50.0), 0.01, "EPSG:4326")
    tgt = xrs.GridMapping.regular(size, (10.0 + shift[0] - 0.05, 50.0 + shift[1] - 0.03), res, "OGC:CRS84",
                                  tile_size=tile, is_j_axis_up=j_up)
    return src, tgt


def _oracle_full(src_gm, tgt_gm, data, method, fill):
    g = ogrid.regular_grid(tgt_gm.size, (tgt_gm.x_min, tgt_gm.y_min), tgt_gm.xy_res, tile_size=tgt_gm.tile_size,
                           is_j_axis_up=tgt_gm.is_j_axis_up)
    xs, ys = src_gm.x_values, src_gm.y_values
    geo = oproj.from_epsg(4326)
    return orep.reproject(data, float(xs[0]), float(ys[0]), src_gm.x_res, src_gm.y_res, float(ys[1] - ys[0]), g, geo,
                          geo, method, fill)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16, np.uint16, np.int32, np.int64])
@pytest.mark.parametrize("method", ["nearest", "bilinear", "triangular"])
@pytest.mark.parametrize("tile", [None, (32, 16)])
def test_identity_transform_is_bit_exact(xrs, dtype, method, tile, monkeypatch):
    src_gm, tgt_gm = _gm_pair(xrs, tile)
    data = _source(dtype, 3, src_gm.height, src_gm.width)
    fill = nan if np.issubdtype(dtype, np.floating) else (255 if dtype == np.uint8 else -1 if dtype != np.uint16 else 65535)
    want = _oracle_full(src_gm, tgt_gm, data, method, fill)
    plan = xrs.rep.ReprojectPlan(src_gm, tgt_gm)
    got = xrs.dev.to_host(plan.run(xrs.dev.to_device(data), method, fill))
    assert_same(got, want, f"{np.dtype(dtype).name}/{method}/{tile}")
    # a 128-byte row pitch lets the TMA-staged kernel run (k3_reproject_staged); border tiles of this
    # case still take its direct per-pixel path
    monkeypatch.setenv("XRS_K3_STAGED", "1")
    got = xrs.dev.to_host(plan.run(xrs.dev.to_device_pitched(data), method, fill))
    assert_same(got, want, f"staged {np.dtype(dtype).name}/{method}/{tile}")


def test_identity_transform_j_axis_up_target_and_2d(xrs):
    src_gm, tgt_gm = _gm_pair(xrs, (16, 16), j_up=True)
    data = _source(np.float32, 1, src_gm.height, src_gm.width)[0]
    for method in ("nearest", "bilinear"):
        want = _oracle_full(src_gm, tgt_gm, data, method, nan)
        got = xrs.dev.to_host(xrs.rep.ReprojectPlan(src_gm, tgt_gm).run(xrs.dev.to_device(data), method, nan))
        assert_same(got, want, method)


def test_bilinear_source_dtype_output(xrs):
    src_gm, tgt_gm = _gm_pair(xrs, None)
    data = _source(np.float32, 2, src_gm.height, src_gm.width)
    want = _oracle_full(src_gm, tgt_gm, data, "bilinear", nan).astype(np.float32)
    got = xrs.dev.to_host(xrs.rep.ReprojectPlan(src_gm, tgt_gm).run(xrs.dev.to_device(data), "bilinear", nan,
                                                                    out_dtype=np.float32))
    assert_same(got, want, "bilinear->f32")


def test_row_bands_and_resident_window_equal_full(xrs):
    src_gm, tgt_gm = _gm_pair(xrs, (32, 16))
    data = _source(np.float32, 4, src_gm.height, src_gm.width)
    full = xrs.dev.to_host(xrs.rep.ReprojectPlan(src_gm, tgt_gm).run(xrs.dev.to_device(data), "bilinear", nan))
    parts = []
    for rows in ((0, 16), (16, 32), (32, 53)):
        plan = xrs.rep.ReprojectPlan(src_gm, tgt_gm, rows=rows)
        i0, j0, i1, j1 = plan.footprint()
        window = xrs.dev.to_device(np.ascontiguousarray(data[:, j0:j1, i0:i1]))
        parts.append(xrs.dev.to_host(plan.run(window, "bilinear", nan, window_origin=(i0, j0))))
    assert_same(np.concatenate(parts, axis=1), full, "bands")


def test_many_bands_chunking(xrs):
    src_gm, tgt_gm = _gm_pair(xrs, None)
    data = _source(np.float32, 53, src_gm.height, src_gm.width)
    want = _oracle_full(src_gm, tgt_gm, data, "nearest", nan)
    got = xrs.dev.to_host(xrs.rep.ReprojectPlan(src_gm, tgt_gm).run(xrs.dev.to_device(data), "nearest", nan))
    assert_same(got, want, "53 bands")


# ---------------------------------------------------------------------------
# real projections against the oracle
# ---------------------------------------------------------------------------
def _case_utm_from_geographic(xrs, n=300, tile=128, lat_origin=990240.0):
    """Scaled-down C3: 0.0001 deg geographic source -> 10 m UTM 32N target near 9 deg N."""
    tgt = xrs.GridMapping.regular((n, n), (399960.0, lat_origin), 10.0, "EPSG:32632", tile_size=tile)
    box = oproj.transform_bounds(oproj.from_epsg(32632), oproj.from_epsg(4326), *tgt.xy_bbox)
    res = 0.0001
    x_min = math_floor_to(box[0], res) - 4 * res
    y_min = math_floor_to(box[1], res) - 4 * res
    w = int(np.ceil((box[2] - x_min) / res)) + 4
    h = int(np.ceil((box[3] - y_min) / res)) + 4
    src = xrs.GridMapping.regular((w, h), (x_min, y_min), res, "EPSG:4326")
    return src, tgt


def math_floor_to(v, step):
    return float(np.floor(v / step) * step)


def _oracle_projected(src_gm, tgt_gm, data, method, fill, src_epsg, tgt_epsg):
    g = ogrid.regular_grid(tgt_gm.size, (tgt_gm.x_min, tgt_gm.y_min), tgt_gm.xy_res, tile_size=tgt_gm.tile_size,
                           is_j_axis_up=tgt_gm.is_j_axis_up)
    xs, ys = src_gm.x_values, src_gm.y_values
    return orep.reproject(data, float(xs[0]), float(ys[0]), src_gm.x_res, src_gm.y_res, float(ys[1] - ys[0]), g,
                          oproj.from_epsg(tgt_epsg), oproj.from_epsg(src_epsg), method, fill)


@pytest.mark.parametrize("method", ["nearest", "bilinear", "triangular"])
def test_geographic_to_utm_matches_oracle(xrs, method):
    src_gm, tgt_gm = _case_utm_from_geographic(xrs)
    rng = np.random.default_rng(1)
    data = rng.random((3, src_gm.height, src_gm.width)).astype(np.float32)
    want = _oracle_projected(src_gm, tgt_gm, data, method, nan, 4326, 32632)
    got = xrs.dev.to_host(xrs.rep.ReprojectPlan(src_gm, tgt_gm).run(xrs.dev.to_device(data), method, nan))
    assert got.shape == want.shape and got.dtype == want.dtype
    if method == "nearest":
        frac = float(np.mean(got != want))
        print(f"nearest mismatch fraction (rounding ties): {frac:.3g}")
        assert frac < 1e-4
    else:
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("src_epsg,tgt_epsg,tgt_args", [
    (32632, 3035, dict(size=(120, 90), xy_min=(4320000.0, 3380000.0), xy_res=25.0)),
    (3035, 4326, dict(size=(100, 100), xy_min=(6.0, 48.0), xy_res=0.002)),
    (4326, 3857, dict(size=(128, 96), xy_min=(1100000.0, 6100000.0), xy_res=150.0)),
    (3857, 32632, dict(size=(96, 128), xy_min=(560000.0, 5930000.0), xy_res=30.0)),
])
def test_other_projection_pairs_match_oracle(xrs, src_epsg, tgt_epsg, tgt_args):
    tgt = xrs.GridMapping.regular(crs=f"EPSG:{tgt_epsg}", tile_size=(64, 48), **tgt_args)
    box = np.array(oproj.transform_bounds(oproj.from_epsg(tgt_epsg), oproj.from_epsg(src_epsg), *tgt.xy_bbox))
    w, h = 140, 120
    res = max((box[2] - box[0]) / (w - 10), (box[3] - box[1]) / (h - 10))
    src = xrs.GridMapping.regular((w, h), (box[0] - 5 * res, box[1] - 5 * res), float(res), f"EPSG:{src_epsg}")
    rng = np.random.default_rng(2)
    data = rng.random((2, h, w)).astype(np.float32)
    plan = xrs.rep.ReprojectPlan(src, tgt)
    for method in ("nearest", "bilinear"):
        want = _oracle_projected(src, tgt, data, method, nan, src_epsg, tgt_epsg)
        got = xrs.dev.to_host(plan.run(xrs.dev.to_device(data), method, nan))
        if method == "nearest":
            assert float(np.mean(got != want)) < 1e-3
        else:
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9, equal_nan=True)


# ---------------------------------------------------------------------------
# the reference's tests/test_reproject.py through reproject_dataset
# ---------------------------------------------------------------------------
def _ds_5x5(xrs, three_d=False):
    """tests/sampledata.py:95-128."""
    x = np.arange(565300.0, 565800.0, 100.0)
    y = np.arange(5934300.0, 5933800.0, -100.0)
    band = np.arange(25).reshape((5, 5))
    crs_attrs = xrs.CRS.from_epsg(32632).to_cf()
    if three_d:
        band = np.repeat(band[np.newaxis], 2, axis=0)
        dims = ("time", "y", "x")
    else:
        dims = ("y", "x")
    return xrs.Dataset(
        data_vars=dict(band_1=xrs.DataArray(band, dims=dims, attrs=dict(grid_mapping="spatial_ref"))),
        coords=dict(x=xrs.DataArray(x, dims="x"), y=xrs.DataArray(y, dims="y"),
                    spatial_ref=xrs.DataArray(np.array(0), dims=(), attrs=crs_attrs)))


EXPECTED_5X5 = np.array([[1, 1, 2, 3, 4], [6, 6, 7, 8, 9], [11, 12, 12, 13, 14], [16, 17, 17, 18, 19],
                         [21, 17, 17, 18, 19]])


def test_ref_reproject_target_gm(xrs):
    """tests/test_reproject.py:21-39."""
    tgt = xrs.GridMapping.regular(size=(5, 5), xy_min=(4320080, 3382480), xy_res=80, crs="epsg:3035")
    out = xrs.reproject_dataset(_ds_5x5(xrs), tgt)
    assert out.band_1.dtype == np.int64
    np.testing.assert_array_equal(out.band_1.values, EXPECTED_5X5)
    assert "spatial_ref" in out.coords


def test_ref_reproject_target_gm_3d(xrs):
    """tests/test_reproject.py:41-76."""
    tgt = xrs.GridMapping.regular(size=(5, 5), xy_min=(4320080, 3382480), xy_res=80, crs="epsg:3035")
    src = _ds_5x5(xrs, three_d=True)
    out = xrs.reproject_dataset(src, tgt)
    np.testing.assert_array_equal(out.band_1.values, np.stack([EXPECTED_5X5, EXPECTED_5X5]))
    assert out.band_1.dims == ("time", "y", "x")


def test_ref_reproject_j_axis_up(xrs):
    """tests/test_reproject.py:78-120: target j-axis up, and source j-axis up."""
    tgt = xrs.GridMapping.regular(size=(5, 5), xy_min=(4320080, 3382480), xy_res=80, crs="epsg:3035",
                                  is_j_axis_up=True)
    np.testing.assert_array_equal(xrs.reproject_dataset(_ds_5x5(xrs), tgt).band_1.values, EXPECTED_5X5[::-1])
    src = _ds_5x5(xrs)
    flipped = xrs.Dataset(
        data_vars=dict(band_1=xrs.DataArray(src.band_1.values[::-1], dims=("y", "x"), attrs=src.band_1.attrs)),
        coords=dict(x=src.coords["x"], y=xrs.DataArray(src.coords["y"].values[::-1], dims="y"),
                    spatial_ref=src.coords["spatial_ref"]))
    tgt = xrs.GridMapping.regular(size=(5, 5), xy_min=(4320080, 3382480), xy_res=80, crs="epsg:3035")
    np.testing.assert_array_equal(xrs.reproject_dataset(flipped, tgt).band_1.values, EXPECTED_5X5)


def test_ref_reproject_finer_and_coarser(xrs):
    """tests/test_reproject.py:122-160."""
    tgt = xrs.GridMapping.regular(size=(5, 5), xy_min=(4320080, 3382480), xy_res=20, crs="epsg:3035")
    np.testing.assert_array_equal(xrs.reproject_dataset(_ds_5x5(xrs), tgt).band_1.values, [
        [15, 16, 16, 16, 16], [15, 16, 16, 16, 16], [15, 16, 16, 16, 16], [20, 21, 21, 21, 21],
        [20, 21, 21, 21, 21]])
    tgt = xrs.GridMapping.regular(size=(3, 3), xy_min=(4320050, 3382500), xy_res=120, crs="epsg:3035")
    np.testing.assert_array_equal(xrs.reproject_dataset(_ds_5x5(xrs), tgt).band_1.values, [
        [0, 1, 2], [5, 6, 7], [15, 16, 17]])


def test_ref_reproject_geographic_targets(xrs):
    """tests/test_reproject.py:162-201."""
    tgt = xrs.GridMapping.regular(size=(5, 5), xy_min=(9.9886, 53.5499), xy_res=0.0006, crs=xrs.CRS_WGS84)
    np.testing.assert_array_equal(xrs.reproject_dataset(_ds_5x5(xrs), tgt).band_1.values, [
        [7, 8, 8, 8, 9], [12, 13, 13, 13, 14], [12, 13, 13, 13, 14], [17, 18, 18, 18, 19], [22, 23, 23, 23, 24]])
    tgt = xrs.GridMapping.regular(size=(5, 5), xy_min=(9.9886, 53.5499), xy_res=0.0003, crs=xrs.CRS_WGS84)
    np.testing.assert_array_equal(xrs.reproject_dataset(_ds_5x5(xrs), tgt).band_1.values, [
        [12, 12, 12, 13, 13], [17, 17, 17, 18, 18], [17, 17, 17, 18, 18], [22, 17, 17, 18, 18],
        [22, 22, 22, 23, 23]])


def test_ref_reproject_complex_array(xrs):
    """tests/test_reproject.py:203-245 (values to 4 decimals; chunk-layout assertions do not apply
    to an eager result)."""
    nt, nx, ny = 10, 100, 100
    x = np.linspace(3900000, 4500000, nx)
    y = np.linspace(2600000, 3200000, ny)
    temp = np.arange(nt * nx * ny, dtype=np.float32).reshape(nt, nx, ny)
    ds = xrs.Dataset(
        data_vars=dict(temperature=xrs.DataArray(temp, dims=("time", "y", "x"), attrs=dict(grid_mapping="spatial_ref")),
                       onedim_data=xrs.DataArray(np.arange(nt), dims="time")),
        coords=dict(time=xrs.DataArray(np.arange(nt), dims="time"), x=xrs.DataArray(x, dims="x"),
                    y=xrs.DataArray(y, dims="y"),
                    spatial_ref=xrs.DataArray(np.array(0), dims=(), attrs=xrs.CRS.from_epsg(3035).to_cf())))
    tgt = xrs.GridMapping.regular(size=(10, 10), xy_min=(6.0, 48.0), xy_res=0.2, crs=xrs.CRS_WGS84, tile_size=(5, 5))
    out = xrs.reproject_dataset(ds, tgt, interp_methods="triangular")
    assert sorted(out.data_vars) == ["onedim_data", "temperature"]
    assert abs(float(out.temperature.values[0, 0, 0]) - 6353.582) < 5e-4
    assert abs(float(out.temperature.values[0, -1, -1]) - 3007.1228) < 5e-4
    out = xrs.reproject_dataset(ds, tgt, interp_methods=1)
    assert abs(float(out.temperature.values[0, 0, 0]) - 6353.5823) < 5e-5
    assert abs(float(out.temperature.values[0, -1, -1]) - 3007.1228) < 5e-5


def test_ref_reproject_raise_not_implemented(xrs):
    """tests/test_reproject.py:247-258."""
    tgt = xrs.GridMapping.regular(size=(5, 5), xy_min=(4320080, 3382480), xy_res=20, crs="epsg:3035")
    with pytest.raises(NotImplementedError, match="interp_methods must be one of 0, 1, 'nearest', 'bilinear', 'triangular'"):
        xrs.reproject_dataset(_ds_5x5(xrs), tgt, interp_methods="cubic")


def test_resample_in_space_dispatches_to_reproject(xrs):
    """spatial.py:158-168."""
    tgt = xrs.GridMapping.regular(size=(5, 5), xy_min=(4320080, 3382480), xy_res=80, crs="epsg:3035")
    out = xrs.resample_in_space(_ds_5x5(xrs), target_gm=tgt)
    np.testing.assert_array_equal(out.band_1.values, EXPECTED_5X5)


@pytest.fixture
def staged_k3(monkeypatch):
    """The TMA-staged reproject kernel is an experiment that measured slower than the direct kernel
    (csrc/reproject.cu); it is compiled only with -DXRS_K3_STAGED_EXPERIMENT and selected by
    XRS_K3_STAGED=1.  In a default build the switch is ignored and both runs below take the direct
    kernel (pitched and unpitched sources), which is still a useful comparison."""
    monkeypatch.setenv("XRS_K3_STAGED", "1")
    yield


def test_band_pipeline_equals_plain_path(xrs):
    """reproject_dataset streams every variable through the device in band chunks (1, 2, 4, ... bands,
    upload / kernel / download overlapped); same bytes as one plain kernel call on the whole stack."""
    src_gm, tgt_gm = _case_utm_from_geographic(xrs, n=160, tile=64)
    rng = np.random.default_rng(8)
    data = rng.random((7, src_gm.height, src_gm.width)).astype(np.float32)
    ds = xrs.Dataset(data_vars=dict(v=xrs.DataArray(data, dims=("band", "lat", "lon"))),
                     coords=dict(lon=xrs.DataArray(src_gm.x_values, dims="lon"),
                                 lat=xrs.DataArray(src_gm.y_values, dims="lat")))
    plan = xrs.rep.ReprojectPlan(src_gm, tgt_gm)
    sd = xrs.dev.to_device(data)
    for method in ("nearest", "bilinear"):
        plain = xrs.dev.to_host(plan.run(sd, method, np.nan))
        piped = xrs.reproject_dataset(ds, tgt_gm, source_gm=src_gm, interp_methods=method)["v"].values
        assert_same(piped, plain, method)


@pytest.mark.parametrize("tile", [128, (100, 77)])
@pytest.mark.parametrize("case", ["utm", "webmerc", "laea", "coarse_target"])
def test_staged_kernel_equals_direct_kernel(xrs, case, tile, staged_k3):
    """k3_reproject_staged (source box of every band staged through TMA) against the direct per-pixel
    kernel on the same inputs, bit for bit: real projections, reference tiles that are not multiples
    of the 32 x 32 CTA tile, float64 and float32 outputs, a row band with a resident window, and a
    target so coarse that the box does not fit the staging buffers (in-kernel fallback)."""
    if case == "utm":
        src_gm, tgt_gm = _case_utm_from_geographic(xrs, n=300, tile=tile)
    elif case == "webmerc":
        src_gm = xrs.GridMapping.regular((420, 300), (5.0, 44.0), 0.01, "EPSG:4326")
        tgt_gm = xrs.GridMapping.regular((400, 380), (570000.0, 5500000.0), 1113.2, "EPSG:3857", tile_size=tile)
    elif case == "laea":
        src_gm = xrs.GridMapping.regular((420, 300), (9.0, 47.0), 0.01, "EPSG:4326")
        tgt_gm = xrs.GridMapping.regular((500, 420), (4180000.0, 2660000.0), 600.0, "EPSG:3035", tile_size=tile)
    else:
        src_gm = xrs.GridMapping.regular((900, 700), (9.0, 47.0), 0.002, "EPSG:4326")
        tgt_gm = xrs.GridMapping.regular((200, 180), (4250000.0, 2700000.0), 700.0, "EPSG:3035", tile_size=tile)
    rng = np.random.default_rng(5)
    data = rng.random((5, src_gm.height, src_gm.width)).astype(np.float32)
    data[1, 40:60, 50:90] = nan
    plain, pitched = xrs.dev.to_device(data), xrs.dev.to_device_pitched(data)
    if (plain.stride(1) * 4) % 16 == 0:
        plain = xrs.dev.to_device(np.pad(data, ((0, 0), (0, 0), (0, 1))))[:, :, :-1]  # force the direct kernel
    plan = xrs.rep.ReprojectPlan(src_gm, tgt_gm)
    for method in ("nearest", "bilinear", "triangular"):
        for out_dtype in ((None, np.float32) if method == "bilinear" else (None,)):
            want = xrs.dev.to_host(plan.run(plain, method, nan, out_dtype=out_dtype))
            got = xrs.dev.to_host(plan.run(pitched, method, nan, out_dtype=out_dtype))
            _same_up_to_contraction(got, want, method, f"{case} {method} out={out_dtype}")
            assert np.isfinite(got).mean() > 0.2
    rows = (64, 230) if tgt_gm.height > 300 else (32, 150)
    band = xrs.rep.ReprojectPlan(src_gm, tgt_gm, rows=rows)
    i0, j0, i1, j1 = band.footprint()
    full = xrs.dev.to_host(plan.run(pitched, "bilinear", nan))
    part = band.run(xrs.dev.to_device_pitched(np.ascontiguousarray(data[:, j0:j1, i0:i1])), "bilinear", nan,
                    window_origin=(i0, j0))
    assert_same(xrs.dev.to_host(part), full[:, rows[0]:rows[1]], f"{case} row band with resident window")


def _same_up_to_contraction(got, want, method, what):
    """Two kernels evaluating the same projection formulas: reproject.cu is compiled with FMA contraction,
    so the transformed coordinates may differ in the last bits between them (far below the 1e-6 of the
    north star); everything after the transform is contraction-free."""
    assert got.shape == want.shape and got.dtype == want.dtype, what
    if method == "nearest":
        frac = float(np.mean(~((got == want) | (np.isnan(got) & np.isnan(want)))))
        assert frac < 1e-4, (what, frac)
    else:
        # float32 outputs: a last-bit difference of the float64 value can cross a float32 rounding boundary
        rtol = 1e-9 if got.dtype == np.float64 else 2.5e-7
        np.testing.assert_allclose(got, want, rtol=rtol, atol=1e-12, equal_nan=True, err_msg=what)


# ---------------------------------------------------------------------------
# lattice form of the transform (csrc/reproject.cu: k3_lattice_setup) against the per-pixel formulas
# ---------------------------------------------------------------------------
def _lattice_case(xrs, case):
    if case == "utm":  # scaled-down C3
        return _case_utm_from_geographic(xrs, n=300, tile=128)
    if case == "laea":  # target LAEA: the generic per-pixel plan
        src_gm = xrs.GridMapping.regular((700, 500), (9.0, 47.0), 0.0005, "EPSG:4326")
        tgt_gm = xrs.GridMapping.regular((500, 420), (4250000.0, 2660000.0), 30.0, "EPSG:3035", tile_size=(100, 77))
        return src_gm, tgt_gm
    if case == "utm_source":  # geographic target, projected source: forward projection per pixel
        tgt_gm = xrs.GridMapping.regular((400, 300), (9.02, 48.01), 0.0002, "EPSG:4326", tile_size=128)
        box = oproj.transform_bounds(oproj.from_epsg(4326), oproj.from_epsg(32632), *tgt_gm.xy_bbox)
        src_gm = xrs.GridMapping.regular((int((box[2] - box[0]) / 20.0) + 8, int((box[3] - box[1]) / 20.0) + 8),
                                         (box[0] - 60.0, box[1] - 60.0), 20.0, "EPSG:32632")
        return src_gm, tgt_gm
    raise AssertionError(case)


@pytest.mark.parametrize("case", ["utm", "laea", "utm_source"])
def test_lattice_transform_equals_exact_transform(xrs, case, monkeypatch):
    """The CTA-level bicubic interpolation of the transform against XRS_K3_EXACT=1 (full formulas per
    pixel): coordinates agree to ~1e-9 px, so float64 bilinear outputs agree to 1e-7 of the data range
    -- and differ somewhere in the last bits, which shows that the lattice path is the one that ran."""
    src_gm, tgt_gm = _lattice_case(xrs, case)
    rng = np.random.default_rng(11)
    data = rng.random((3, src_gm.height, src_gm.width)).astype(np.float32)
    sd = xrs.dev.to_device(data)
    plan = xrs.rep.ReprojectPlan(src_gm, tgt_gm)
    out = {}
    for exact in ("1", "0"):
        monkeypatch.setenv("XRS_K3_EXACT", exact)
        out[exact] = {m: xrs.dev.to_host(plan.run(sd, m, nan)) for m in ("nearest", "bilinear")}
    got, want = out["0"], out["1"]
    assert np.isfinite(want["bilinear"]).mean() > 0.5
    assert got["bilinear"].dtype == np.float64
    np.testing.assert_allclose(got["bilinear"], want["bilinear"], rtol=0, atol=1e-7, equal_nan=True)
    assert np.any(got["bilinear"] != want["bilinear"]), "lattice path did not run"
    frac = float(np.mean(~((got["nearest"] == want["nearest"]) | (np.isnan(got["nearest"]) & np.isnan(want["nearest"])))))
    assert frac < 1e-4, frac


def test_lattice_falls_back_to_exact_on_coarse_grids(xrs, monkeypatch):
    """5 km pixels: the cubic through a 320 km tile misses the exact centre by far more than 1e-8 px,
    so every CTA takes the per-pixel formulas -- bit-identical to XRS_K3_EXACT=1."""
    src_gm = xrs.GridMapping.regular((900, 700), (-10.0, 30.0), 0.05, "EPSG:4326")
    tgt_gm = xrs.GridMapping.regular((300, 280), (3200000.0, 1700000.0), 5000.0, "EPSG:3035", tile_size=128)
    rng = np.random.default_rng(12)
    data = rng.random((2, src_gm.height, src_gm.width)).astype(np.float32)
    sd = xrs.dev.to_device(data)
    plan = xrs.rep.ReprojectPlan(src_gm, tgt_gm)
    monkeypatch.setenv("XRS_K3_EXACT", "1")
    want = xrs.dev.to_host(plan.run(sd, "bilinear", nan))
    monkeypatch.setenv("XRS_K3_EXACT", "0")
    got = xrs.dev.to_host(plan.run(sd, "bilinear", nan))
    assert np.isfinite(want).mean() > 0.5
    assert_same(got, want, "coarse grid: exact path")


def test_lattice_is_independent_of_the_row_band(xrs):
    """CTA tiles are anchored at absolute target rows, so a row band that does not start at a multiple
    of the tile height gives the same bits as the whole image."""
    src_gm, tgt_gm = _case_utm_from_geographic(xrs, n=300, tile=128)
    rng = np.random.default_rng(13)
    data = rng.random((4, src_gm.height, src_gm.width)).astype(np.float32)
    sd = xrs.dev.to_device(data)
    full = xrs.dev.to_host(xrs.rep.ReprojectPlan(src_gm, tgt_gm).run(sd, "bilinear", nan))
    for rows in ((45, 211), (1, 33), (290, 300)):
        part = xrs.dev.to_host(xrs.rep.ReprojectPlan(src_gm, tgt_gm, rows=rows).run(sd, "bilinear", nan))
        assert_same(part, full[:, rows[0]:rows[1]], f"rows {rows}")

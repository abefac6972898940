"""GPU parity of the affine / coarsen path (xrs_affine, xrs_coarsen through ctypes) against the
oracle: scipy.ndimage.affine_transform + numpy reducers, i.e. what the reference executes.

float results: bit-exact expected for order 0/1 resampling, min/max/median/first/last/center and
for mean with factors 2, 4, 8 (numpy's summation order is replicated); 1e-6 relative is the
north-star tolerance and is asserted as the fallback bound.  Integer results: bit-exact.
"""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import resample as ores

from .helpers import assert_same

pytestmark = pytest.mark.gpu
nan = np.nan


@pytest.fixture(scope="module")
def xrs():
    import torch

    assert torch.cuda.is_available()
    import xcube_resampling_b200 as pkg
    from xcube_resampling_b200 import _dev, affine

    pkg.dev = _dev
    pkg.aff = affine
    return pkg


def _gpu_resample(xrs, arr, matrix, out_shape, interp, agg, fill):
    out = xrs.aff._resample_array_dev(xrs.dev.to_device(arr), matrix, out_shape[-2:], interp, agg, False, fill)
    return xrs.dev.to_host(out)


REFL = np.array([[0, 1, 0, 2, 0, 3, 0, 4], [2, 0, 3, 0, 4, 0, 1, 0], [0, 4, 0, nan, 0, 2, 0, 3],
                 [1, 0, 2, 0, 3, 0, 4, 0], [0, 3, 0, 4, 0, 1, 0, 2], [4, 0, 1, 0, 2, 0, 3, 0]], dtype=np.float64)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16, np.int32, np.uint16])
@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("matrix", [
    ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0)),
    ((0.7, 0.0, -2.2), (0.0, 1.3, 0.1)),
    ((2 / 3, 0.0, 3.0), (0.0, 0.5, -6.0)),
    ((1.0, 0.0, 0.5), (0.0, 1.0, 2.499999999999986)),
    ((0.37, 0.0, 5.5), (0.0, 0.91, 1e-9)),
])
def test_upscale_matches_scipy(xrs, dtype, order, matrix):
    rng = np.random.default_rng(3)
    if np.issubdtype(dtype, np.floating):
        a = (rng.random((37, 41)) * 100).astype(dtype)
        a[5, 7] = nan
        a[20, 3] = np.inf
        fill = nan
    else:
        info = np.iinfo(dtype)
        a = rng.integers(max(info.min, -1000), min(info.max, 1000), (37, 41)).astype(dtype)
        fill = ores.default_fill(np.dtype(dtype))
    shape = (50, 45)
    ref = ores.upscale(a, matrix, shape, order, False, fill)
    got = _gpu_resample(xrs, a, matrix, shape, order, "mean", fill)
    assert_same(got, ref, f"{np.dtype(dtype)} order {order}")


def test_3d_slice_blend_matches_scipy(xrs):
    rng = np.random.default_rng(5)
    for n, bad in ((3, (1, 4, 4)), (3, (2, 4, 4)), (3, (0, 4, 4)), (2, (1, 2, 2)), (1, (0, 3, 3)), (4, (3, 5, 1))):
        a = rng.random((n, 10, 12)).astype(np.float32)
        a[bad] = nan
        m = ((0.9, 0.0, 0.3), (0.0, 1.0, 0.0))
        ref = ores.upscale(a, m, (n, 11, 13), 1, False, nan)
        got = _gpu_resample(xrs, a, m, (n, 11, 13), 1, "mean", nan)
        assert_same(got, ref, f"n={n} bad={bad}")


@pytest.mark.parametrize("agg", ["mean", "min", "max", "median", "first", "last", "center", "sum", "prod", "count",
                                 "std", "var"])
@pytest.mark.parametrize("f", [2, 3, 4, 8, (2, 4), 16])
def test_coarsen_float32_matches_numpy(xrs, agg, f):
    f_j, f_i = (f, f) if isinstance(f, int) else f
    rng = np.random.default_rng(11)
    a = rng.random((6 * f_j, 9 * f_i)).astype(np.float32)
    a[rng.random(a.shape) < 0.05] = nan
    a[0:f_j, 0:f_i] = nan  # one all-NaN window
    ref = ores.coarsen(a, f_j, f_i, agg)
    got = xrs.dev.to_host(xrs.aff.coarsen_dev(xrs.dev.to_device(a), (f_j, f_i), agg))
    assert got.shape == ref.shape
    if agg in ("std", "var", "prod") or (agg in ("mean", "sum") and f_i not in (2, 4, 8, 16) and f_j * f_i > 64):
        np.testing.assert_allclose(got, ref.astype(got.dtype), rtol=1e-6, atol=0, equal_nan=True)
    else:
        assert_same(got, np.asarray(ref).astype(got.dtype), f"{agg} f={f}")


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int32])
@pytest.mark.parametrize("agg", ["mean", "min", "max", "median", "mode", "first", "last", "center", "sum", "count",
                                 "std", "var"])
@pytest.mark.parametrize("f", [2, 4, 8])
def test_coarsen_integers_match_numpy(xrs, dtype, agg, f):
    rng = np.random.default_rng(13)
    coarse = rng.integers(0, 20, (5, 7))
    a = np.repeat(np.repeat(coarse, 11, axis=0), 11, axis=1)[: 6 * f, : 8 * f].astype(dtype)
    a[rng.random(a.shape) < 0.2] = 3
    ref = np.asarray(ores.coarsen(a, f, f, agg))
    got = xrs.dev.to_host(xrs.aff.coarsen_dev(xrs.dev.to_device(a), (f, f), agg))
    assert np.array_equal(got.astype(np.int64), ref.astype(np.int64)), f"{agg}: {got} vs {ref}"
    if agg in ("mode", "count", "sum"):
        assert got.dtype == np.int64


@pytest.mark.parametrize("agg", ["mean", "median", "max", "min"])
@pytest.mark.parametrize("scale", [2.0, 4.0, 2.5, 8.0])
def test_downscale_matches_oracle(xrs, agg, scale):
    rng = np.random.default_rng(17)
    a = rng.random((96, 120)).astype(np.float32)
    a[rng.random(a.shape) < 0.01] = nan
    matrix = ((scale, 0.0, 1.0), (0.0, scale, -2.0))
    shape = (int(96 / scale), int(120 / scale))
    ref = ores.resample_array(a, matrix, shape, 1, agg, False, nan)
    got = _gpu_resample(xrs, a, matrix, shape, 1, agg, nan)
    assert_same(got, np.asarray(ref).astype(np.float32), f"{agg} scale {scale}")


# ---------------------------------------------------------------------------
# entry point, mirroring the reference's tests/test_affine.py
# ---------------------------------------------------------------------------
def _source_ds(xrs, three_d=False):
    res = 0.1
    data = np.stack([REFL, REFL]) if three_d else REFL
    dims = ("time", "lat", "lon") if three_d else ("lat", "lon")
    coords = dict(lon=50.0 + res * np.arange(0, 8) + 0.5 * res, lat=10.6 - res * np.arange(0, 6) - 0.5 * res)
    if three_d:
        coords["time"] = np.array([0, 1])
    return xrs.Dataset(data_vars=dict(refl=(dims, data)), coords=coords)


AFFINE_CASES = [
    # (size, xy_min, res factor, kwargs, expected) -- tests/test_affine.py:46-478
    ((3, 3), (50.0, 10.0), 1, {}, [[1, 0, 2], [0, 3, 0], [4, 0, 1]]),
    ((3, 3), (50.1, 10.1), 1, {}, [[4, nan, nan], [0, 2, 0], [3, 0, 4]]),
    ((3, 3), (50.05, 10.05), 1, {}, [[1.25, 1.5, nan], [1.0, 1.25, 1.5], [1.75, 1.0, 1.25]]),
    ((8, 6), (50, 10), 2, {}, [[nan] * 8, [nan] * 8, [nan] * 8,
                               [0.75, 1.0, 1.75, 1.25, nan, nan, nan, nan],
                               [1.25, 1.0, 1.25, 1.75, nan, nan, nan, nan],
                               [1.75, 1.25, 0.75, 1.25, nan, nan, nan, nan]]),
    ((8, 6), (49.8, 9.8), 2, {}, [[nan] * 8, [nan] * 8,
                                  [nan, 0.75, 1.0, 1.75, 1.25, nan, nan, nan],
                                  [nan, 1.25, 1.0, 1.25, 1.75, nan, nan, nan],
                                  [nan, 1.75, 1.25, 0.75, 1.25, nan, nan, nan], [nan] * 8]),
    ((8, 6), (50, 10), 0.5, {}, [[1.0, 0.5, 0.0, 1.0, 2.0, 1.0, 0.0, 1.5],
                                 [0.5, 1.0, 1.5, 1.25, 1.0, 1.5, 2.0, 1.75],
                                 [0.0, 1.5, 3.0, 1.5, 0.0, 2.0, 4.0, 2.0],
                                 [2.0, 1.75, 1.5, 1.0, 0.5, 1.25, 2.0, 1.5],
                                 [4.0, 2.0, 0.0, 0.5, 1.0, 0.5, 0.0, 1.0], [nan] * 8]),
    ((8, 6), (50.2, 10.1), 1, {}, [[nan] * 8, [0.0, 2.0, 0.0, 3.0, 0.0, 4.0, nan, nan],
                                   [nan, nan, 4.0, 0.0, 1.0, 0.0, nan, nan], [nan, nan, 0.0, 2.0, 0.0, 3.0, nan, nan],
                                   [2.0, 0.0, 3.0, 0.0, 4.0, 0.0, nan, nan], [0.0, 4.0, 0.0, 1.0, 0.0, 2.0, nan, nan]]),
    ((8, 6), (49.8, 9.9), 1, {}, [[nan, nan, 2.0, 0.0, nan, nan, 4.0, 0.0], [nan, nan, 0.0, 4.0, nan, nan, 0.0, 2.0],
                                  [nan, nan, 1.0, 0.0, 2.0, 0.0, 3.0, 0.0], [nan, nan, 0.0, 3.0, 0.0, 4.0, 0.0, 1.0],
                                  [nan, nan, 4.0, 0.0, 1.0, 0.0, 2.0, 0.0], [nan] * 8]),
]


@pytest.mark.parametrize("size,xy_min,res_factor,kwargs,expected", AFFINE_CASES)
def test_affine_transform_dataset_reference_cases(xrs, size, xy_min, res_factor, kwargs, expected):
    ds = _source_ds(xrs)
    source_gm = xrs.GridMapping.from_dataset(ds)
    assert source_gm.is_regular and not source_gm.is_j_axis_up
    target_gm = xrs.GridMapping.regular(size, xy_min, 0.1 * res_factor, source_gm.crs)
    out = xrs.affine_transform_dataset(ds, target_gm, source_gm=source_gm, interp_methods=1, **kwargs)
    assert set(out.variables) == {"refl", "lon", "lat", "spatial_ref"}
    np.testing.assert_almost_equal(out["refl"].values, np.array(expected, dtype=np.float64))


def test_affine_transform_dataset_3d_and_mapping_options(xrs):
    # tests/test_affine.py:142-246
    ds = _source_ds(xrs, three_d=True)
    target_gm = xrs.GridMapping.regular((3, 3), (50.0, 10.0), 0.1, "EPSG:4326")
    out = xrs.affine_transform_dataset(ds, target_gm, interp_methods={"refl": "bilinear"})
    exp = np.array([[1, 0, 2], [0, 3, 0], [4, 0, 1]], dtype=np.float64)
    np.testing.assert_almost_equal(out["refl"].values, np.stack([exp, exp]))


def test_affine_transform_dataset_crs_rules(xrs):
    # tests/test_affine.py:248-293
    ds = _source_ds(xrs)
    source_gm = xrs.GridMapping.from_dataset(ds)
    expected = np.array([[1.25, 1.5, nan], [1.0, 1.25, 1.5], [1.75, 1.0, 1.25]])
    for crs in (xrs.CRS_WGS84, xrs.CRS_CRS84):
        target_gm = xrs.GridMapping.regular((3, 3), (50.05, 10.05), 0.1, crs)
        out = xrs.affine_transform_dataset(ds, target_gm, source_gm=source_gm, interp_methods=1)
        np.testing.assert_almost_equal(out["refl"].values, expected)
    target_gm = xrs.GridMapping.regular((3, 3), (50.05, 10.05), 0.1, "EPSG:3035")
    with pytest.raises(AssertionError) as e:
        xrs.affine_transform_dataset(ds, target_gm, source_gm=source_gm)
    assert ("Affine transformation cannot be applied to source CRS 'WGS 84' and target CRS "
            "'ETRS89-extended / LAEA Europe'") in str(e.value)


def test_affine_order_above_one_raises_value_error(xrs):
    # tests/test_affine.py:480-497
    ds = _source_ds(xrs)
    target_gm = xrs.GridMapping.regular((8, 6), (50.2, 10.1), 0.1, "EPSG:4326")
    with pytest.raises(ValueError) as e:
        xrs.affine_transform_dataset(ds, target_gm, interp_methods=3)
    assert "interp_methods must be one of 0, 1, 'nearest', 'bilinear'." in str(e.value)


def test_config1_bilinear_downsample_4096(xrs):
    """BASELINE config 1: 2x bilinear downsample of a 4096^2 float32 variable (identity + 2x2 nanmean)."""
    rng = np.random.default_rng(0)
    n = 1024  # quarter linear scale keeps the scipy oracle fast; the kernel path is size independent
    a = rng.random((n, n)).astype(np.float32)
    src = ogrid.regular_grid((n, n), (0, 0), 0.01, tile_size=256)
    tgt = ogrid.regular_grid((n // 2, n // 2), (0, 0), 0.02, tile_size=256)
    ref = ores.affine_transform(a, src, tgt, interp=1)
    source_gm = xrs.GridMapping.regular((n, n), (0, 0), 0.01, "EPSG:4326", tile_size=256)
    target_gm = xrs.GridMapping.regular((n // 2, n // 2), (0, 0), 0.02, "EPSG:4326", tile_size=256)
    ds = xrs.Dataset(data_vars=dict(refl=(("lat", "lon"), a)),
                     coords=dict(lon=source_gm.x_coords.values, lat=source_gm.y_coords.values))
    out = xrs.affine_transform_dataset(ds, target_gm, source_gm=source_gm, interp_methods=1)
    assert_same(out["refl"].values, ref, "C1")


# ---------------------------------------------------------------------------
# recover_nans (affine.py:344-360)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("matrix,shape", [
    (((0.5, 0.0, 0.25), (0.0, 0.5, -0.25)), (14, 18)),   # upscale
    (((1.0, 0.0, 0.5), (0.0, 1.0, 0.5)), (6, 8)),        # half-pixel shift
    (((2.0, 0.0, 0.0), (0.0, 2.0, 0.0)), (3, 4)),        # aligned downscale -> aggregation
    (((2.5, 0.0, 1.0), (0.0, 2.5, -1.0)), (3, 3)),       # fractional downscale -> divisor 3
])
def test_recover_nans_matches_oracle(xrs, dtype, matrix, shape):
    rng = np.random.default_rng(23)
    a = rng.random((7, 9)).astype(dtype)
    a[2, 3] = nan
    a[5, 0] = nan
    a[0, 8] = np.inf
    for agg in ("mean", "max"):
        ref = np.asarray(ores.resample_array(a, matrix, shape, 1, agg, True, nan))
        got = _gpu_resample_recover(xrs, a, matrix, shape, agg)
        assert got.dtype == ref.dtype == np.float64
        assert_same(got, ref, f"{np.dtype(dtype).name} {agg} {matrix}")
    # no NaN in the array: the ordinary path runs and the dtype stays (affine.py:349)
    b = np.nan_to_num(a, nan=0.5, posinf=1.0)
    got = _gpu_resample_recover(xrs, b, matrix, shape, "mean")
    assert got.dtype == np.dtype(dtype)
    assert_same(got, np.asarray(ores.resample_array(b, matrix, shape, 1, "mean", True, nan)).astype(dtype), "no-nan")


def _gpu_resample_recover(xrs, arr, matrix, out_shape, agg):
    out = xrs.aff._resample_array_dev(xrs.dev.to_device(arr), matrix, out_shape[-2:], 1, agg, True, nan)
    return xrs.dev.to_host(out)


def test_recover_nans_3d_and_reference_expectation(xrs):
    """tests/test_affine.py:118-140: the NaN neighbour is recovered as 0.6666667."""
    rng = np.random.default_rng(29)
    a = rng.random((3, 6, 7))
    a[1, 2, 2] = nan
    matrix = ((0.7, 0.0, 0.3), (0.0, 0.7, 0.1))
    ref = np.asarray(ores.resample_array(a, matrix, (3, 8, 9), 1, "mean", True, nan))
    got = xrs.dev.to_host(xrs.aff._resample_array_dev(xrs.dev.to_device(a), matrix, (8, 9), 1, "mean", True, nan))
    assert_same(got, ref, "3-D recover")
    ds = _source_ds(xrs)
    target_gm = xrs.GridMapping.regular((3, 3), (50.05, 10.05), 0.1, "EPSG:4326")
    out = xrs.affine_transform_dataset(ds, target_gm, interp_methods=1, recover_nans=True)
    np.testing.assert_almost_equal(out["refl"].values, [[1.25, 1.5, 0.6666667], [1.0, 1.25, 1.5], [1.75, 1.0, 1.25]])

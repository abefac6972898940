"""Streaming block-aggregation kernels (resample_fast.cu) against scipy + numpy, bit for bit.

These are the aligned integer-factor cases of xrs_affine / xrs_coarsen (windows 2x2, 4x4, 8x8 whose
samples fall exactly on source pixels).  Covered here beyond tests/test_resample_gpu.py: order-1
zero-weight contamination by non-finite right / lower neighbours incl. scipy's edge mirror, -0.0,
integer offsets, unaligned pitches (scalar-load path), 3-D arrays, partially / fully NaN windows in
the sorting-network median, and that the fast path is really the one that runs."""

import numpy as np
import pytest

from oracle import resample as ores

from .helpers import assert_same

pytestmark = pytest.mark.gpu
nan, inf = np.nan, np.inf


@pytest.fixture(scope="module")
def xrs():
    import torch

    assert torch.cuda.is_available()
    import xcube_resampling_b200 as pkg
    from xcube_resampling_b200 import _dev, _lib, affine

    pkg.dev, pkg.aff, pkg.lib = _dev, affine, _lib
    return pkg


def _kernels_used(xrs, fn):
    xrs.lib.profile_collect()
    xrs.lib.profile_enable(True)
    try:
        out = fn()
    finally:
        xrs.lib.profile_enable(False)
    return out, set(xrs.lib.profile_collect())


@pytest.mark.parametrize("f", [2, 4, 8])
@pytest.mark.parametrize("agg", ["mean", "median", "min", "max", "sum", "std", "center", "count"])
def test_order1_identity_contamination(xrs, f, agg):
    """Down-scaling by an integer factor on aligned grids: scale/f == 1, offsets integer."""
    rng = np.random.default_rng(f)
    h, w = 6 * f + 3, 7 * f + 5
    a = rng.normal(size=(h, w)).astype(np.float32)
    a[rng.random(a.shape) < 0.03] = nan
    a[rng.random(a.shape) < 0.01] = inf
    a[rng.random(a.shape) < 0.01] = -inf
    a[rng.random(a.shape) < 0.02] = -0.0
    a[h - 1, 3] = nan   # last row: reached only through the mirrored tap
    a[2, w - 1] = inf   # last column likewise
    for off in ((0, 0), (3, 5), (1, 0)):
        out_h, out_w = (h - off[0]) // f, (w - off[1]) // f
        matrix = ((float(f), 0.0, float(off[1])), (0.0, float(f), float(off[0])))
        ref = ores.resample_array(a, matrix, (out_h, out_w), 1, agg, False, nan)
        got, used = _kernels_used(xrs, lambda: xrs.dev.to_host(
            xrs.aff._resample_array_dev(xrs.dev.to_device(a), matrix, (out_h, out_w), 1, agg, False, nan)))
        assert any(k.startswith("k5_window_reduce") for k in used), used
        if agg in ("std",):
            np.testing.assert_allclose(got, np.asarray(ref).astype(got.dtype), rtol=1e-6, equal_nan=True)
        else:
            assert_same(got, np.asarray(ref).astype(got.dtype), f"f={f} {agg} off={off}")


def test_exact_fit_uses_mirror_taps(xrs):
    """Windows reaching the last row / column: taps len -> len-2 (ni_interpolation.c)."""
    a = np.arange(64, dtype=np.float64).reshape(8, 8)
    a[6, 2] = nan   # mirror source of the last row's lower tap
    a[3, 6] = inf   # mirror source of the last column's right tap
    matrix = ((2.0, 0.0, 0.0), (0.0, 2.0, 0.0))
    for agg in ("mean", "max", "median"):
        ref = ores.resample_array(a, matrix, (4, 4), 1, agg, False, nan)
        got = xrs.dev.to_host(xrs.aff._resample_array_dev(xrs.dev.to_device(a), matrix, (4, 4), 1, agg, False, nan))
        assert_same(got, np.asarray(ref), agg)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int8, np.uint16, np.int16, np.int32, np.uint32,
                                   np.int64])
@pytest.mark.parametrize("f", [2, 4, 8])
def test_coarsen_all_dtypes_and_unaligned_rows(xrs, dtype, f):
    rng = np.random.default_rng(5)
    h, w = 5 * f, 9 * f + 1  # odd width -> row pitch not a multiple of the vector size
    if np.issubdtype(dtype, np.floating):
        a = rng.normal(size=(h, w)).astype(dtype)
        a[rng.random(a.shape) < 0.1] = nan
    else:
        info = np.iinfo(dtype)
        a = rng.integers(max(info.min, -1000), min(info.max, 1000), size=(h, w)).astype(dtype)
    src = a[:, : 9 * f]
    for agg in ("mean", "min", "max", "median", "first", "last"):
        ref = np.asarray(ores.coarsen(src, f, f, agg))
        dev = xrs.dev.to_device(a)[:, : 9 * f]  # strided view: pitch = w elements
        got, used = _kernels_used(xrs, lambda: xrs.dev.to_host(xrs.aff.coarsen_dev(dev, (f, f), agg)))
        if agg == "median" and dtype in (np.int8, np.uint32, np.int64):
            assert used == {"k4_affine_generic"}  # sorting network not instantiated for these
        else:
            assert any(k.startswith("k5_window_reduce") for k in used), used
        assert_same(got, ref.astype(got.dtype), f"{np.dtype(dtype).name} f={f} {agg}")


def test_three_d_coarsen_and_mode(xrs):
    rng = np.random.default_rng(9)
    a = rng.integers(0, 6, size=(3, 16, 24)).astype(np.uint8)
    for f in (2, 4, 8):
        ref = np.asarray(ores.coarsen(a, f, f, "mode"))
        got = xrs.dev.to_host(xrs.aff.coarsen_dev(xrs.dev.to_device(a), (f, f), "mode"))
        assert got.dtype == np.int64
        assert_same(got, ref.astype(np.int64), f"mode f={f}")
    b = rng.normal(size=(2, 16, 16)).astype(np.float32)
    ref = np.asarray(ores.coarsen(b, 4, 4, "mean"))
    assert_same(xrs.dev.to_host(xrs.aff.coarsen_dev(xrs.dev.to_device(b), (4, 4), "mean")), ref, "3-D mean")


def test_median_nan_patterns(xrs):
    rng = np.random.default_rng(21)
    for f in (2, 4, 8):
        a = rng.normal(size=(4 * f, 6 * f)).astype(np.float32)
        a[0:f, 0:f] = nan                      # all NaN
        a[0:f, f:2 * f].flat[1:] = nan          # a single valid value
        a[f:2 * f, 0:f].flat[::2] = nan         # half
        a[f:2 * f, f:2 * f].flat[:3] = inf      # infinities are values, not missing
        ref = np.asarray(ores.coarsen(a, f, f, "median"))
        got = xrs.dev.to_host(xrs.aff.coarsen_dev(xrs.dev.to_device(a), (f, f), "median"))
        assert_same(got, ref.astype(np.float32), f"median f={f}")


@pytest.mark.parametrize("f", [2, 4, 8])
@pytest.mark.parametrize("agg", ["mean", "min", "max", "median", "mode", "first", "last", "center", "sum", "count"])
def test_uint8_packed_path(xrs, f, agg):
    """uint8 rasters with 4-aligned output width take the SIMD-in-a-word kernel (4 windows / thread)."""
    rng = np.random.default_rng(100 + f)
    h, w = 7 * f, 16 * f * 3
    a = rng.integers(0, 256, size=(2, h, w)).astype(np.uint8)
    a[0, : 2 * f] = rng.integers(0, 4, size=(2 * f, w)).astype(np.uint8)  # few classes: mode ties, zeros for count
    a[1, 3 * f: 4 * f] = 255
    ref = np.asarray(ores.coarsen(a, f, f, agg))
    got, used = _kernels_used(xrs, lambda: xrs.dev.to_host(xrs.aff.coarsen_dev(xrs.dev.to_device(a), (f, f), agg)))
    assert used == ({"k5_u8x2_sort"} if agg in ("median", "mode") else {"k5_u8x4_reduce"}), used
    assert got.dtype == (np.int64 if agg in ("mode", "sum", "count") else np.uint8)
    assert np.array_equal(got.astype(np.int64), ref.astype(np.int64)), f"{agg} f={f}"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("agg", ["mean", "max", "min", "sum", "first", "center", "std"])
def test_blend2_four_windows_per_thread(xrs, dtype, agg):
    """2x2 windows of order-1 samples with a 4-aligned output width: k5_blend2_x4."""
    rng = np.random.default_rng(77)
    h, w = 2 * 9 + 2, 2 * 24 + 4
    a = rng.normal(size=(h, w)).astype(dtype)
    a[rng.random(a.shape) < 0.04] = nan
    a[rng.random(a.shape) < 0.02] = inf
    a[rng.random(a.shape) < 0.02] = -0.0
    a[h - 1, 5] = nan
    a[3, w - 1] = -inf
    for off, out_hw in (((0, 0), (10, 24)), ((2, 4), (9, 24)), ((0, 0), (9, 24))):
        matrix = ((2.0, 0.0, float(off[1])), (0.0, 2.0, float(off[0])))
        ref = np.asarray(ores.resample_array(a, matrix, out_hw, 1, agg, False, nan))
        got, used = _kernels_used(xrs, lambda: xrs.dev.to_host(
            xrs.aff._resample_array_dev(xrs.dev.to_device(a), matrix, out_hw, 1, agg, False, nan)))
        assert used == {"k5_blend2_x4"}, used
        if agg == "std":
            np.testing.assert_allclose(got, ref.astype(got.dtype), rtol=1e-6, equal_nan=True)
        else:
            assert_same(got, ref.astype(got.dtype), f"{agg} off={off} out={out_hw}")

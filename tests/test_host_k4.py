"""K4's arithmetic -- the device helpers of csrc/resample.cu (``affine_sample``: what
scipy.ndimage.affine_transform computes per output sample for order 0 / 1; ``reduce_window``: the 13 reducers of
coarsen.py / constants.py:51-65 with numpy's semantics) compiled for the HOST (tests/hostmath) and driven by
the loop nest of ``k4_affine_generic`` -- against the reference goldens, numpy's reducers (the oracle) and scipy
itself, bit for bit, without a GPU."""

import numpy as np
import pytest
from scipy import ndimage

from oracle import resample as ores

from .helpers import assert_same, load_golden

nan = np.nan
AGGS = ("center", "count", "first", "last", "max", "mean", "median", "mode", "min", "prod", "std", "sum", "var")


@pytest.fixture(scope="module")
def k4(tmp_path_factory):
    from . import hostmath

    try:
        so = hostmath.build_k4(str(tmp_path_factory.mktemp("k4host")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise
    return lambda *a, **k: hostmath.affine(so, *a, **k)


def _as_library_dtype(want, src, agg):
    """numpy gives uint64 for unsigned sum / prod where the library's plane is int64 viewed as uint64 by the
    Python layer (affine.py); compare bit patterns."""
    return want.view(np.int64) if want.dtype == np.uint64 else want


@pytest.mark.parametrize("case", [str(c) for c in load_golden("coarsen.npz")["cases"]])
def test_reducers_against_the_reference_goldens(k4, case):
    """tests/golden/coarsen.npz: outputs of the reference's own reducers (coarsen.py) on seeded blocks."""
    z = load_golden("coarsen.npz")
    src = z[f"{case}/input"]
    f_j, f_i = (int(v) for v in z[f"{case}/factors"])
    for agg in AGGS:
        key = f"{case}/{agg}"
        if key not in z.files:
            continue
        want = _as_library_dtype(z[key], src, agg)
        got = k4(src, (src.shape[0] // f_j, src.shape[1] // f_i), agg=agg, factors=(f_j, f_i))
        assert_same(got, want, key)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16, np.uint16, np.int32])
@pytest.mark.parametrize("factors", [(2, 2), (4, 4), (8, 8), (3, 5), (1, 6), (16, 16)])
def test_reducers_against_numpy(k4, dtype, factors):
    f_j, f_i = factors
    rng = np.random.default_rng(f_j * 31 + f_i)
    h, w = 6 * f_j, 5 * f_i
    kind = np.dtype(dtype).kind
    if kind == "f":
        src = ((rng.random((h, w)) - 0.4) * 100).astype(dtype)
        src[rng.random((h, w)) < 0.08] = nan
        src[:f_j, :f_i] = nan  # one all-NaN window
        src[h - 1, w - 1] = np.inf
    else:
        src = rng.integers(0, 7 if kind == "u" else 9, (h, w)).astype(dtype) - (0 if kind == "u" else 4)
    for agg in AGGS:
        if agg == "mode" and kind == "f":
            continue  # the reference's mode is defined on integer classes (coarsen.py:114-155)
        want = _as_library_dtype(ores.coarsen(src, f_j, f_i, agg), src, agg)
        got = k4(src, (h // f_j, w // f_i), agg=agg, factors=factors)
        assert_same(got, want, f"{np.dtype(dtype).name} {agg} {factors}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16])
@pytest.mark.parametrize("order", [0, 1])
def test_affine_samples_are_scipys(k4, dtype, order):
    """Scales, offsets, the image border (cval without tolerance), the mirrored upper tap, NaN / inf neighbours
    with zero weight, and scipy's rounding into integer outputs."""
    rng = np.random.default_rng(order)
    h, w = 23, 31
    src = (rng.random((h, w)) * 200).astype(dtype)
    if np.dtype(dtype).kind == "f":
        src[5, 7] = nan
        src[11, 3] = np.inf
    cval = nan if np.dtype(dtype).kind == "f" else 7.0
    for (sj, si), (oj, oi), out_hw in (((1.0, 1.0), (0.0, 0.0), (23, 31)), ((2.0, 2.0), (0.5, 0.5), (11, 15)),
                                       ((0.5, 0.5), (-0.25, -0.25), (46, 62)), ((1.7, 0.6), (-1.3, 2.2), (16, 40)),
                                       ((1.0, 1.0), (3.0, -4.0), (23, 31)), ((0.37, 2.9), (21.9, -0.1), (9, 12))):
        want = ndimage.affine_transform(src, np.diag([sj, si]), offset=(oj, oi), order=order, output_shape=out_hw,
                                        mode="constant", cval=cval)
        got = k4(src, out_hw, scale_ji=(sj, si), offset_ji=(oj, oi), order=order, cval=cval)
        assert_same(got, want, f"{np.dtype(dtype).name} order {order} scale {(sj, si)} offset {(oj, oi)}")


def test_bilinear_downscale_of_config_c1_in_small(k4):
    """affine.py:277-313 for a 2x bilinear down-scaling on aligned grids (BASELINE config C1): order-1 samples
    at twice the target resolution, then the mean of 2 x 2 windows -- fused in one pass."""
    rng = np.random.default_rng(3)
    src = rng.random((64, 96)).astype(np.float32)
    src[10, 10] = nan
    matrix = ((2.0, 0.0, 0.0), (0.0, 2.0, 0.0))
    want = ores.resample_array(src, matrix, (32, 48), 1, "mean", False, nan)
    got = k4(src, (32, 48), scale_ji=(1.0, 1.0), offset_ji=(0.0, 0.0), order=1, cval=nan, agg="mean", factors=(2, 2))
    assert_same(got, want, "C1-like 2x bilinear downsample")

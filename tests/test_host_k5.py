"""K5's arithmetic -- the reducers of the streaming kernels on register windows (csrc/resample_fast.cu: the
bitonic sorting network, ``reduce_simple``, ``reduce_sort``; what ``k5_window_reduce`` runs for the aligned
2 / 4 / 8 block aggregations of BASELINE config C4) compiled for the HOST (tests/hostmath) -- against the
reference's coarsen goldens and numpy's reducers (the oracle), bit for bit, without a GPU."""

import numpy as np
import pytest

from oracle import resample as ores

from .helpers import assert_same, load_golden

nan = np.nan
AGGS = ("center", "count", "first", "last", "max", "mean", "median", "mode", "min", "prod", "std", "sum", "var")


@pytest.fixture(scope="module")
def k5(tmp_path_factory):
    from . import hostmath

    try:
        so = hostmath.build_k5(str(tmp_path_factory.mktemp("k5host")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise
    return lambda src, f, agg: hostmath.fast_reduce(so, src, f, agg)


def _bits(want):
    return want.view(np.int64) if want.dtype == np.uint64 else want  # unsigned sum / prod: same 64 bits


@pytest.mark.parametrize("case", ["f32_f2", "f32_f4", "f32_f8", "f64_f4", "u8_f4", "u8_f8", "i16_f2"])
def test_reference_goldens(k5, case):
    z = load_golden("coarsen.npz")
    src = z[f"{case}/input"]
    f = int(z[f"{case}/factors"][0])
    assert tuple(z[f"{case}/factors"]) == (f, f)
    for agg in AGGS:
        if f"{case}/{agg}" in z.files:
            assert_same(k5(src, f, agg), _bits(z[f"{case}/{agg}"]), f"{case}/{agg}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16, np.uint16, np.int32])
@pytest.mark.parametrize("f", [2, 4, 8])
def test_against_numpy(k5, dtype, f):
    rng = np.random.default_rng(f)
    h, w = 9 * f, 7 * f
    kind = np.dtype(dtype).kind
    if kind == "f":
        src = ((rng.random((h, w)) - 0.4) * 100).astype(dtype)
        src[rng.random((h, w)) < 0.1] = nan     # medians over 1 .. f*f - 1 valid values
        src[:f, :f] = nan                        # an all-NaN window
        src[f:2 * f, :f] = 3.25                  # a constant window
        src[h - 1, w - 1] = np.inf
        src[h - f, w - f] = -np.inf
        src[2 * f, 0], src[2 * f, 1] = 0.0, -0.0
    else:
        src = (rng.integers(0, 6, (h, w)) - (0 if kind == "u" else 3)).astype(dtype)  # few classes: mode ties
        src[:f, :f] = 5
    for agg in AGGS:
        if agg == "mode" and kind == "f":
            continue
        assert_same(k5(src, f, agg), _bits(ores.coarsen(src, f, f, agg)), f"{np.dtype(dtype).name} /{f} {agg}")


def test_class_raster_mode_ties_take_the_lowest_class(k5):
    """coarsen.py:138-155: the most frequent value, the lowest one among equally frequent values."""
    src = np.array([[3, 1, 2, 2], [1, 3, 0, 0], [7, 7, 9, 9], [7, 9, 9, 7]], dtype=np.uint8)
    assert k5(src, 2, "mode").tolist() == [[1, 0], [7, 9]]
    assert k5(src, 4, "mode").tolist() == [[7]]
    assert k5(src, 2, "mode").dtype == np.int64
    assert_same(k5(src, 2, "mode"), ores.coarsen(src, 2, 2, "mode"))

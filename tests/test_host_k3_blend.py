"""K3's per-pixel blend -- the helpers of csrc/reproject.cu (``k3_blend``, ``diff_as_f64``, ``cast_like_numpy``,
``window_index``) compiled for the HOST (tests/hostmath) -- against numpy evaluating the reference's own
expressions (reproject.py:301-328) in the arrays' dtypes, bit for bit, without a GPU.

What is pinned: numpy subtracts the taps in the ARRAY's dtype before the float64 promotion (unsigned and narrow
integers wrap), bilinear results are float64, triangular results and ``out_dtype=dtype`` bilinear results are
cast back the way numpy's ``astype`` does it on x86-64 (cvttsd2si: NaN and out-of-range values become the
"integer indefinite" pattern, then truncated to the destination width).

One documented platform dependence: float64 values outside [0, 2^32) cast to uint32.  C leaves that undefined;
numpy's result depends on its build (this image's AVX-512 build gives 0 above 2^32, the scalar path the low 32
bits of the 64-bit conversion, which is what the kernel does).  Such values are compared in range only."""

import platform
import warnings

import numpy as np
import pytest

nan = np.nan
DTYPES = [np.float32, np.float64, np.uint8, np.int8, np.uint16, np.int16, np.int32, np.uint32, np.int64]


@pytest.fixture(scope="module")
def blend(tmp_path_factory):
    from . import hostmath

    try:
        so = hostmath.build_k3_blend(str(tmp_path_factory.mktemp("k3host")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise
    return (lambda *a, **k: hostmath.k3_blend(so, *a, **k)), so


def _taps(dtype, n, rng, full_range):
    if np.dtype(dtype).kind == "f":
        taps = [((rng.random(n) - 0.5) * 1000).astype(dtype) for _ in range(4)]
        taps[0][5], taps[1][6], taps[2][7], taps[3][8] = nan, np.inf, -np.inf, nan
        return taps
    info = np.iinfo(dtype)
    lo, hi = (info.min, info.max) if full_range else (0 if info.min == 0 else -100, 100)
    if np.dtype(dtype).itemsize == 8:
        lo, hi = max(lo, -2**62), min(hi, 2**62)
    return [rng.integers(lo, hi, n, dtype=np.int64, endpoint=True).astype(dtype) for _ in range(4)]


def _numpy_blends(p00, p01, p10, p11, u, v):
    """reproject.py:301-328 as numpy evaluates it."""
    with np.errstate(over="ignore", invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        top = p00 + u * (p01 - p00)
        bottom = p10 + u * (p11 - p10)
        bilinear = top + v * (bottom - top)
        near = p00 + u * (p01 - p00) + v * (p10 - p00)
        far = p11 + (1.0 - u) * (p10 - p11) + (1.0 - v) * (p01 - p11)
        triangular_f64 = np.where(u + v < 1.0, near, far)
        return bilinear, triangular_f64


@pytest.mark.skipif(platform.machine() not in ("x86_64", "AMD64"), reason="numpy's float -> int casts are x86 semantics here")
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("full_range", [False, True])
def test_blend_equals_numpy_in_every_dtype(blend, dtype, full_range):
    fn, _ = blend
    rng = np.random.default_rng(np.dtype(dtype).num + 100 * full_range)
    n = 5000
    u, v = rng.random(n), rng.random(n)
    u[:20], v[:20] = 0.0, 0.0
    u[20:40] = 1.0
    v[40:60] = 1.0
    u[60:80], v[60:80] = 0.5, 0.5  # u + v == 1: the far triangle
    taps = _taps(dtype, n, rng, full_range)
    bilinear, triangular_f64 = _numpy_blends(*taps, u, v)
    assert bilinear.dtype == np.float64  # numpy promotes (reproject.py:325-327)
    got = fn(*taps, u, v, "bilinear", True)
    assert np.array_equal(got, bilinear, equal_nan=True), "bilinear, float64 out"
    with np.errstate(invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want_tri, want_bil = triangular_f64.astype(dtype), bilinear.astype(dtype)
    got_tri, got_bil = fn(*taps, u, v, "triangular", False), fn(*taps, u, v, "bilinear", False)
    kind = np.dtype(dtype).kind
    if dtype == np.uint32:  # see the module docstring
        ok_t = (triangular_f64 >= 0) & (triangular_f64 < 2.0**32)
        ok_b = (bilinear >= 0) & (bilinear < 2.0**32)
        assert ok_t.mean() > 0.3 and np.array_equal(got_tri[ok_t], want_tri[ok_t])
        assert np.array_equal(got_bil[ok_b], want_bil[ok_b])
    else:
        assert np.array_equal(got_tri, want_tri, equal_nan=kind == "f"), "triangular, cast back to the dtype"
        assert np.array_equal(got_bil, want_bil, equal_nan=kind == "f"), "bilinear with out_dtype = dtype"


def test_window_index_is_numpys_negative_indexing(blend):
    """reproject.py:284,295-298: indices into the tile's window wrap once from the end, like numpy's; what is
    still outside raises IndexError there and reads as "no data" here."""
    import ctypes

    _, so = blend
    lib = ctypes.CDLL(so)
    out = ctypes.c_long(0)
    n = 7
    arr = np.arange(n)
    for k in range(-2 * n - 1, 2 * n + 2):
        ok = lib.xrsh_window_index(ctypes.c_long(k), n, ctypes.byref(out))
        try:
            want = int(arr[k])
        except IndexError:
            want = None
        assert bool(ok) == (want is not None), k
        if want is not None:
            assert out.value == want, k

"""K2 as kernels -- k2_gather_staged / k2_gather_direct (csrc/gather.cu) and the two-method k2_gather_dual
(csrc/gather_dual.cu, the dominant kernel of the benchmark step) compiled UNCHANGED for the host
(tests/hostmath.build_k2) -- against the oracle's gather (rectify.py:579-734), bit for bit, without a GPU.

The shim writes out what the hardware does: a CTA is 256 host threads, `tma_load_2d` copies the tile's
64x48 source box into the staging buffer (zero outside the tensor) and completes an mbarrier phase,
`mbar_wait` polls the phase parity.  So the CPU suite exercises the kernels' own tile logic: the CTA-wide box
reduction, the 16-byte alignment of the box origin, the 4-stage ring over more bands than stages, launches
split at 24 bands, boxes that reach outside the resident window, tiles whose box exceeds the staging buffer
(global taps inside the same kernel), tiles without any source pixel, ragged image edges."""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import rectify as orect

from .helpers import assert_same, covering_grid_args, hand_made_ij, swath

nan = np.nan
FILLS = {np.float32: nan, np.float64: nan, np.uint8: 255, np.int16: -1}


@pytest.fixture(scope="module")
def k2_so(tmp_path_factory):
    from . import hostmath

    try:
        return hostmath.build_k2(str(tmp_path_factory.mktemp("k2host")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise


@pytest.fixture(scope="module")
def k1_so(tmp_path_factory):
    from . import hostmath

    try:
        return hostmath.build_k1(str(tmp_path_factory.mktemp("k1host_for_k2")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise


@pytest.fixture(scope="module")
def scene():
    w, h = 120, 90
    x, y = swath(w, h, theta=20.0, seed=3)
    size, xy_min = covering_grid_args(x, y, 0.0027)
    g = ogrid.regular_grid(size, xy_min, 0.0027, tile_size=64)
    return (w, h), orect.rectify_ij(x, y, g)   # target 203 x 126: ragged 32x32 tiles, corners without a source


def _source(dtype, n_bands, h, w, seed=0):
    rng = np.random.default_rng(seed)
    if np.issubdtype(dtype, np.floating):
        src = (rng.random((n_bands, h, w)) * 100).astype(dtype)
        src[0][rng.random((h, w)) < 0.02] = nan
        return src
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, (n_bands, h, w), endpoint=True).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16])
@pytest.mark.parametrize("method", ["nearest", "bilinear", "triangular"])
def test_staged_and_direct_kernels(k2_so, scene, dtype, method):
    from . import hostmath

    (w, h), ij = scene
    src = _source(dtype, 6, h, w)   # 6 bands through a ring of 4 stages
    want = orect.gather(src, ij, method, FILLS[dtype])
    out, (boxes, oob) = hostmath.k2_gather(k2_so, src, ij, method, FILLS[dtype], staged=True)
    assert_same(out, want, f"k2_gather_staged<{dtype.__name__}, {method}>")
    assert boxes > 0 and boxes % 6 == 0 and 0 < oob < boxes  # every staged tile pulled one box per band
    out, _ = hostmath.k2_gather(k2_so, src, ij, method, FILLS[dtype], staged=False)
    assert_same(out, want, f"k2_gather_direct<{dtype.__name__}, {method}>")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16])
@pytest.mark.parametrize("method", ["bilinear", "triangular"])
def test_two_method_kernel(k2_so, scene, dtype, method):
    from . import hostmath

    (w, h), ij = scene
    src = _source(dtype, 5, h, w, seed=1)
    fill_i, fill_n = (nan, nan) if np.issubdtype(dtype, np.floating) else (FILLS[dtype], 7)
    out_i, out_n, (boxes, _) = hostmath.k2_gather_dual(k2_so, src, ij, method, fill_i, fill_n)
    assert_same(out_i, orect.gather(src, ij, method, fill_i), f"k2_gather_dual<{dtype.__name__}, {method}>: interpolated")
    assert_same(out_n, orect.gather(src, ij, "nearest", fill_n), f"k2_gather_dual<{dtype.__name__}, {method}>: nearest")
    assert boxes > 0 and boxes % 5 == 0


@pytest.mark.parametrize("n_bands", [1, 4, 9, 24, 25])
def test_band_counts_around_the_ring_and_the_launch_split(k2_so, n_bands):
    from . import hostmath

    h, w, H, W = 70, 96, 64, 80
    ij = hand_made_ij(True, h, w, H, W)
    src = _source(np.float32, n_bands, h, w, seed=n_bands)
    out, (boxes, _) = hostmath.k2_gather(k2_so, src, ij, "bilinear", nan, staged=True)
    assert_same(out, orect.gather(src, ij, "bilinear", nan), f"{n_bands} bands, staged")
    assert boxes > 0 and boxes % n_bands == 0   # one box per band for every staged tile (the others tap global memory)
    out_i, out_n, _ = hostmath.k2_gather_dual(k2_so, src, ij, "bilinear", nan, nan)
    assert_same(out_i, orect.gather(src, ij, "bilinear", nan), f"{n_bands} bands, dual")
    assert_same(out_n, orect.gather(src, ij, "nearest", nan), f"{n_bands} bands, dual nearest")


@pytest.mark.parametrize("method", ["nearest", "bilinear", "triangular"])
def test_ties_edges_and_scattered_taps(k2_so, method):
    """hand_made_ij: exact half-pixel fractions, last row / column (clamped neighbours), zeros, NaN in one
    plane; smooth = every tile staged, scattered = every tile's box exceeds 64x48 and the taps come from
    global memory inside the staged kernels."""
    from . import hostmath

    h, w, H, W = 70, 96, 64, 80
    src = _source(np.int16, 3, h, w, seed=9)
    for smooth in (True, False):
        ij = hand_made_ij(smooth, h, w, H, W)
        out, (boxes, _) = hostmath.k2_gather(k2_so, src, ij, method, -1, staged=True)
        assert_same(out, orect.gather(src, ij, method, -1), f"{method}, smooth={smooth}")
        assert (boxes > 0) is smooth
        if method != "nearest":
            out_i, out_n, (boxes, _) = hostmath.k2_gather_dual(k2_so, src, ij, method, -1, -2)
            assert_same(out_i, orect.gather(src, ij, method, -1), f"dual {method}, smooth={smooth}")
            assert_same(out_n, orect.gather(src, ij, "nearest", -2), f"dual nearest, smooth={smooth}")
            assert (boxes > 0) is smooth


def test_resident_window_only(k2_so, scene):
    """The multi-GPU path hands the kernels a WINDOW of the source (pointer to its origin, its extent as the
    tensor's): boxes that reach beyond it are zero-filled by the copy and never used."""
    from . import hostmath

    (w, h), ij = scene
    rows = slice(0, 30)                        # a row band of the target: it sees one corner of the swath
    band_ij = ij[:, rows]
    fi, fj = band_ij
    i0, i1 = int(np.nanmin(fi)) // 32 * 32, min(w, int(np.nanmax(fi)) + 2)   # 32-column alignment as footprint_segments
    j0, j1 = int(np.nanmin(fj)), min(h, int(np.nanmax(fj)) + 2)
    assert i0 > 0 and (i1 - i0) * (j1 - j0) < 0.4 * w * h
    src = _source(np.float32, 5, h, w, seed=2)
    poisoned = np.full_like(src, 1e30)
    poisoned[:, j0:j1, i0:i1] = src[:, j0:j1, i0:i1]
    for method in ("nearest", "bilinear"):
        want = orect.gather(src, band_ij, method, nan)
        for staged in (True, False):
            out, _ = hostmath.k2_gather(k2_so, poisoned, band_ij, method, nan, staged=staged, window=(i0, j0, i1, j1))
            assert_same(out, want, f"{method}, staged={staged}, window")
    out_i, out_n, _ = hostmath.k2_gather_dual(k2_so, poisoned, band_ij, "bilinear", nan, nan, window=(i0, j0, i1, j1))
    assert_same(out_i, orect.gather(src, band_ij, "bilinear", nan), "dual, window")
    assert_same(out_n, orect.gather(src, band_ij, "nearest", nan), "dual nearest, window")


@pytest.mark.parametrize("tile,j_up,rows", [(48, False, None), ((40, 24), True, None), (64, False, (32, 96))])
def test_fused_resolve_and_gather(k2_so, k1_so, tile, j_up, rows):
    """xrs_rectify_gather end to end on the host: K1's claim stage (host build of rectify_ij.cu), then the staged
    gather in FUSED mode, which resolves the claim words in registers (rectify_common.cuh resolve_pixel) instead of
    reading an ij image -- equal to the oracle's rectification of the same bands."""
    from . import hostmath

    w, h = 110, 100
    x, y = swath(w, h, theta=-25.0, seed=31)
    x[60, 20:30] = nan
    size, xy_min = covering_grid_args(x, y, 0.0027)
    g = ogrid.regular_grid(size, xy_min, 0.0027, tile_size=tile, is_j_axis_up=j_up)
    windows = orect.source_windows(x, y, g)
    r0, r1 = rows if rows is not None else (0, g.height)
    _, claims, _ = hostmath.k1(k1_so, x, y, windows, g, rows=rows)
    src = _source(np.float32, 5, h, w, seed=6)
    for method in ("nearest", "bilinear", "triangular"):
        want, _ = orect.rectify(x, y, src, g, method, nan)
        got = hostmath.k2_gather_fused(k2_so, src, x, y, windows, claims, g, method, nan, rows=rows)
        assert_same(got, want[:, r0:r1], f"fused {method}, rows {rows}")

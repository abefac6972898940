"""K1 as a whole -- the text of csrc/rectify_ij.cu (k1_init_claims, k1_scatter, k1_scatter_slow, k1_resolve)
compiled UNCHANGED for the host (tests/hostmath.build_k1: threadIdx / blockIdx as thread-local variables, a
warp as 32 host threads in lock step, shuffles and ballots through a shared array, GCC atomics) -- against
the oracle's ij image, bit for bit, without a GPU.

What this pins on the CPU: the pixel-space fast path of the scatter (stepped edge functions, sign-bit
decisions, margins, centre-tight candidate boxes, magic division, single-tile test), its hand-over of
undecidable quads to the generic kernel (the reference's arithmetic, rectify.py:458-576), the first-writer
rule as an atomicMin over quad indices, row-band calls with and without a caller-supplied footprint, and the
resolve step.  The cases are the GPU suite's adversarial geometries (tests/test_rectify_gpu.py) at sizes the
warp emulation finishes in seconds."""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import rectify as orect

from .helpers import assert_same, covering_grid_args, quad_footprints_np, swath
from .test_host_resolve import NOCLAIM, claims_and_ij

nan = np.nan


@pytest.fixture(scope="module")
def k1_so(tmp_path_factory):
    from . import hostmath

    try:
        return hostmath.build_k1(str(tmp_path_factory.mktemp("k1host")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise


def _check(k1_so, x, y, g, uv_delta=1e-3):
    from . import hostmath

    windows = orect.source_windows(x, y, g)
    want = orect.rectify_ij(x, y, g, uv_delta=uv_delta, windows=windows)
    ij, claims, queued = hostmath.k1(k1_so, x, y, windows, g, uv_delta=uv_delta)
    assert_same(ij, want, "K1 (host build of rectify_ij.cu) vs the oracle")
    assert np.array_equal(claims != NOCLAIM, ~np.isnan(want[0]))
    return windows, want, claims, queued


@pytest.mark.parametrize("shape,theta,res_factor,tile,j_up", [
    ((46, 38), 12.0, 1.0, 16, False),        # several small reference tiles, quads ~1 px
    ((40, 33), -35.0, 0.6, (23, 9), False),  # finer target (quads span 2-3 px), ragged non-square tiles
    ((52, 30), 77.0, 1.7, None, False),      # coarser target (several quads per pixel), one tile
    ((37, 41), 5.0, 1.0, 12, True),          # j axis up
])
def test_seeded_swaths_claims_and_ij(k1_so, shape, theta, res_factor, tile, j_up):
    w, h = shape
    x, y = swath(w, h, theta=theta, seed=w * h)
    res = 0.0027 * res_factor
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile, is_j_axis_up=j_up)
    windows, _, claims, queued = _check(k1_so, x, y, g)
    # the claim words themselves: smallest accepting quad index and its triangle, as the reference's
    # sequential first-writer walk leaves them
    walk_claims, _ = claims_and_ij(x, y, g, windows)
    assert np.array_equal(claims, walk_claims)
    assert queued < 0.3 * (w - 1) * (h - 1)  # the fast path decides most quads itself


@pytest.mark.parametrize("tile", [None, 16, (24, 10)])
@pytest.mark.parametrize("ratio", [1.0, 2.0, 0.5, 3.0])
def test_axis_aligned_source_on_pixel_centres(k1_so, tile, ratio):
    """Vertices exactly on target pixel centres / corners: every triangle edge passes through pixel centres
    (u, v exactly 0 or 1; the uv tolerance decides them, rectify.py:558-573)."""
    res = 0.25
    w, h = 41, 33
    x = np.broadcast_to(10.0 + ratio * res * (np.arange(w) + 0.5), (h, w)).copy()
    y = np.broadcast_to((50.0 - ratio * res * (np.arange(h) + 0.5))[:, None], (h, w)).copy()
    size = (int(w * ratio) + 4, int(h * ratio) + 4)
    g = ogrid.regular_grid(size, (10.0 - 2 * res, 50.0 - (size[1] - 2) * res), res, tile_size=tile)
    _check(k1_so, x, y, g)


def test_holes_duplicates_and_folds(k1_so):
    """NaN holes, an infinite vertex, duplicated rows / columns (zero-area quads) and a fold."""
    x, y = swath(120, 100, theta=25.0, seed=11)
    x[20:24, 30:50] = nan
    y[20:24, 30:50] = nan
    x[50, :] = x[49, :]
    y[50, :] = y[49, :]
    x[:, 70] = x[:, 69]
    y[:, 70] = y[:, 69]
    x[75:90] = x[75:90][::-1].copy()
    y[75:90] = y[75:90][::-1].copy()
    x[5, 5] = np.inf
    res = 0.0027
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=64)
    _check(k1_so, x, y, g)


def test_projected_coordinate_magnitudes(k1_so):
    """UTM-like metres (5e5 / 5e6 with 300 m pixels): the rounding of the reference's pixel centres is an ulp
    of the CRS coordinate, far above that of the kernel's own pixel coordinates."""
    lon, lat = swath(110, 90, theta=-8.0, seed=13)
    x = 500000.0 + (lon - 10.0) * 78000.0
    y = 5000000.0 + (lat - 45.0) * 111000.0
    res = 300.0
    size, xy_min = covering_grid_args(x, y, res)
    for tile in (None, 50):
        _check(k1_so, x, y, ogrid.regular_grid(size, xy_min, res, tile_size=tile))


def test_huge_quads_and_loose_tolerance_take_the_generic_path(k1_so):
    x, y = swath(12, 10, theta=33.0, seed=17)
    res = 0.0027 / 90.0  # quads of ~90 target pixels (> K1_MAX_EXTENT)
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=256)
    _, _, _, queued = _check(k1_so, x, y, g)
    assert queued > 0.5 * 11 * 9
    x, y = swath(40, 30, theta=-20.0, seed=19)
    size, xy_min = covering_grid_args(x, y, 0.0027)
    g = ogrid.regular_grid(size, xy_min, 0.0027, tile_size=32)
    _, _, _, queued = _check(k1_so, x, y, g, uv_delta=0.05)  # 0.05 * 66 px > 0.25: no fast path at all
    assert queued == 39 * 29


def test_row_bands_with_and_without_a_footprint(k1_so):
    """A row-band call (multi-GPU path) gives the rows of the whole image; the caller's quad footprint
    (xrs_band_quad_footprints, restated in tests/helpers.py) only removes quads that cannot reach the band."""
    from . import hostmath

    x, y = swath(80, 110, theta=30.0, seed=23)
    res = 0.0027
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=48)
    windows = orect.source_windows(x, y, g)
    want = orect.rectify_ij(x, y, g, windows=windows)
    edges = [0, 32, 72, g.height]
    fp = quad_footprints_np(x, y, g, edges, group=32)
    for b in range(3):
        rows = (edges[b], edges[b + 1])
        ij, _, _ = hostmath.k1(k1_so, x, y, windows, g, rows=rows)
        assert_same(ij, want[:, rows[0]:rows[1]], f"rows {rows}")
        ij, _, _ = hostmath.k1(k1_so, x, y, windows, g, rows=rows, fp_cols=fp[b])
        assert_same(ij, want[:, rows[0]:rows[1]], f"rows {rows} with the band's quad footprint")

"""K0 -- the kernels of csrc/rectify.cu (k0_tile_windows with the tile table in shared memory, in global
memory and in min-form; k0_init_table, k0_finalize, k0_fold_minform, k0_finalize_minform) compiled UNCHANGED
for the host (tests/hostmath.build_k0: a thread block = 256 host threads, __syncthreads a barrier, dynamic
shared memory a static array, __match_any_sync / __reduce_*_sync as warp and peer-group collectives) --
against the reference's compute_ij_bboxes (gridmapping/bboxes.py:28-106): its golden outputs, the
expectations of tests/gridmapping/test_bboxes.py, and the oracle on seeded swaths.  The tile axis tables
come from the product's own host code (rectify._separable_axes / _xy_border)."""

import numpy as np
import pytest

import xcube_resampling_b200 as xrs
from oracle import grid as ogrid
from oracle import rectify as orect
from xcube_resampling_b200.rectify import _separable_axes, _xy_border

from .helpers import assert_same, covering_grid_args, load_golden, swath

FORMS = [(0, 0), (1, 0), (2, 16), (3, 11)]  # (kernel form, slab rows): see hostmath.K0_EXPORT


@pytest.fixture(scope="module")
def k0_so(tmp_path_factory):
    from . import hostmath

    try:
        return hostmath.build_k0(str(tmp_path_factory.mktemp("k0host")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise


def _k0(k0_so, x, y, boxes, xy_border, ij_border, form=0, slab=0):
    from . import hostmath

    x_lo, x_hi, y_lo, y_hi, _, _ = _separable_axes(np.asarray(boxes, dtype=np.float64), xy_border)
    return hostmath.k0(k0_so, x, y, x_lo, x_hi, y_lo, y_hi, ij_border, form, slab)


@pytest.mark.parametrize("form,slab", [(0, 0), (2, 6)])  # (all four forms: the two tests below)
def test_reference_goldens(k0_so, form, slab):
    z = load_golden("ij_bboxes.npz")  # outputs of the reference's own numba kernel (tests/golden/make_golden.py)
    for k in range(int(z["n_cases"])):
        border, ij_border = z[f"case{k}/params"]
        # the golden boxes are independent random rectangles; K0 takes separable tile grids (what a regular target
        # grid has), so every box goes through as a grid of one tile
        for box, want in zip(z[f"case{k}/boxes"], z[f"case{k}/result"]):
            got = _k0(k0_so, z["x"], z["y"], box[None], float(border), int(ij_border), form, slab)
            assert_same(got, want[None], f"case{k}, box {box}, form {form}")


def test_reference_unit_test_expectations(k0_so):
    # tests/gridmapping/test_bboxes.py:40-139
    lon, lat = np.meshgrid(np.linspace(10.0, 20.0, 11), np.linspace(50.0, 60.0, 11))
    assert _k0(k0_so, lon, lat, [[10.0, 50.0, 20.0, 60.0]], 0.0, 0).tolist() == [[0, 0, 11, 11]]
    tiles = [[10.0, 50.0, 15.0, 55.0], [15.0, 50.0, 20.0, 55.0], [10.0, 55.0, 15.0, 60.0], [15.0, 55.0, 20.0, 60.0]]
    for form, slab in [(0, 0), (1, 0), (2, 5), (3, 4)]:
        assert _k0(k0_so, lon, lat, tiles, 0.0, 0, form, slab).tolist() == \
            [[0, 0, 6, 6], [5, 0, 11, 6], [0, 5, 6, 11], [5, 5, 11, 11]]
        assert _k0(k0_so, lon, lat, (np.array(tiles) + 11.0).tolist(), 0.0, 0, form, slab).tolist() == [[-1] * 4] * 4
    box = [[12.4, 51.6, 12.6, 51.7]]
    assert _k0(k0_so, lon, lat, box, 0.0, 0).tolist() == [[-1, -1, -1, -1]]
    assert _k0(k0_so, lon, lat, box, 0.5, 0).tolist() == [[2, 2, 4, 3]]
    assert _k0(k0_so, lon, lat, box, 1.0, 0).tolist() == [[2, 1, 4, 3]]
    assert _k0(k0_so, lon, lat, box, 2.0, 0).tolist() == [[1, 0, 5, 4]]
    assert _k0(k0_so, lon, lat, box, 2.0, 2).tolist() == [[0, 0, 7, 6]]
    lon_nan = lon.copy()
    lon_nan[3:9, 2:10] = np.nan  # NaN coordinates match no tile (bboxes.py:60-69: every comparison is false)
    assert_same(_k0(k0_so, lon_nan, lat, tiles, 0.0, 0), orect.ij_bboxes(lon_nan, lat, tiles, 0.0, 0), "NaN block")


@pytest.mark.parametrize("shape,theta,res_factor,tile,j_up", [
    ((46, 38), 12.0, 1.0, 16, False),
    ((300, 70), -35.0, 0.6, (23, 9), False),  # two column chunks of 256 threads, 30 x 43 tiles
    ((52, 30), 77.0, 1.7, None, False),       # one tile
    ((37, 41), 5.0, 1.0, 12, True),           # j axis up: the y axis tables ascend with the tile row
    ((530, 40), 20.0, 1.0, 64, False),        # three column chunks, the last one ragged
])
@pytest.mark.parametrize("form,slab", FORMS)
def test_seeded_swaths_through_the_products_axis_tables(k0_so, shape, theta, res_factor, tile, j_up, form, slab):
    from . import hostmath

    w, h = shape
    x, y = swath(w, h, theta=theta, seed=w * h)
    x[h // 3, w // 4:w // 2] = np.nan
    res = 0.0027 * res_factor
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile, is_j_axis_up=j_up)
    gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=tile, is_j_axis_up=j_up)
    x_lo, x_hi, y_lo, y_hi, ntx, nty = _separable_axes(gm.xy_bboxes, _xy_border(gm))
    assert (nty, ntx) == tuple(g.n_tiles)
    got = hostmath.k0(k0_so, x, y, x_lo, x_hi, y_lo, y_hi, 1, form, slab)
    assert_same(got, orect.source_windows(x, y, g), f"K0 form {form} (host build of rectify.cu) vs the oracle")

"""K2's per-pixel arithmetic -- the text of csrc/gather_common.cuh (make_taps / interp_value: tap selection,
clamping at the image edge, nearest ties, bilinear and triangular blends in float64, the final C cast) compiled
for the HOST (tests/hostmath) -- against the reference goldens and the oracle, bit for bit, without a GPU."""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import rectify as orect

from .helpers import assert_same, covering_grid_args, hand_made_ij, load_golden, swath

nan = np.nan


@pytest.fixture(scope="module")
def gather_so(tmp_path_factory):
    from . import hostmath

    try:
        return hostmath.build_gather(str(tmp_path_factory.mktemp("gatherhost")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise


def _golden_cases():
    return [str(c) for c in load_golden("rectify.npz")["cases"]]


@pytest.mark.parametrize("case", _golden_cases())
def test_reference_goldens(gather_so, case):
    """Outputs of the reference's own numba kernels (tests/golden/rectify.npz)."""
    from . import hostmath

    z = load_golden("rectify.npz")
    ij = z[f"{case}/ij"]
    for vname, fill in (("f32", nan), ("u8", 255), ("i16", -1), ("f64", nan)):
        src = z[f"{case}/src_{vname}"]
        for method in ("nearest", "bilinear", "triangular"):
            got = hostmath.gather(gather_so, src, ij, method, fill)
            assert_same(got, z[f"{case}/out_{vname}_{method}"], f"{case} {vname}/{method}")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16, np.uint16, np.int32, np.int64])
@pytest.mark.parametrize("smooth", [True, False])
def test_ties_edges_and_holes_against_the_oracle(gather_so, dtype, smooth):
    from . import hostmath

    h, w = 40, 70
    ij = hand_made_ij(smooth, h, w, 64, 96)
    data = (np.random.default_rng(6).random((3, h, w)) * 200).astype(dtype)
    if np.dtype(dtype).kind == "f":
        data[0, 3:6, 10:30] = nan
        data[1, 20, 20] = np.inf
    fill = nan if np.dtype(dtype).kind == "f" else 9
    for method in ("nearest", "bilinear", "triangular"):
        assert_same(hostmath.gather(gather_so, data, ij, method, fill), orect.gather(data, ij, method, fill),
                    f"{np.dtype(dtype).name} {method}")


def test_seeded_swath_against_the_oracle(gather_so):
    from . import hostmath

    x, y = swath(300, 240, theta=-22.0, seed=21)
    size, xy_min = covering_grid_args(x, y, 0.0027)
    g = ogrid.regular_grid(size, xy_min, 0.0027, tile_size=96)
    ij = orect.rectify_ij(x, y, g)
    data = np.random.default_rng(2).random((4, 240, 300)).astype(np.float32)
    for method in ("nearest", "bilinear", "triangular"):
        assert_same(hostmath.gather(gather_so, data, ij, method, nan), orect.gather(data, ij, method, nan), method)

"""Pin the CPU oracle: it must reproduce (a) the outputs of the reference's own
kernels stored in tests/golden/*.npz and (b) the expected arrays of the
reference's unit tests (file:line cited per case).  CPU only."""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import rectify as orect

from .helpers import assert_same, grid_from_golden, load_golden

nan = np.nan


# ---------------------------------------------------------------------------
# (a) reference-kernel outputs
# ---------------------------------------------------------------------------
def _rectify_cases():
    z = load_golden("rectify.npz")
    return z, [str(c) for c in z["cases"]]


@pytest.mark.parametrize("case", _rectify_cases()[1])
def test_rectify_windows_and_ij_match_reference_kernels(case):
    z, _ = _rectify_cases()
    g = grid_from_golden(z[f"{case}/grid"])
    x, y = z[f"{case}/x"], z[f"{case}/y"]
    assert_same(orect.source_windows(x, y, g), z[f"{case}/windows"], "windows")
    assert_same(orect.rectify_ij(x, y, g), z[f"{case}/ij"], "ij")


@pytest.mark.parametrize("case", _rectify_cases()[1])
@pytest.mark.parametrize("method", ["nearest", "bilinear", "triangular"])
def test_rectify_gather_matches_reference_kernels(case, method):
    z, _ = _rectify_cases()
    ij = z[f"{case}/ij"]
    for vname, fill in (("f32", nan), ("u8", 255), ("i16", -1), ("f64", nan)):
        out = orect.gather(z[f"{case}/src_{vname}"], ij, method, fill)
        assert_same(out, z[f"{case}/out_{vname}_{method}"], f"{vname}/{method}")


@pytest.mark.parametrize("case", ["f32_tiled32", "f32_tiled_17x40_jup"])
def test_rectify_float32_coordinates(case):
    """The reference kernels were run on float32 coordinate images (rectify.py:480-501 keeps the
    coordinates' dtype for the vertex arrays); the C ABI takes float64 only, so callers up-cast --
    which must give the same windows and the same ij image, bit for bit."""
    z = load_golden("rectify_f32coords.npz")
    g = grid_from_golden(z[f"{case}/grid"])
    x, y = z[f"{case}/x"], z[f"{case}/y"]
    assert x.dtype == np.float32 and y.dtype == np.float32
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    assert_same(orect.source_windows(x64, y64, g), z[f"{case}/windows"], "windows")
    assert_same(orect.rectify_ij(x64, y64, g), z[f"{case}/ij"], "ij")


def test_ij_bboxes_match_reference_kernel():
    z = load_golden("ij_bboxes.npz")
    for k in range(int(z["n_cases"])):
        border, ij_border = z[f"case{k}/params"]
        got = orect.ij_bboxes(z["x"], z["y"], z[f"case{k}/boxes"], border, int(ij_border))
        assert_same(got, z[f"case{k}/result"], f"case{k}")


# ---------------------------------------------------------------------------
# (b) expectations of the reference's unit tests
# ---------------------------------------------------------------------------
LON_2X2 = np.array([[1.0, 6.0], [0.0, 2.0]])
LAT_2X2 = np.array([[56.0, 53.0], [52.0, 50.0]])
RAD_2X2 = np.array([[1.0, 2.0], [3.0, 4.0]])

# tests/test_rectify.py:529-547 (expected_rad_13x13)
EXPECTED_RAD_13X13 = np.array([
    [nan, nan, 1.0, nan, nan, nan, nan, nan, nan, nan, nan, nan, nan],
    [nan, nan, 1.0, 1.0, nan, nan, nan, nan, nan, nan, nan, nan, nan],
    [nan, nan, 1.0, 1.0, 1.0, 1.0, nan, nan, nan, nan, nan, nan, nan],
    [nan, nan, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, nan, nan, nan, nan, nan],
    [nan, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 2.0, 2.0, nan, nan, nan, nan],
    [nan, 3.0, 3.0, 1.0, 1.0, 1.0, 1.0, 2.0, 2.0, 2.0, 2.0, nan, nan],
    [nan, 3.0, 3.0, 3.0, 3.0, 1.0, 1.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0],
    [nan, 3.0, 3.0, 3.0, 3.0, 3.0, 1.0, 2.0, 2.0, 2.0, 2.0, nan, nan],
    [3.0, 3.0, 3.0, 3.0, 3.0, 4.0, 4.0, 2.0, 2.0, 2.0, nan, nan, nan],
    [nan, 3.0, 3.0, 3.0, 4.0, 4.0, 4.0, 4.0, 2.0, nan, nan, nan, nan],
    [nan, nan, 3.0, 4.0, 4.0, 4.0, 4.0, nan, nan, nan, nan, nan, nan],
    [nan, nan, nan, 4.0, 4.0, 4.0, nan, nan, nan, nan, nan, nan, nan],
    [nan, nan, nan, nan, 4.0, nan, nan, nan, nan, nan, nan, nan, nan],
])


def _rect(lon, lat, rad, size, xy_min, res, method="nearest", **kw):
    g = ogrid.regular_grid(size, xy_min, res, **kw)
    out, _ = orect.rectify(lon, lat, rad, g, method, nan)
    return out


def test_ref_rectify_2x2_to_default():
    # tests/test_rectify.py:42-61
    out = _rect(LON_2X2, LAT_2X2, RAD_2X2, (4, 4), (-1, 49), 2)
    np.testing.assert_almost_equal(out, np.array([
        [nan, nan, nan, nan], [nan, 1.0, 2.0, nan], [3.0, 3.0, 2.0, nan], [nan, 4.0, nan, nan]]))


@pytest.mark.parametrize("tile_size", [None, 7, 5, (3, 13), (13, 3)])
@pytest.mark.parametrize("j_up", [False, True])
def test_ref_rectify_2x2_to_13x13(tile_size, j_up):
    # tests/test_rectify.py:261-387
    out = _rect(LON_2X2, LAT_2X2, RAD_2X2, (13, 13), (-0.25, 49.75), 0.5, tile_size=tile_size, is_j_axis_up=j_up)
    np.testing.assert_almost_equal(out, EXPECTED_RAD_13X13[::-1] if j_up else EXPECTED_RAD_13X13)


RAD_7X7_SRC = RAD_2X2 + np.array([[0.0, 0.0], [0.0, 1.0]])


def test_ref_rectify_2x2_to_7x7_nearest():
    # tests/test_rectify.py:112-144
    out = _rect(LON_2X2, LAT_2X2, RAD_7X7_SRC, (7, 7), (-0.5, 49.5), 1.0)
    np.testing.assert_almost_equal(out, np.array([
        [nan, 1.0, nan, nan, nan, nan, nan],
        [nan, 1.0, 1.0, nan, nan, nan, nan],
        [nan, 1.0, 1.0, 1.0, 2.0, nan, nan],
        [nan, 3.0, 3.0, 1.0, 2.0, 2.0, 2.0],
        [3.0, 3.0, 3.0, 5.0, 2.0, nan, nan],
        [nan, 3.0, 5.0, 5.0, nan, nan, nan],
        [nan, nan, 5.0, nan, nan, nan, nan]]))


def test_ref_rectify_2x2_to_7x7_triangular():
    # tests/test_rectify.py:146-181
    out = _rect(LON_2X2, LAT_2X2, RAD_7X7_SRC, (7, 7), (-0.5, 49.5), 1.0, method="triangular")
    np.testing.assert_almost_equal(out, np.array([
        [nan, 1.000, nan, nan, nan, nan, nan],
        [nan, 1.478, 1.391, nan, nan, nan, nan],
        [nan, 1.957, 1.870, 1.784, 1.697, nan, nan],
        [nan, 2.435, 2.348, 2.261, 2.174, 2.087, 2.000],
        [3.000, 3.000, 3.000, 3.000, 3.000, nan, nan],
        [nan, 4.000, 4.000, 4.000, nan, nan, nan],
        [nan, nan, 5.000, nan, nan, nan, nan]]), decimal=3)


def test_ref_rectify_2x2_to_7x7_bilinear():
    # tests/test_rectify.py:183-218
    out = _rect(LON_2X2, LAT_2X2, RAD_7X7_SRC, (7, 7), (-0.5, 49.5), 1.0, method="bilinear")
    np.testing.assert_almost_equal(out, np.array([
        [nan, 1.000, nan, nan, nan, nan, nan],
        [nan, 1.488, 1.410, nan, nan, nan, nan],
        [nan, 1.994, 1.949, 1.858, 1.722, nan, nan],
        [nan, 2.520, 2.506, 2.448, 2.344, 2.195, 2.000],
        [3.000, 3.112, 3.163, 3.153, 3.082, nan, nan],
        [nan, 4.000, 4.041, 4.020, nan, nan, nan],
        [nan, nan, 5.000, nan, nan, nan, nan]]), decimal=3)


def test_ref_rectify_2x2_to_7x7_subset():
    # tests/test_rectify.py:230-259
    out = _rect(LON_2X2, LAT_2X2, RAD_2X2, (7, 7), (1.5, 50.5), 1.0)
    np.testing.assert_almost_equal(out, np.array([
        [nan, nan, nan, nan, nan, nan, nan],
        [nan, nan, nan, nan, nan, nan, nan],
        [1.0, nan, nan, nan, nan, nan, nan],
        [1.0, 1.0, 2.0, nan, nan, nan, nan],
        [3.0, 1.0, 2.0, 2.0, 2.0, nan, nan],
        [3.0, 4.0, 2.0, nan, nan, nan, nan],
        [4.0, 4.0, nan, nan, nan, nan, nan]]))


@pytest.mark.parametrize("xy_min", [(10.0, 50.0), (-10.0, 50.0), (0.0, 58.0), (0.0, 42.0)])
def test_ref_rectify_no_overlap_is_all_nan(xy_min):
    # tests/test_rectify.py:426-459
    out = _rect(LON_2X2, LAT_2X2, RAD_2X2, (13, 13), xy_min, 0.5)
    assert np.isnan(out).all()


def test_ref_rectify_invalid_method():
    # tests/test_rectify.py:220-228
    g = ogrid.regular_grid((7, 7), (-0.5, 49.5), 1.0)
    with pytest.raises(NotImplementedError):
        orect.rectify(LON_2X2, LAT_2X2, RAD_2X2, g, "cubic", nan)


def test_ref_compute_ij_bboxes():
    # tests/gridmapping/test_bboxes.py:40-139
    lon, lat = np.meshgrid(np.linspace(10.0, 20.0, 11), np.linspace(50.0, 60.0, 11))
    assert orect.ij_bboxes(lon, lat, [[10.0, 50.0, 20.0, 60.0]]).tolist() == [[0, 0, 11, 11]]
    tiles = [[10.0, 50.0, 15.0, 55.0], [15.0, 50.0, 20.0, 55.0], [10.0, 55.0, 15.0, 60.0], [15.0, 55.0, 20.0, 60.0]]
    assert orect.ij_bboxes(lon, lat, tiles).tolist() == [[0, 0, 6, 6], [5, 0, 11, 6], [0, 5, 6, 11], [5, 5, 11, 11]]
    far = (np.array(tiles) + 11.0).tolist()
    assert orect.ij_bboxes(lon, lat, far).tolist() == [[-1] * 4] * 4
    box = [[12.4, 51.6, 12.6, 51.7]]
    assert orect.ij_bboxes(lon, lat, box, 0.0, 0).tolist() == [[-1, -1, -1, -1]]
    assert orect.ij_bboxes(lon, lat, box, 0.5, 0).tolist() == [[2, 2, 4, 3]]
    assert orect.ij_bboxes(lon, lat, box, 1.0, 0).tolist() == [[2, 1, 4, 3]]
    assert orect.ij_bboxes(lon, lat, box, 2.0, 0).tolist() == [[1, 0, 5, 4]]
    assert orect.ij_bboxes(lon, lat, box, 2.0, 2).tolist() == [[0, 0, 7, 6]]


# ---------------------------------------------------------------------------
# affine / coarsen
# ---------------------------------------------------------------------------
from oracle import resample as ores  # noqa: E402


def _coarsen_cases():
    z = load_golden("coarsen.npz")
    return z, [str(c) for c in z["cases"]]


@pytest.mark.parametrize("case", _coarsen_cases()[1])
def test_coarsen_reducers_match_reference(case):
    # outputs of the reference's AGG_METHODS table (constants.py:51-65, coarsen.py:50-155)
    z, _ = _coarsen_cases()
    a = z[f"{case}/input"]
    f_j, f_i = (int(v) for v in z[f"{case}/factors"])
    for agg in ores.AGG_NAMES:
        key = f"{case}/{agg}"
        if key not in z:
            continue
        got = np.asarray(ores.coarsen(a, f_j, f_i, agg))
        ref = z[key]
        assert got.dtype == ref.dtype and np.array_equal(got, ref, equal_nan=got.dtype.kind == "f"), f"{case}/{agg}"


def test_ref_coarsen_unit_test_values():
    # tests/test_coarsen.py:35-61
    arr_float = np.array([[1.0, 2.0], [3.0, 4.0]])
    arr_int = np.array([[1, 2], [3, 4]])
    arr_mode = np.array([[1, 2, 2], [3, 2, 2]])
    assert ores.coarsen(arr_float, 2, 2, "first") == 1.0
    assert ores.coarsen(arr_float, 2, 2, "last") == 4.0
    assert ores.coarsen(arr_float, 2, 2, "center") == 4.0
    assert ores.coarsen(arr_float, 2, 2, "mean") == 2.5
    assert ores.coarsen(arr_int, 2, 2, "mean") == 2
    assert ores.coarsen(arr_float, 2, 2, "median") == 2.5
    np.testing.assert_almost_equal(ores.coarsen(arr_float, 2, 2, "std"), np.std(arr_float))
    assert ores.coarsen(arr_int, 2, 2, "sum") == 10
    np.testing.assert_almost_equal(ores.coarsen(arr_float, 2, 2, "var"), np.var(arr_float))
    assert ores.coarsen(arr_mode, 2, 3, "mode") == 2


REFL_8X6 = np.array([[0, 1, 0, 2, 0, 3, 0, 4], [2, 0, 3, 0, 4, 0, 1, 0], [0, 4, 0, nan, 0, 2, 0, 3],
                     [1, 0, 2, 0, 3, 0, 4, 0], [0, 3, 0, 4, 0, 1, 0, 2], [4, 0, 1, 0, 2, 0, 3, 0]], dtype=np.float64)


def _affine(size, xy_min, res, **kw):
    src = ogrid.regular_grid((8, 6), (50, 10), 0.1)  # what GridMapping.from_dataset derives for the 8x6 sample
    tgt = ogrid.regular_grid(size, xy_min, res)
    return ores.affine_transform(REFL_8X6, src, tgt, interp=1, **kw)


def test_ref_affine_subset_and_recover_nans():
    # tests/test_affine.py:46-140
    np.testing.assert_almost_equal(_affine((3, 3), (50.0, 10.0), 0.1), [[1, 0, 2], [0, 3, 0], [4, 0, 1]])
    np.testing.assert_almost_equal(_affine((3, 3), (50.1, 10.1), 0.1), [[4, nan, nan], [0, 2, 0], [3, 0, 4]])
    np.testing.assert_almost_equal(_affine((3, 3), (50.05, 10.05), 0.1),
                                   [[1.25, 1.5, nan], [1.0, 1.25, 1.5], [1.75, 1.0, 1.25]])
    np.testing.assert_almost_equal(_affine((3, 3), (50.05, 10.05), 0.1, recover_nan=True),
                                   [[1.25, 1.5, 0.6666667], [1.0, 1.25, 1.5], [1.75, 1.0, 1.25]])


def test_ref_affine_downscale_upscale_shift():
    # tests/test_affine.py:295-478 (incl. the zero-weight NaN contamination value 1.0 at [3, 1])
    np.testing.assert_almost_equal(_affine((8, 6), (50, 10), 0.2), [
        [nan] * 8, [nan] * 8, [nan] * 8, [0.75, 1.0, 1.75, 1.25, nan, nan, nan, nan],
        [1.25, 1.0, 1.25, 1.75, nan, nan, nan, nan], [1.75, 1.25, 0.75, 1.25, nan, nan, nan, nan]])
    np.testing.assert_almost_equal(_affine((8, 6), (49.8, 9.8), 0.2), [
        [nan] * 8, [nan] * 8, [nan, 0.75, 1.0, 1.75, 1.25, nan, nan, nan], [nan, 1.25, 1.0, 1.25, 1.75, nan, nan, nan],
        [nan, 1.75, 1.25, 0.75, 1.25, nan, nan, nan], [nan] * 8])
    np.testing.assert_almost_equal(_affine((8, 6), (50, 10), 0.05), [
        [1.0, 0.5, 0.0, 1.0, 2.0, 1.0, 0.0, 1.5], [0.5, 1.0, 1.5, 1.25, 1.0, 1.5, 2.0, 1.75],
        [0.0, 1.5, 3.0, 1.5, 0.0, 2.0, 4.0, 2.0], [2.0, 1.75, 1.5, 1.0, 0.5, 1.25, 2.0, 1.5],
        [4.0, 2.0, 0.0, 0.5, 1.0, 0.5, 0.0, 1.0], [nan] * 8])
    np.testing.assert_almost_equal(_affine((8, 6), (50.2, 10.1), 0.1), [
        [nan] * 8, [0.0, 2.0, 0.0, 3.0, 0.0, 4.0, nan, nan], [nan, nan, 4.0, 0.0, 1.0, 0.0, nan, nan],
        [nan, nan, 0.0, 2.0, 0.0, 3.0, nan, nan], [2.0, 0.0, 3.0, 0.0, 4.0, 0.0, nan, nan],
        [0.0, 4.0, 0.0, 1.0, 0.0, 2.0, nan, nan]])
    np.testing.assert_almost_equal(_affine((8, 6), (49.8, 9.9), 0.1), [
        [nan, nan, 2.0, 0.0, nan, nan, 4.0, 0.0], [nan, nan, 0.0, 4.0, nan, nan, 0.0, 2.0],
        [nan, nan, 1.0, 0.0, 2.0, 0.0, 3.0, 0.0], [nan, nan, 0.0, 3.0, 0.0, 4.0, 0.0, 1.0],
        [nan, nan, 4.0, 0.0, 1.0, 0.0, 2.0, 0.0], [nan] * 8])


def test_ref_affine_order_above_one():
    # tests/test_affine.py:480-497
    with pytest.raises(ValueError):
        ores.upscale(REFL_8X6, ((1.0, 0, 0), (0, 1.0, 0)), (6, 8), 3, False, nan)


def test_ref_affine_matrix_algebra():
    # gridmapping/base.py:436-478 with the affine package's algebra: exact offsets the goldens need
    src = ogrid.regular_grid((8, 6), (50, 10), 0.1)
    assert ogrid.ij_transform_to(ogrid.regular_grid((3, 3), (50.0, 10.0), 0.1), src) == ((1.0, 0.0, 0.0), (0.0, 1.0, 3.0))
    assert ogrid.ij_transform_to(ogrid.regular_grid((8, 6), (50, 10), 0.2), src) == ((2.0, 0.0, 0.0), (0.0, 2.0, -6.0))

"""KC -- the kernels of csrc/coords.cu (kc_stats_init, kc_coords_stats, kc_lon_360) compiled for the host
(tests/hostmath.build_kc) -- against the numpy expressions of gridmapping/coords.py:102-103, 226-252 as the
product's host path evaluates them (gridmapping._estimate_resolution_2d): the reductions behind
``GridMapping.from_device_coords``, without a GPU."""

import math

import numpy as np
import pytest

from xcube_resampling_b200.gridmapping import _abs_no_nan

from .helpers import swath

_ER = 6371000.0


@pytest.fixture(scope="module")
def kc_so(tmp_path_factory):
    from . import hostmath

    try:
        return hostmath.build_kc(str(tmp_path_factory.mktemp("kchost")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise


def _areas_numpy(x, y, geographic):
    """gridmapping._estimate_resolution_2d up to the area image (coords.py:232-250)."""
    x_x, x_y = _abs_no_nan(np.diff(x, axis=1)), _abs_no_nan(np.diff(x, axis=0))
    y_x, y_y = _abs_no_nan(np.diff(y, axis=1)), _abs_no_nan(np.diff(y, axis=0))
    x_x = np.concatenate([x_x, x_x[:, -1:]], axis=1)
    y_x = np.concatenate([y_x, y_x[:, -1:]], axis=1)
    x_y = np.concatenate([x_y, x_y[-1:, :]], axis=0)
    y_y = np.concatenate([y_y, y_y[-1:, :]], axis=0)
    x_abs = np.sqrt(np.square(x_x) + np.square(x_y))
    y_abs = np.sqrt(np.square(y_x) + np.square(y_y))
    if geographic:
        x_r, y_r = np.radians(x_abs), np.radians(y_abs)
        x_abs, y_abs = _ER * np.cos(x_r) * y_r, _ER * y_r
    areas = (x_abs * y_abs).ravel()
    return areas[(areas > 0) & np.isfinite(areas)]


@pytest.mark.parametrize("geographic", [False, True])
@pytest.mark.parametrize("blocks", [1, 3, 7])
def test_extreme_cell_areas_and_lon_flag(kc_so, geographic, blocks):
    from . import hostmath

    x, y = swath(97, 61, theta=17.0, seed=4)
    x[10, 20:30] = np.nan           # holes: their differences count as 0 (_abs_no_nan)
    y[40:43, 5] = np.nan
    x[50, :] = x[49, :]             # a duplicated row: zero-area cells are skipped
    y[50, :] = y[49, :]
    if not geographic:
        x, y = 500000.0 + (x - 10.0) * 78000.0, 5000000.0 + (y - 45.0) * 111000.0
    gt, a_min, a_max = hostmath.coords_stats(kc_so, x, y, geographic, blocks)
    want = _areas_numpy(x, y, geographic)
    assert gt is bool(np.nanmax(x) > 180)
    if geographic:  # libm's cos vs numpy's: a few ulp
        assert a_min == pytest.approx(want.min(), rel=1e-13) and a_max == pytest.approx(want.max(), rel=1e-13)
    else:
        assert a_min == want.min() and a_max == want.max()
    # the resolution estimate the host derives from the two areas (coords.py:253-264)
    res = 0.7 * math.sqrt(a_min) + 0.3 * math.sqrt(a_max)
    assert res == pytest.approx(0.7 * math.sqrt(want.min()) + 0.3 * math.sqrt(want.max()), rel=1e-13)


def test_degenerate_images_and_lon_360(kc_so):
    from . import hostmath

    flat = np.zeros((5, 6))
    gt, a_min, a_max = hostmath.coords_stats(kc_so, flat, flat, True)
    assert gt is False and math.isnan(a_min) and math.isnan(a_max)  # no cell with a positive area
    lon = np.array([[170.0, 175.0, 181.0], [171.0, 176.0, 182.5]])
    lat = np.array([[10.0, 10.0, 10.0], [9.0, 9.0, 9.0]])
    assert hostmath.coords_stats(kc_so, lon, lat, True)[0] is True
    assert hostmath.coords_stats(kc_so, lon - 10.0, lat, True)[0] is False
    x = np.array([[179.0, -179.5, -0.0, 0.0], [np.nan, -1e-300, 12.0, -180.0]])
    got = hostmath.lon_360(kc_so, x)   # helpers.py to_lon_360: x < 0 -> x + 360, in place
    want = np.where(x < 0, x + 360.0, x)
    assert np.array_equal(got, want, equal_nan=True) and math.copysign(1.0, got[0, 2]) == -1.0

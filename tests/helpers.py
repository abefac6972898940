"""Shared test helpers: golden loading, synthetic swaths, oracle grids."""

import os

import numpy as np

from oracle import grid as ogrid

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def grid_from_golden(vec) -> ogrid.RegularGrid:
    w, h, tw, th, x_min, y_min, x_max, y_max, x_res, y_res, j_up = vec
    return ogrid.RegularGrid(int(w), int(h), int(tw), int(th), float(x_min), float(y_min), float(x_max),
                             float(y_max), float(x_res), float(y_res), bool(j_up))


from xcube_resampling_b200.synthetic import covering_grid_args, swath  # noqa: E402,F401


def assert_same(a, b, what=""):
    """Bit-exact comparison with NaN == NaN."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} != {b.shape}"
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} != {b.dtype}"
    if not np.array_equal(a, b, equal_nan=a.dtype.kind == "f"):
        bad = ~((a == b) | ((a != a) & (b != b))) if a.dtype.kind == "f" else (a != b)
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{what}: {bad.sum()} of {a.size} elements differ, first at {idx.tolist()}: "
                             f"{a[bad][:5]} vs {b[bad][:5]}")

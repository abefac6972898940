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


def swath(width, height, res=0.0027, theta=12.0, seed=0, lon0=10.0, lat0=45.0):
    """OLCI-like rotated swath with a smooth sub-pixel perturbation."""
    i = np.arange(width, dtype=np.float64)[None, :]
    j = np.arange(height, dtype=np.float64)[:, None]
    a = (i - width / 2) * res
    b = (height / 2 - j) * res
    th = np.deg2rad(theta)
    lat = lat0 + a * np.sin(th) + b * np.cos(th)
    lon = lon0 + (a * np.cos(th) - b * np.sin(th)) / np.cos(np.deg2rad(lat))
    lon = lon + 0.1 * res * np.sin(i / 37.0 + seed) * np.cos(j / 29.0)
    lat = lat + 0.1 * res * np.cos(i / 31.0) * np.sin(j / 41.0 + seed)
    return lon, lat


def covering_grid_args(x, y, res):
    """(size, xy_min) of a regular grid at *res* covering finite coordinates x, y."""
    xf, yf = x[np.isfinite(x)], y[np.isfinite(y)]
    w = int(np.ceil((xf.max() - xf.min()) / res)) + 1
    h = int(np.ceil((yf.max() - yf.min()) / res)) + 1
    return (w, h), (float(xf.min()) - res / 2, float(yf.min()) - res / 2)


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

"""Seeded synthetic inputs of the benchmark configurations.

Used by bench.py, the smoke test and the GPU parity tests; no real scene data
ships with the repository (the reference's OLCI sample is a missing blob).
"""

from __future__ import annotations

import numpy as np

OLCI_WIDTH, OLCI_HEIGHT, OLCI_BANDS = 4865, 4091, 21
OLCI_RES_DEG = 0.0027  # ~300 m


def swath(width: int, height: int, res: float = OLCI_RES_DEG, theta: float = 12.0, seed: int = 0,
          lon0: float = 10.0, lat0: float = 45.0) -> tuple[np.ndarray, np.ndarray]:
    """OLCI-like swath rotated by ``theta`` degrees with a smooth <=0.1 px perturbation.

    Returns float64 (lon, lat) images of shape (height, width).
    """
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


def covering_grid_args(x: np.ndarray, y: np.ndarray, res: float) -> tuple[tuple[int, int], tuple[float, float]]:
    """(size, xy_min) of a regular grid at ``res`` covering the finite coordinates."""
    xf, yf = x[np.isfinite(x)], y[np.isfinite(y)]
    w = int(np.ceil((xf.max() - xf.min()) / res)) + 1
    h = int(np.ceil((yf.max() - yf.min()) / res)) + 1
    return (w, h), (float(xf.min()) - res / 2, float(yf.min()) - res / 2)


def band_stack(n_bands: int, height: int, width: int, seed: int = 0) -> np.ndarray:
    """(n_bands, height, width) float32 uniform [0, 1)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_bands, height, width), dtype=np.float32)
    for b in range(n_bands):
        rng.random(out=out[b], dtype=np.float32)
    return out

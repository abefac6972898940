"""The reference's OWN numba kernels as a CPU baseline (test / measurement infrastructure).

Not a restatement: this module loads the unmodified ``xcube_resampling`` package and calls its
jitted kernels --

* ``gridmapping/bboxes.py:28-106``  ``compute_ij_bboxes`` (numba ``prange`` over tiles),
* ``rectify.py:424-455``            ``_compute_target_source_ij_sequential`` per target tile,
* ``rectify.py:640-660``            ``_compute_var_image_sequential`` per target tile,

-- tile by tile on a thread pool, which is what dask's threaded scheduler does with the reference's
``nogil`` kernels (``rectify.py:373-419, 605-635``; SURVEY.md 8d "CPU baseline beside it").  The only
glue restated here is the per-tile slicing of those two block functions.

The package is looked for in ``oracle/_ref`` (``pip install --no-deps --target oracle/_ref
<reference>``, done by ``__graft_entry__.build()`` in the build container; git-ignored, it travels to
the GPU box with the snapshot) and, failing that, in ``/root/reference``.  Third-party imports of
the package that the kernels never touch (dask, xarray, pyproj, affine, zarr, dask_image) are
replaced by inert stubs when they are not installed; numba JIT stays on.

Only ``tests/`` and ``bench.py``'s CPU legs may use this module.
"""

from __future__ import annotations

import os
import sys
import types
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIRS = (os.path.join(HERE, "_ref"), "/root/reference")
_STUBBED = ["dask", "dask.array", "dask.array.core", "xarray", "pyproj", "pyproj.crs", "pyproj.transformer",
            "dask_image", "dask_image.ndinterp", "affine", "zarr", "zarr.convenience"]


class _Any:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, n):
        if n.startswith("__") and n.endswith("__"):
            raise AttributeError(n)
        return _Any()

    def __call__(self, *a, **k):
        return _Any()

    def __or__(self, o):
        return self

    def __ror__(self, o):
        return self

    def __getitem__(self, k):
        return _Any()

    def __mro_entries__(self, bases):
        return ()


class _Stub(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__") and n.endswith("__"):
            raise AttributeError(n)
        return _Any()


_loaded = None


def load():
    """(rectify module, bboxes module, where) of the reference, or raises ImportError with the reason."""
    global _loaded
    if _loaded is not None:
        return _loaded
    try:
        import numba  # noqa: F401
    except Exception as e:  # pragma: no cover
        raise ImportError(f"numba is not importable ({e})")
    where = next((d for d in REF_DIRS if os.path.isdir(os.path.join(d, "xcube_resampling"))), None)
    if where is None:
        raise ImportError("reference package not found in oracle/_ref (run __graft_entry__.build() where "
                          "/root/reference exists) nor in /root/reference")
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join("/tmp", "numba_cache_xrs_ref"))
    for name in _STUBBED:
        if name in sys.modules:
            continue
        try:
            __import__(name)
        except Exception:
            m = _Stub(name)
            m.__path__ = []
            sys.modules[name] = m
    if where not in sys.path:
        sys.path.insert(0, where)
    import xcube_resampling.gridmapping.bboxes as B
    import xcube_resampling.rectify as R

    _loaded = (R, B, where)
    return _loaded


# ---------------------------------------------------------------------------
# tile geometry of a regular target grid (gridmapping/base.py:521-533, rectify.py:329-345)
# ---------------------------------------------------------------------------
def tile_boxes(g):
    """Per reference tile: pixel box (i0, j0, i1, j1) and x/y box (x_min, y_min, x_max, y_max)."""
    nty = -(-g.height // g.tile_h)
    ntx = -(-g.width // g.tile_w)
    ij = np.array([[tx * g.tile_w, ty * g.tile_h, min((tx + 1) * g.tile_w, g.width), min((ty + 1) * g.tile_h, g.height)]
                   for ty in range(nty) for tx in range(ntx)], dtype=np.int64)
    if g.is_j_axis_up:
        off = np.array([g.x_min, g.y_min, g.x_min, g.y_min])
        sc = np.array([g.x_res, g.y_res, g.x_res, g.y_res])
        xy = off + sc * ij
    else:
        off = np.array([g.x_min, g.y_max, g.x_min, g.y_max])
        sc = np.array([g.x_res, -g.y_res, g.x_res, -g.y_res])
        xy = off + sc * ij
        xy[:, [1, 3]] = xy[:, [3, 1]]
    return ij, xy


def xy_border(g) -> float:
    ntx_f, nty_f = g.width / g.tile_w, g.height / g.tile_h
    return min(min(2 * ntx_f * g.x_res, 2 * nty_f * g.y_res), min(0.5 * (g.x_max - g.x_min), 0.5 * (g.y_max - g.y_min)))


def rectify_pass(lon, lat, bands, g, methods, n_threads: int, tile_ids=None, uv_delta: float = 1e-3, fill=np.nan,
                 keep_outputs: bool = False):
    """One rectification of ``bands`` (n, h, w) onto grid ``g`` (an ``oracle.grid.RegularGrid``) with
    every method in ``methods``, through the reference kernels: source windows, the ij image (once,
    shared by the methods -- rectify.py:146) and one gather per method, restricted to the target
    tiles ``tile_ids`` (default all).  Returns (output pixel*bands computed, outputs or None)."""
    R, B, _ = load()
    ij_boxes, xy_boxes = tile_boxes(g)
    ids = np.arange(len(ij_boxes)) if tile_ids is None else np.asarray(tile_ids, dtype=np.int64)
    windows = np.full((len(ids), 4), -1, dtype=np.int64)
    B.compute_ij_bboxes(lon, lat, np.ascontiguousarray(xy_boxes[ids]), xy_border(g), 1, windows)
    n_b = bands.shape[0]
    outs = {m: np.full((n_b, g.height, g.width), fill, dtype=bands.dtype) for m in methods} if keep_outputs else None
    y_scale = g.y_res if g.is_j_axis_up else -g.y_res

    def one_tile(k):
        i0, j0, i1, j1 = (int(v) for v in ij_boxes[ids[k]])
        bb = windows[k]
        blk = np.full((2, j1 - j0, i1 - i0), np.nan)
        if bb[0] == -1:
            return (j1 - j0) * (i1 - i0)
        # rectify.py:373-419 (_compute_target_source_ij_block)
        xs = lon[bb[1]:bb[3] + 1, bb[0]:bb[2] + 1]
        ys = lat[bb[1]:bb[3] + 1, bb[0]:bb[2] + 1]
        x_off = g.x_min + i0 * g.x_res
        y_off = g.y_min + j0 * g.y_res if g.is_j_axis_up else g.y_max - j0 * g.y_res
        R._compute_target_source_ij_sequential(xs, ys, int(bb[0]), int(bb[1]), blk, x_off, y_off, g.x_res, y_scale,
                                               uv_delta)
        # rectify.py:605-635 (_compute_var_image_block), once per method
        if not np.all(np.isnan(blk[0])):
            bbox = (int(np.nanmin(blk[0])), int(np.nanmin(blk[1])), min(int(np.nanmax(blk[0])) + 2, bands.shape[-1]),
                    min(int(np.nanmax(blk[1])) + 2, bands.shape[-2]))
            for m in methods:
                win = bands[..., bbox[1]:bbox[3], bbox[0]:bbox[2]].astype(np.float64)
                dst = np.full((n_b, j1 - j0, i1 - i0), fill, dtype=bands.dtype)
                R._compute_var_image_sequential(win, blk, dst, bbox, m)
                if outs is not None:
                    outs[m][:, j0:j1, i0:i1] = dst
        return (j1 - j0) * (i1 - i0)

    with ThreadPoolExecutor(max_workers=max(1, int(n_threads))) as pool:
        px = sum(pool.map(one_tile, range(len(ids))))
    return px * n_b * len(methods), outs

"""Target row-band partitioning for multi-GPU runs (one process per GPU).

The resampling path shards without any exchange step: the target
image is cut into contiguous row bands, rank *r* computes band *r* and needs only
the source footprint of that band.  No collective is involved; results of a band
do not depend on the split because every kernel evaluates pixels independently
(rectify's first-writer rule is resolved per pixel, see DESIGN.md "K1").
"""

from __future__ import annotations

import numpy as np

from .gridmapping import GridMapping


def row_bands(height: int, n_bands: int, align: int = 32) -> list[tuple[int, int]]:
    """Split ``height`` rows into ``n_bands`` contiguous bands with boundaries on multiples
    of ``align`` (bands may be empty when there are fewer aligned blocks than bands)."""
    if n_bands < 1:
        raise ValueError("n_bands must be >= 1")
    n_blocks = -(-height // align)
    edges = [min(height, align * ((n_blocks * k) // n_bands)) for k in range(n_bands)] + [height]
    return [(edges[k], edges[k + 1]) for k in range(n_bands)]


def weighted_row_bands(row_weights, n_bands: int, align: int = 32) -> list[tuple[int, int]]:
    """Contiguous row bands of (approximately) equal total weight, boundaries on multiples of
    ``align``.  ``row_weights[r]`` is the cost of target row r -- for a rotated swath the rows near
    the middle of the image hold far more valid pixels than those at the top and bottom, so equal
    row counts would leave the outer ranks idle."""
    w = np.asarray(row_weights, dtype=np.float64)
    height = w.shape[0]
    if n_bands < 1:
        raise ValueError("n_bands must be >= 1")
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    edges = [0]
    for k in range(1, n_bands):
        if total > 0:
            r = int(np.searchsorted(cum, total * k / n_bands, side="left"))
        else:
            r = height * k // n_bands
        r = min(height, max(edges[-1], align * int(round(r / align))))
        edges.append(r)
    edges.append(height)
    return [(edges[k], edges[k + 1]) for k in range(n_bands)]


def band_tile_rows(target_gm: GridMapping, rows: tuple[int, int]) -> tuple[int, int]:
    """Reference tile rows [ty0, ty1) intersecting target rows [r0, r1)."""
    r0, r1 = rows
    th = target_gm.tile_height
    return r0 // th, -(-r1 // th)


def rectify_band_footprint(tile_boxes: np.ndarray, target_gm: GridMapping, rows: tuple[int, int],
                           src_size: tuple[int, int]) -> tuple[int, int, int, int] | None:
    """Source window (i0, j0, i1, j1), end-exclusive, that a rectify row band can touch.

    Union of the reference tiles' source windows (``rectify.py:342-345, 397-399``) over the
    tile rows intersecting the band, grown by one row/column because a gather tap may read
    index+1 (``rectify.py:719-720``).  ``None`` if no source point maps into the band.
    """
    src_w, src_h = src_size
    ntx = -(-target_gm.width // target_gm.tile_width)
    ty0, ty1 = band_tile_rows(target_gm, rows)
    boxes = np.asarray(tile_boxes).reshape(-1, ntx, 4)[ty0:ty1].reshape(-1, 4)
    boxes = boxes[boxes[:, 0] >= 0]
    if boxes.shape[0] == 0:
        return None
    i0, j0 = int(boxes[:, 0].min()), int(boxes[:, 1].min())
    i1 = min(int(boxes[:, 2].max()) + 2, src_w)
    j1 = min(int(boxes[:, 3].max()) + 2, src_h)
    return i0, j0, i1, j1


def reproject_band_footprint(win_i0: np.ndarray, win_j0: np.ndarray, win_w: int, win_h: int, target_gm: GridMapping,
                             rows: tuple[int, int], src_size: tuple[int, int]) -> tuple[int, int, int, int] | None:
    """Source window (i0, j0, i1, j1), end-exclusive and clipped to the source, that a reproject
    row band can read: the union of the reference tiles' source windows (``reproject.py:385-469``)
    over the tile rows intersecting the band.  ``None`` if it lies entirely outside the source."""
    src_w, src_h = src_size
    ty0, ty1 = band_tile_rows(target_gm, rows)
    i0, j0 = np.asarray(win_i0)[ty0:ty1], np.asarray(win_j0)[ty0:ty1]
    i_lo, j_lo = max(int(i0.min()), 0), max(int(j0.min()), 0)
    i_hi, j_hi = min(int(i0.max()) + int(win_w), src_w), min(int(j0.max()) + int(win_h), src_h)
    if i_lo >= i_hi or j_lo >= j_hi:
        return None
    return i_lo, j_lo, i_hi, j_hi


# ---------------------------------------------------------------------------
# cross-rank bookkeeping (torch.distributed; NCCL on GPUs, gloo in the CPU tests).
# The data path has no collective: these only aggregate timings and unit counts.
# ---------------------------------------------------------------------------
def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())

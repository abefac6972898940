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

"""Multi-GPU host logic on CPU: target row-band partition, per-band source footprints and the
cross-rank bookkeeping, exercised by two ``gloo`` processes (the N>1 path has no data-path
collective -- each rank computes its own band from its own footprint).

The kernels are stood in for by the oracle here (CPU); what is under test is the sharding: a band
computed from ONLY its footprint of the source must equal the same rows of the whole-image result.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import grid as ogrid
from oracle import proj as oproj
from oracle import rectify as orect
from oracle import reproject as orep
from xcube_resampling_b200 import GridMapping
from xcube_resampling_b200 import bands as xbands
from xcube_resampling_b200.synthetic import covering_grid_args, swath


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_row_bands_partition():
    for height, n, align in ((5013, 8, 32), (100, 3, 32), (31, 4, 32), (4500 * 8, 8, 4500)):
        bands = xbands.row_bands(height, n, align)
        assert len(bands) == n and bands[0][0] == 0 and bands[-1][1] == height
        for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(a0 % align == 0 for a0, _ in bands)
    with pytest.raises(ValueError):
        xbands.row_bands(10, 0)


def test_weighted_row_bands_balance():
    h = 5013
    r = np.arange(h)
    weights = np.exp(-((r - h / 2) / (h / 4)) ** 2) + 0.1  # a rotated swath: heavy middle rows
    for n in (2, 4, 8):
        bands = xbands.weighted_row_bands(weights, n, align=32)
        assert bands[0][0] == 0 and bands[-1][1] == h and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        loads = [weights[a:b].sum() for a, b in bands]
        equal = [weights[a:b].sum() for a, b in xbands.row_bands(h, n, align=32)]
        assert max(loads) < 1.1 * sum(loads) / n
        assert max(loads) <= max(equal)
    assert xbands.weighted_row_bands(np.zeros(100), 4, 32)[-1][1] == 100


def _rectify_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, res, tile = 160, 120, 0.0027, 64
        lon, lat = swath(w, h, res=res, theta=12.0, seed=2)
        size, xy_min = covering_grid_args(lon, lat, res)
        gm = GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=tile)
        g = ogrid.regular_grid(size, xy_min, res, tile_size=tile)
        data = np.random.default_rng(0).random((3, h, w)).astype(np.float32)
        rows = xbands.row_bands(gm.height, world, align=32)[rank]
        boxes = orect.source_windows(lon, lat, g)
        fp = xbands.rectify_band_footprint(boxes, gm, rows, (w, h))
        assert fp is not None
        i0, j0, i1, j1 = fp
        # the band's ij values (whole-image oracle, band rows) must stay inside the footprint
        ij = orect.rectify_ij(lon, lat, g, windows=boxes)[:, rows[0]:rows[1]]
        ok = np.isfinite(ij[0])
        assert ok.any()
        assert int(ij[0][ok].min()) >= i0 and min(int(ij[0][ok].max()) + 1, w - 1) <= i1 - 1
        assert int(ij[1][ok].min()) >= j0 and min(int(ij[1][ok].max()) + 1, h - 1) <= j1 - 1
        # gather from ONLY the footprint window, with shifted indices and true-edge clamping
        masked = np.full_like(data, np.nan)
        masked[:, j0:j1, i0:i1] = data[:, j0:j1, i0:i1]
        for method in ("nearest", "bilinear"):
            part = orect.gather(masked, ij, method, np.nan)
            full = orect.gather(data, orect.rectify_ij(lon, lat, g, windows=boxes), method, np.nan)
            assert np.array_equal(part, full[:, rows[0]:rows[1]], equal_nan=True), (rank, method)
        # bookkeeping: per-rank unit counts sum, timings max
        units = xbands.sum_over_ranks(float((rows[1] - rows[0]) * gm.width))
        assert units == float(gm.height * gm.width)
        assert xbands.max_over_ranks(float(rank + 1)) == float(world)
        # every rank's rows gathered -> exact cover of the image
        got = [None] * world
        dist.all_gather_object(got, rows)
        assert got[0][0] == 0 and got[-1][1] == gm.height and all(a[1] == b[0] for a, b in zip(got, got[1:]))
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _reproject_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        utm, geo = oproj.from_epsg(32632), oproj.from_epsg(4326)
        tgt_gm = GridMapping.regular((96, 128), (399960.0, 990240.0), 10.0, "EPSG:32632", tile_size=32)
        g = ogrid.regular_grid((96, 128), (399960.0, 990240.0), 10.0, tile_size=32)
        box = oproj.transform_bounds(utm, geo, *tgt_gm.xy_bbox)
        res = 0.0001
        x_min, y_min = np.floor(box[0] / res) * res - 4 * res, np.floor(box[1] / res) * res - 4 * res
        w = int(np.ceil((box[2] - x_min) / res)) + 4
        h = int(np.ceil((box[3] - y_min) / res)) + 4
        sg = ogrid.regular_grid((w, h), (x_min, y_min), res)
        xs, ys = ogrid.x_centres(sg), ogrid.y_centres(sg)
        data = np.random.default_rng(1).random((2, h, w)).astype(np.float32)
        args = (float(xs[0]), float(ys[0]), res, res, float(ys[1] - ys[0]))
        win = orep.source_windows(*args, w, h, g, utm, geo)
        rows = xbands.row_bands(tgt_gm.height, world, align=32)[rank]
        fp = xbands.reproject_band_footprint(win["i0"], win["j0"], win["win_w"], win["win_h"], tgt_gm, rows, (w, h))
        assert fp is not None
        i0, j0, i1, j1 = fp
        assert (i1 - i0) * (j1 - j0) < w * h  # a band needs less than the whole source
        masked = np.full_like(data, -777.0)   # anything outside the footprint must never be read
        masked[:, j0:j1, i0:i1] = data[:, j0:j1, i0:i1]
        for method in ("nearest", "bilinear"):
            full = orep.reproject(data, *args, g, utm, geo, method, np.nan)
            part = orep.reproject(masked, *args, g, utm, geo, method, np.nan)
            assert np.array_equal(part[:, rows[0]:rows[1]], full[:, rows[0]:rows[1]], equal_nan=True), (rank, method)
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _footprint_worker(rank, world, port, result_dir):
    """The product multi-GPU path's host logic (multigpu.py) with the kernels stood in for by numpy /
    the oracle: slab scans, the MIN exchange over the process group, ragged upload plan, and the
    claim that a band computed from ONLY its footprint equals the whole-image result."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xcube_resampling_b200 import multigpu

        from .helpers import INT32_MAX, quad_footprints_np

        w, h, res, tile, group = 200, 170, 0.0027, 64, 32
        lon, lat = swath(w, h, res=res, theta=-17.0, seed=5)
        lon[60:63, 40:70] = np.nan  # a hole: quads with fewer than three finite vertices do not count
        size, xy_min = covering_grid_args(lon, lat, res)
        gm = GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=tile)
        g = ogrid.regular_grid(size, xy_min, res, tile_size=tile)
        edges = multigpu.default_band_edges(gm.height, world)
        assert edges[0] == 0 and edges[-1] == gm.height and len(edges) == world + 1
        # slab scan (numpy stand-in for xrs_band_quad_footprints) + exchange
        slabs = multigpu.source_slabs(h, world, group)
        assert slabs[0][0] == 0 and slabs[-1][1] == h and all(a[1] == b[0] and a[0] % group == 0 for a, b in zip(slabs, slabs[1:]))
        s0, s1 = slabs[rank]
        part = quad_footprints_np(lon, lat, g, edges, group, rows=(s0, min(h, s1 + 1)))
        merged = multigpu.DistExchange().merge(rank, torch.from_numpy(part.copy())).numpy()
        whole = quad_footprints_np(lon, lat, g, edges, group)
        assert np.array_equal(merged, whole), "merged slab tables differ from the whole-swath table"
        assert np.array_equal(multigpu.merge_minform_host([part, whole]), whole)
        # upload plan of this rank's band
        window, segments, n_px = multigpu.footprint_segments(merged[rank], h, w, group, merge_groups=2, align=8)
        assert window is not None and 0 < n_px < h * w, (window, n_px)
        assert segments[0][0] == window[0] and segments[-1][1] == window[1]
        assert all(a[1] <= b[0] for a, b in zip(segments, segments[1:])), "segments overlap"
        keep = np.zeros((h, w), dtype=bool)
        for j0, j1, i0, i1 in segments:
            keep[j0:j1, i0:i1] = True
        # every vertex of a footprint quad and every tap reachable through it (+2) is uploaded
        fp = merged[rank].astype(np.int64)
        for gi in range(fp.shape[0]):
            if fp[gi, 0] == INT32_MAX:
                continue
            lo, hi = fp[gi, 0], -fp[gi, 1]
            rows_needed = slice(gi * group, min(h, gi * group + group + 2))
            assert keep[rows_needed, lo:min(w, hi + 3)].all(), f"group {gi}: footprint not covered by the segments"
        # K1 restricted to the footprint (quads outside never read) + gather from the uploaded data only
        rows = (edges[rank], edges[rank + 1])
        boxes = orect.source_windows(lon, lat, g)
        vert = np.zeros((h, w), dtype=bool)
        for gi in range(fp.shape[0]):
            if fp[gi, 0] == INT32_MAX:
                continue
            lo, hi = fp[gi, 0], -fp[gi, 1]
            vert[gi * group:min(h, gi * group + group + 1), lo:hi + 2] = True
        assert not (vert & ~keep).any()
        lon_m, lat_m = np.where(vert, lon, np.nan), np.where(vert, lat, np.nan)
        ij_full = orect.rectify_ij(lon, lat, g, windows=boxes)
        ij_band = orect.rectify_ij(lon_m, lat_m, g, windows=boxes)[:, rows[0]:rows[1]]
        assert np.array_equal(ij_band, ij_full[:, rows[0]:rows[1]], equal_nan=True), "band ij from the footprint differs"
        data = np.random.default_rng(0).random((3, h, w)).astype(np.float32)
        masked = np.where(keep[None], data, np.float32(-777.0))
        for method in ("nearest", "bilinear", "triangular"):
            part_out = orect.gather(masked, ij_band, method, np.nan)
            full_out = orect.gather(data, ij_full, method, np.nan)[:, rows[0]:rows[1]]
            assert np.array_equal(part_out, full_out, equal_nan=True), (rank, method)
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_rectify_worker, _reproject_worker, _footprint_worker])
def test_two_rank_row_bands_gloo(worker, tmp_path):
    world = 2
    mp.spawn(worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_band_partitions_cover_the_image_for_any_band_count():
    """Seeded sweep over heights, band counts (1..11, more bands than aligned blocks included), alignments and
    row weights (zero rows, all-zero images): the bands are contiguous, ordered and cover [0, height)."""
    import numpy as np

    from xcube_resampling_b200.bands import row_bands, weighted_row_bands

    rng = np.random.default_rng(0)
    for _ in range(1500):
        h, n, align = int(rng.integers(1, 400)), int(rng.integers(1, 12)), int(rng.choice([1, 8, 32]))
        w = rng.random(h) * (rng.random(h) > 0.3)
        if rng.random() < 0.1:
            w[:] = 0
        for bands in (weighted_row_bands(w, n, align=align), row_bands(h, n, align=align)):
            assert len(bands) == n and bands[0][0] == 0 and bands[-1][1] == h
            assert all(a <= b for a, b in bands) and all(bands[k][1] == bands[k + 1][0] for k in range(n - 1))

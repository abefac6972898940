"""Target row bands on several GPUs -- the product-level multi-GPU path.

The reference distributes every algorithm over target tiles as dask tasks (``rectify.py:263-309``,
``reproject.py:189-265``, ``dask.py:41-135``).  On one box with several B200s the unit of
distribution is a **row band of the target**: participant *k* of *N* (a host thread in
``rectify_dataset(..., devices=[...])``, or one ``torchrun`` rank calling :func:`rectify_band`)
computes target rows ``band_edges[k]:band_edges[k+1]`` on its GPU and writes them into its rows of
the (page-locked) output arrays.  A participant uploads only what its band needs:

rectify
    1. a **slab** (1/N of the rows) of the source coordinates, scanned once for the per-tile source
       windows (K0) and for the bands' ragged quad footprints (``xrs_band_quad_footprints``);
    2. the partial tables of all participants are merged with an element-wise MIN -- the only
       exchange step of the path (an NCCL all-reduce of a few KB between processes,
       ``numpy.minimum`` between threads);
    3. the coordinates and the data bands of the band's footprint only (strided 2-D copies of the
       ragged, for a rotated swath diagonal, strip), K1 restricted to that footprint, K2 per band
       chunk, download into the band's rows.

reproject
    the band's source footprint is a rectangle known on the host (``ReprojectPlan.footprint``): no
    exchange step at all.

Results are bit-identical to the single-GPU path (every quad that can claim a pixel of the band is
inside the band's footprint; K1's first-writer rule is resolved per pixel).
"""

from __future__ import annotations

import threading

import numpy as np
import torch

from ._affinity import bind_to_device
from . import _dev
from ._lib import check, load
from ._pipeline import GatherPipeline, SourceGroup  # noqa: F401
from .bands import row_bands

INT32_MAX = np.iinfo(np.int32).max


# ---------------------------------------------------------------------------
# exchange of the min-form tables
# ---------------------------------------------------------------------------
class NoExchange:
    """A single participant: nothing to merge."""

    n = 1

    def merge(self, k: int, table: torch.Tensor) -> torch.Tensor:
        return table


class ThreadExchange:
    """N host threads of one process, one GPU each: tables meet in host memory."""

    def __init__(self, n: int):
        self.n = int(n)
        self._barrier = threading.Barrier(self.n)
        self._parts: list = [None] * self.n
        self._merged = None

    def merge(self, k: int, table: torch.Tensor) -> torch.Tensor:
        self._parts[k] = _dev.to_host(table)
        if self._barrier.wait() == 0:
            self._merged = np.minimum.reduce(self._parts)
        self._barrier.wait()
        out = torch.from_numpy(self._merged).to(table.device)
        self._barrier.wait()  # nobody overwrites _parts before everyone has read the result
        return out

    def abort(self):
        self._barrier.abort()


class DistExchange:
    """One process per GPU (``torchrun``): NCCL all-reduce(MIN) over NVLink on the rank's current
    stream (gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self._dist = dist
        self.group = group
        self.n = dist.get_world_size(group)

    def merge(self, k: int, table: torch.Tensor) -> torch.Tensor:
        self._dist.all_reduce(table, op=self._dist.ReduceOp.MIN, group=self.group)
        return table


def merge_minform_host(parts) -> np.ndarray:
    """Element-wise MIN of min-form tables (what every exchange computes)."""
    return np.minimum.reduce([np.asarray(p, dtype=np.int32) for p in parts])


# ---------------------------------------------------------------------------
# geometry helpers (host)
# ---------------------------------------------------------------------------
def source_slabs(src_h: int, n: int, group: int) -> list[tuple[int, int]]:
    """Split the source rows into ``n`` slabs whose starts are multiples of ``group`` (the quad row
    group of the footprint table); slabs may be empty when the image is short."""
    n_groups = -(-src_h // group)
    edges = [min(src_h, group * ((n_groups * k) // n)) for k in range(n)] + [src_h]
    return [(edges[k], edges[k + 1]) for k in range(n)]


def footprint_segments(fp_minform: np.ndarray, src_h: int, src_w: int, group: int, merge_groups: int = 1,
                       align: int = 32):
    """Upload plan of one band from its row of the footprint table.

    ``fp_minform``: (n_groups, 2) int32 (c_min, -c_max) per group of ``group`` quad rows.  Returns
    ``(window, segments, n_px)``: the resident source rows ``(j0, j1)``, a list of
    ``(j0, j1, i0, i1)`` rectangles to upload (vertex rows / columns; they cover the quads'
    vertices and every tap a gather through those quads can read: index + 2 in both directions,
    ``rectify.py:689-727``) and the number of source pixels in them; ``(None, [], 0)`` if the band
    sees no source quad."""
    fp = np.asarray(fp_minform, dtype=np.int64).reshape(-1, 2)
    lo, hi = fp[:, 0], -fp[:, 1]
    valid = fp[:, 0] != INT32_MAX
    if not valid.any():
        return None, [], 0
    g_idx = np.nonzero(valid)[0]
    g_min, g_max = int(g_idx[0]), int(g_idx[-1])
    blocks = []  # (j0, j1, i0, i1) per row block
    for rb in range(g_min, g_max + 2):
        j0 = rb * group
        if j0 >= src_h:
            break
        cands = [g for g in (rb - 1, rb) if 0 <= g < len(valid) and valid[g]]
        if not cands:
            continue
        i0 = min(int(lo[g]) for g in cands)
        i1 = max(int(hi[g]) for g in cands) + 3
        own = rb < len(valid) and valid[rb]
        j1 = min(src_h, j0 + group) if own else min(src_h, j0 + 2)
        i0 = max(0, (i0 // align) * align)
        i1 = min(src_w, -(-i1 // align) * align)
        blocks.append((j0, j1, i0, i1))
    segments = []
    k = 0
    while k < len(blocks):
        j0, j1, i0, i1 = blocks[k]
        n = 1
        while (n < merge_groups and k + n < len(blocks) and blocks[k + n][0] == j1
               and blocks[k + n][1] - blocks[k + n][0] == group):
            j1 = blocks[k + n][1]
            i0, i1 = min(i0, blocks[k + n][2]), max(i1, blocks[k + n][3])
            n += 1
        segments.append((j0, j1, i0, i1))
        k += n
    window = (segments[0][0], segments[-1][1])
    n_px = sum((j1 - j0) * (i1 - i0) for j0, j1, i0, i1 in segments)
    return window, segments, n_px


# ---------------------------------------------------------------------------
# rectify: one participant
# ---------------------------------------------------------------------------
class RectifyBandStats:
    """Byte counts of one band call (what bench.py reports as h2d / d2h bytes)."""

    def __init__(self):
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.src_px = 0       # source pixels of the footprint (per band of data)
        self.window = None


class RectifyBandJob:
    """Participant ``k``'s share of ONE scene: target rows ``band_edges[k]:band_edges[k+1]`` of every
    target in ``groups`` (``_pipeline.SourceGroup`` list; ``Target.out_host`` arrays are (bands, H, W)
    -- or band-only arrays with ``Target.row0`` -- and the rows of the band are filled) from host
    coordinates ``x``, ``y`` (h, w) float64.

    Three phases, so that a rank working through a sequence of scenes can overlap them
    (:func:`rectify_band_stream`): :meth:`prologue` -- slab scan, table exchange, footprint, K1, all on
    ``stream`` (blocks only on that stream: the tiny footprint table has to reach the host);
    :meth:`enqueue` -- the band-chunk pipeline of all data variables on the current stream;
    :meth:`wait`.  Every participant must run the same sequence of prologues with the same ``x``,
    ``y``, ``target_gm`` and ``band_edges``; ``exchange`` merges the partial tables."""

    def __init__(self, x: np.ndarray, y: np.ndarray, groups: list, target_gm, band_edges, k: int, exchange,
                 device=None, plan=None, stats: RectifyBandStats | None = None, chunk_bands: int = 4, stream=None):
        from .rectify import RectifyPlan

        self.lib = load()
        self.dev = _dev.require_cuda(device)
        self.stats = stats if stats is not None else RectifyBandStats()
        self.exchange, self.k, self.n = exchange, int(k), exchange.n
        self.edges = [int(e) for e in band_edges]
        if len(self.edges) != self.n + 1:
            raise ValueError("band_edges must have one more entry than there are participants")
        self.rows = (self.edges[k], self.edges[k + 1])
        if x.dtype != np.float64 or y.dtype != np.float64 or y.shape != x.shape or x.ndim != 2:
            raise TypeError("source coordinates must be 2-D float64 arrays of the same shape")
        if x.strides[1] != 8 or y.strides[1] != 8:
            x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
        self.x, self.y, self.groups, self.gm = x, y, groups, target_gm
        self.h, self.w = x.shape
        self.group = int(self.lib.xrs_quad_row_group())
        self.n_groups = -(-(self.h - 1) // self.group)
        rows = self.rows
        if plan is None or plan.rows != (rows if rows[1] > rows[0] else (0, 1)) or plan.device != self.dev:
            plan = RectifyPlan(target_gm, self.dev, rows=(rows if rows[1] > rows[0] else (0, 1)))
        self.plan = plan
        self.chunk_bands = chunk_bands
        self.stream = stream
        self.pipe = None
        self._ij = self._ready = self._window = self._segments = None
        self._done = False

    def _fill_band(self):
        r0, r1 = self.rows
        for grp in self.groups:
            for tgt in grp.targets:
                tgt.out_host[:, r0 - tgt.row0:r1 - tgt.row0, :] = np.asarray(tgt.fill).astype(tgt.out_dtype)

    def prologue(self) -> None:
        from ._pipeline import copy2d

        dev, lib, plan, st = self.dev, self.lib, self.plan, self.stats
        h, w, n, k, group, n_groups = self.h, self.w, self.n, self.k, self.group, self.n_groups
        torch.cuda.set_device(dev)
        stream = self.stream if self.stream is not None else torch.cuda.current_stream(dev)
        with torch.cuda.stream(stream):
            n_tiles = plan.ntx * plan.nty
            # 1. slab scan: per-tile source windows + band footprints, partial tables in min-form
            table = _dev.empty((4 * n_tiles + 2 * n * n_groups,), np.int32, dev)
            check(lib.xrs_minform_init(_dev.ptr(table), table.numel(), _dev.stream_ptr(dev)), "xrs_minform_init")
            s0, s1 = source_slabs(h, n, group)[k]
            wp = -(-w // 16) * 16
            if s1 > s0:
                s1v = min(h, s1 + 1)  # first vertex row of the next slab closes this slab's last quad row
                xs = _dev.empty((s1v - s0, wp), np.float64, dev)
                ys = _dev.empty((s1v - s0, wp), np.float64, dev)
                for buf, host in ((xs, self.x), (ys, self.y)):
                    copy2d(buf.data_ptr(), wp * 8, 0, host.__array_interface__["data"][0] + s0 * host.strides[0],
                           host.strides[0], 0, w * 8, s1v - s0, 1, dev)
                st.h2d_bytes += 2 * (s1v - s0) * w * 8
                plan.scan_slab(xs, ys, s0, s1 - s0, h, w, self.edges, table)
            # 2. the exchange step
            table = self.exchange.merge(k, table)
            if self.rows[1] <= self.rows[0]:
                self._done = True
                return
            tile_boxes = plan.finalize_windows(table, w, h)
            fp = _dev.to_host(table[4 * n_tiles:].view(n, n_groups, 2)[k])
            window, segments, n_px = footprint_segments(fp, h, w, group)
            st.window, st.src_px = window, n_px
            if window is None:  # no source quad reaches the band: everything is fill
                self._fill_band()
                self._done = True
                return
            # 3. coordinates of the footprint, K1 restricted to it
            fj0, fj1 = window
            xw = _dev.empty((fj1 - fj0, wp), np.float64, dev)
            yw = _dev.empty((fj1 - fj0, wp), np.float64, dev)
            for buf, host in ((xw, self.x), (yw, self.y)):
                for (j0, j1, i0, i1) in segments:
                    copy2d(buf.data_ptr() + ((j0 - fj0) * wp + i0) * 8, wp * 8, 0,
                           host.__array_interface__["data"][0] + j0 * host.strides[0] + i0 * 8, host.strides[0], 0,
                           (i1 - i0) * 8, j1 - j0, 1, dev)
            st.h2d_bytes += 2 * n_px * 8
            col_ranges = table[4 * n_tiles:].view(n, n_groups, 2)[k]
            self._ij = plan.ij_window(xw, yw, fj0, h, w, tile_boxes, col_ranges)
            self._ready = torch.cuda.Event()
            self._ready.record(stream)
            self._window, self._segments = window, segments
            self._buffers = (xw, yw, table)  # alive until the kernels that read them have run

    def enqueue(self) -> None:
        """The data bands of the footprint through K2, band chunk by band chunk, on the current stream."""
        from . import rectify as xrect
        from .rectify import gather_ij

        if self._done:
            return
        dev = self.dev
        torch.cuda.current_stream(dev).wait_event(self._ready)
        fj0 = self._window[0]
        ij, h, w = self._ij, self.h, self.w
        self.pipe = GatherPipeline(dev, (h, w), self.gm.width, self.rows, src_window=self._window,
                                   segments=self._segments, chunk_bands=self.chunk_bands)

        def process(src_view, tgt, out_view, b0):
            gather_ij(src_view, ij, tgt.method, tgt.fill, out=out_view, window_origin=(0, fj0), full_size=(w, h))

        def process_pair(src_view, tgt_interp, tgt_near, out_interp, out_near, b0):
            xrect.gather_ij_pair(src_view, ij, tgt_interp.method, tgt_interp.fill, tgt_near.fill, out_interp, out_near,
                                 window_origin=(0, fj0), full_size=(w, h))

        self.pipe.run(self.groups, process, wait=False, process_pair=process_pair if xrect.DUAL_GATHER else None)

    def wait(self) -> None:
        if self.pipe is not None:
            self.pipe.wait()
            self.stats.h2d_bytes += self.pipe.h2d_bytes
            self.stats.d2h_bytes += self.pipe.d2h_bytes
            self.pipe = None
        self._buffers = None
        self._done = True


def rectify_band(x: np.ndarray, y: np.ndarray, groups: list, target_gm, band_edges, k: int, exchange,
                 device=None, plan=None, stats: RectifyBandStats | None = None, chunk_bands: int = 4) -> None:
    """One scene, one participant, start to finish (see :class:`RectifyBandJob`).  Blocks until the
    band's rows are in host memory."""
    job = RectifyBandJob(x, y, groups, target_gm, band_edges, k, exchange, device, plan, stats, chunk_bands)
    job.prologue()
    job.enqueue()
    job.wait()


def rectify_band_stream(scenes, target_gm, band_edges, k: int, exchange, device=None, plans=None,
                        chunk_bands: int = 4) -> list:
    """Participant ``k`` working through a SEQUENCE of scenes (an iterable of ``(x, y, groups)``): while
    the band-chunk pipeline of scene s streams data through the copy engines, the prologue of scene
    s + 1 (slab scan, table exchange, footprint, K1) already runs on a second, high-priority stream
    with its own plan, so the copy engines never wait for it.  Returns one
    :class:`RectifyBandStats` per scene.  Single host thread: the exchange steps of all participants
    stay in scene order."""
    from .rectify import RectifyPlan

    dev = _dev.require_cuda(device)
    torch.cuda.set_device(dev)
    edges = [int(e) for e in band_edges]
    rows = (edges[k], edges[k + 1])
    rows_p = rows if rows[1] > rows[0] else (0, 1)
    if plans is None:
        plans = [RectifyPlan(target_gm, dev, rows=rows_p) for _ in range(2)]
    side = torch.cuda.Stream(dev, priority=-1)
    stats, prev = [], None
    for s, (x, y, groups) in enumerate(scenes):
        st = RectifyBandStats()
        stats.append(st)
        job = RectifyBandJob(x, y, groups, target_gm, edges, k, exchange, dev, plans[s % 2], st, chunk_bands,
                             stream=side)
        job.prologue()          # overlaps the previous scene's streaming
        job.enqueue()
        if prev is not None:
            prev.wait()
        prev = job
    if prev is not None:
        prev.wait()
    return stats


def default_band_edges(height: int, n: int, align: int = 32) -> list[int]:
    """Equal-height bands (equal output bytes per GPU: the end-to-end path is bound by the D2H copy
    of the results), boundaries on multiples of ``align`` rows."""
    bands = row_bands(height, n, align)
    return [b[0] for b in bands] + [height]


def run_on_devices(devices, worker) -> None:
    """``worker(k, device, exchange)`` on one host thread per device; re-raises the first error."""
    devs = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in devices]
    n = len(devs)
    if n == 1:
        worker(0, devs[0], NoExchange())
        return
    exchange = ThreadExchange(n)
    errors: list = [None] * n

    def run(k):
        try:
            torch.cuda.set_device(devs[k])
            bind_to_device(devs[k].index or 0)  # this thread page-locks the band's staging buffers: keep them local
            worker(k, devs[k], exchange)
        except BaseException as e:  # noqa: BLE001 - reported to the caller below
            errors[k] = e
            exchange.abort()

    threads = [threading.Thread(target=run, args=(k,), name=f"xrs-band-{k}") for k in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e

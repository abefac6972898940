#!/usr/bin/env python
"""One rank's share of the N-rank scan phase, timed alone on one GPU.

    python tools/microbench/scan_slab_bench.py [--ranks 8] [--reps 50]

Scans slab 0..N-1 (1/N of the rows of the bench's OLCI-shaped swath each) the way rank r does inside
bench.py's step -- `RectifyPlan.scan_slab`: tile windows of the slab's points + quad footprints of
N target row bands, both into one min-form table -- and prints the time per scan from CUDA events
and from libxrs's own per-kernel events.  The slab kernels are latency-bound (a few hundred rows),
which is why they are looked at separately from the step."""

import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    import bench
    from xcube_resampling_b200 import GridMapping, _dev, _lib, multigpu
    from xcube_resampling_b200 import rectify as xrect

    lon, lat, _bands, size, xy_min, res = bench.make_scene(n_bands=1)
    h, w = lon.shape
    dev = torch.device("cuda", 0)
    gm = GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=bench.TILE)
    x_dev, y_dev = _dev.to_device(lon, dev), _dev.to_device(lat, dev)
    lib = _lib.load()
    n = args.ranks
    group = int(lib.xrs_quad_row_group())
    n_groups = -(-(h - 1) // group)
    plan = xrect.RectifyPlan(gm, dev)
    n_tiles = plan.ntx * plan.nty
    edges = [round(k * size[1] / n) for k in range(n)] + [size[1]]
    table = _dev.empty((4 * n_tiles + 2 * n * n_groups,), np.int32, dev)
    slabs = multigpu.source_slabs(h, n, group)

    def scan(r):
        s0, s1 = slabs[r]
        _lib.check(lib.xrs_minform_init(_dev.ptr(table), table.numel(), _dev.stream_ptr(dev)), "init")
        plan.scan_slab(x_dev[s0:min(h, s1 + 1)], y_dev[s0:min(h, s1 + 1)], s0, s1 - s0, h, w, edges, table)

    for r in range(n):
        scan(r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        for r in range(n):
            scan(r)
    e1.record()
    torch.cuda.synchronize()
    _lib.profile_enable(True)  # a second, eager pass with libxrs's events around every launch
    for r in range(n):
        scan(r)
    kernels = {k: round(1e3 * ms / cnt, 2) for k, (ms, cnt) in _lib.profile_collect().items()}
    _lib.profile_enable(False)
    print(json.dumps({"ranks": n, "slab_rows": slabs[0][1] - slabs[0][0], "us_per_scan": 1e3 * e0.elapsed_time(e1) / (args.reps * n),
                      "kernels_us": kernels}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Copy-only probe of the host <-> device fabric with N ranks copying at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/microbench/pcie_probe_ranks.py [--mb 1024] [--reps 5]

Every rank page-locks two host buffers on the NUMA node of its GPU and times, with CUDA events
between barriers: H2D alone, D2H alone, both directions at once (two streams).  Rank 0 prints one
JSON line with the per-rank minimum / median and the aggregate rates.  This is the ceiling of the
end-to-end leg of bench.py at N ranks: no kernel runs here.
"""

import argparse
import json
import os

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.mb << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    h_out.fill_(2)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(h2d, d2h):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        main = torch.cuda.current_stream(dev)
        e0.record(main)
        s1.wait_stream(main)
        s2.wait_stream(main)
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        main.wait_stream(s1)
        main.wait_stream(s2)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    out = {}
    for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        run(h2d, d2h)
        ms = [run(h2d, d2h) for _ in range(args.reps)]
        gbs = (int(h2d) + int(d2h)) * n / (np.median(ms) * 1e-3) / 1e9
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            vals = [float(v.item()) for v in allv]
        else:
            vals = [gbs]
        out[name] = {"per_rank_gbs_min": min(vals), "per_rank_gbs_median": float(np.median(vals)),
                     "aggregate_gbs": float(sum(vals))}
    if rank == 0:
        print(json.dumps({"probe": "pcie copy-only", "ranks": world, "mb_per_direction": args.mb, **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

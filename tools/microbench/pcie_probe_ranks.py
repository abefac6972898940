#!/usr/bin/env python
"""Copy-only probe of the host <-> device fabric with N ranks copying at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/microbench/pcie_probe_ranks.py [--mb 1024] [--reps 5]

Every rank page-locks two host buffers (with --bind: after pinning itself to the CPUs NVML lists as
local to its GPU, so that first touch places them on that NUMA node) and times, with CUDA events
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
    ap.add_argument("--bind", action="store_true", help="bind the rank to its GPU's CPUs before allocating")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
    from xcube_resampling_b200._affinity import bind_to_device, device_cpus
    allowed = sorted(os.sched_getaffinity(0))
    local_cpus = sorted(device_cpus(local))
    bound = sorted(bind_to_device(local)) if args.bind else []
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
    def span(c):
        return f"{c[0]}-{c[-1]} ({len(c)})" if c else "none"
    info = f"rank {rank}: allowed {span(allowed)} device-local {span(local_cpus)} bound {span(bound)}"
    if world > 1:
        infos = [None] * world
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        out["cpu_sets"] = infos
        out["bind"] = bool(args.bind)
        print(json.dumps({"probe": "pcie copy-only", "ranks": world, "mb_per_direction": args.mb, **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

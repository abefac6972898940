#!/usr/bin/env python
"""bench.py -- throughput of the rectify hot path on B200 (BASELINE.json configs[1], "C2").

Workload: Sentinel-3 OLCI-shaped swath (4865 x 4091 lon/lat float64, 21 float32 bands) rectified to
a regular 0.0027 deg (~300 m) EPSG:4326 grid with nearest AND bilinear interpolation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One *step* on N GPUs rectifies N scenes; every scene is cut into N target row bands and rank r
computes band r of every scene (no collective; weak scaling: one scene-equivalent per GPU and step).
A scene pass = rectify(nearest) + rectify(bilinear), each a complete K0 (tile windows) + K1 (ij
image) + K2 (gather of all 21 bands).  ``value`` is output Mpix*band/s with the inputs resident in
HBM; ``e2e`` is the same metric through ``rectify_dataset`` with host (pinned) inputs and host
outputs, copies inside the timed region.  Rank 0 prints ONE JSON line.

The device-resident part is timed twice over K steps each: eagerly, launch after launch, with a
CUDA event pair around every kernel (the per-kernel durations of ``roofline``), then as K replays
of a CUDA graph of the same step in which the two methods' passes run as concurrent chains
(``value``; ``--no-graph`` takes it from the eager region instead).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "output Mpix*band/s"
UNIT = "Mpix*band/s"
TILE = 512
METHODS = ("nearest", "bilinear")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the scene (debugging only; 1.0 = BASELINE config)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (debugging only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (debugging only)")
    ap.add_argument("--two-step", action="store_true",
                    help="device-resident leg: xrs_rectify_ij + xrs_gather_ij instead of the fused xrs_rectify_gather")
    ap.add_argument("--no-graph", action="store_true", help="take `value` from the eager, sequential region instead of the CUDA-graph replays")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------
def make_scene(scale=1.0, seed=0, n_bands=None):
    from xcube_resampling_b200 import synthetic as syn

    w = max(64, int(round(syn.OLCI_WIDTH * scale)))
    h = max(64, int(round(syn.OLCI_HEIGHT * scale)))
    nb = syn.OLCI_BANDS if n_bands is None else n_bands
    lon, lat = syn.swath(w, h, res=syn.OLCI_RES_DEG, theta=12.0, seed=seed)
    size, xy_min = syn.covering_grid_args(lon, lat, syn.OLCI_RES_DEG)
    bands = syn.band_stack(nb, h, w, seed=seed)
    return lon, lat, bands, size, xy_min, syn.OLCI_RES_DEG


def workload_config(w, h, nb, size, n_gpus):
    return {
        "workload": "rectify_dataset: OLCI-shaped swath -> regular 300 m EPSG:4326 grid, nearest + bilinear",
        "source": f"{w}x{h} lon/lat float64, {nb} float32 bands",
        "target": f"{size[0]}x{size[1]} @0.0027deg, reference tile_size {TILE}",
        "scene_pass": "rectify(nearest)+rectify(bilinear); each = K0 tile windows + K1 claims + K2 gather of "
                      "21 bands (fused xrs_rectify_gather: ij resolved in registers)",
        "scenes_per_step": n_gpus,
        "partition": "target row bands of equal work (valid pixels per row), rank r = band r of every scene, "
                     "no collective; the timed step is one CUDA graph, nearest and bilinear passes as two chains on two streams",
        "l2": "inputs (2.0 GB) and outputs (3.3 GB per method) exceed the 126 MB L2; no explicit flush",
    }


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled while the timed region runs.

    In-process NVML (the counters behind nvidia-smi's clocks.sm / clocks_event_reasons.* columns)
    every 10 ms from a thread that is joined when the timed region ends -- the device-resident
    region lasts tens of milliseconds, too short for ``nvidia-smi -lms``.  Falls back to an
    ``nvidia-smi -lms 100`` child (killed and reaped in stop()) when the NVML binding is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.sm, self.smax, self.reasons = [], [], set()
        self.source = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.smax.append(float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)))

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                        mask = int(reasons_fn(handle))
                        self.reasons.update(name for name, bit in self.REASON_BITS.items() if mask & bit)
                    except pynvml.NVMLError:
                        pass
                    self.stop_flag.wait(0.01)

            self.source = "nvml"
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:  # no NVML binding: the nvidia-smi child below
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _nvml_index(self):
        # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a list of indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except (ValueError, IndexError):
                pass
        return self.gpu

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join()
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.kill()
            self.proc.wait()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for ln in self.lines:
                f = [t.strip() for t in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    self.sm.append(float(f[1]))
                    self.smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        self.reasons.add(name)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.smax) if self.smax else None, "samples": len(self.sm),
                "source": self.source, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's kernels on the host cores
# ---------------------------------------------------------------------------
def cpu_scene_pass(orect, ogrid, lon, lat, bands, size, xy_min, res):
    """One scene pass with the C restatement of the reference kernels (all OpenMP threads)."""
    g = ogrid.regular_grid(size, xy_min, res, tile_size=TILE)
    n = 0
    for method in METHODS:
        windows = orect.source_windows(lon, lat, g)
        ij = orect.rectify_ij(lon, lat, g, windows=windows)
        out = orect.gather(bands, ij, method, np.nan)
        n += out.size
    return n


def run_cpu(steps, warmup, scale):
    import oracle
    from oracle import grid as ogrid
    from oracle import rectify as orect

    oracle.build()
    lon, lat, bands, size, xy_min, res = make_scene(scale=scale)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would
    # otherwise make the CPU arm single-threaded)
    try:
        n_host = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n_host = os.cpu_count() or 1
    oracle.lib().xrso_set_num_threads(n_host)
    cores = oracle.lib().xrso_num_threads()
    for _ in range(warmup):
        cpu_scene_pass(orect, ogrid, lon, lat, bands, size, xy_min, res)
    t0 = time.perf_counter()
    units = 0
    for _ in range(steps):
        units += cpu_scene_pass(orect, ogrid, lon, lat, bands, size, xy_min, res)
    dt = time.perf_counter() - t0
    h, w = lon.shape
    sample = (f"{steps} scene pass(es) of a {w}x{h} swath ({scale:g}x linear scale of the workload), "
              f"{bands.shape[0]} bands, -> {size[0]}x{size[1]}, nearest+bilinear, {dt:.1f} s")
    return units / dt / 1e6, cores, sample, dt / steps * 1e3, (w, h, bands.shape[0], size)


def reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port, kind 'port') on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scale = args.scale
    value, cores, sample, ms, _sample_dims = run_cpu(args.steps, args.warmup, scale)
    # describe the full workload (same strings as the GPU arm); the sample is named separately
    from xcube_resampling_b200 import synthetic as syn

    w, h = max(64, int(round(syn.OLCI_WIDTH * args.scale))), max(64, int(round(syn.OLCI_HEIGHT * args.scale)))
    lon, lat = syn.swath(w, h, res=syn.OLCI_RES_DEG, theta=12.0, seed=0)
    size, _ = syn.covering_grid_args(lon, lat, syn.OLCI_RES_DEG)
    nb = syn.OLCI_BANDS
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(w, h, nb, size, args.gpus) | {"bounded_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def pin_to_gpu_numa_node(gpu_index):
    """Restrict this process to the CPUs NVML reports as local to the GPU, so that page-locked host
    buffers (first touched here) sit on the NUMA node of the GPU's PCIe root.  Returns the original
    CPU set (restored before the CPU baseline runs); a no-op when NVML is unavailable."""
    try:
        original = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        return None
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (max(original) // 64) + 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        local = {64 * k + b for k, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        local &= original
        if local:
            os.sched_setaffinity(0, local)
    except Exception:  # NVML missing or not permitted: keep the inherited affinity
        pass
    return original


def ours(args):
    import torch
    import torch.distributed as dist

    import xcube_resampling_b200 as xrs
    from xcube_resampling_b200 import _dev, _lib, bands as xbands, rectify as xrect

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    all_cpus = pin_to_gpu_numa_node(local_rank)  # pinned host buffers land next to the GPU's PCIe root
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        return xbands.max_over_ranks(ms, dev)

    def sum_over_ranks(v):
        return xbands.sum_over_ranks(v, dev)

    # ---- scene + geometry (every rank builds the same synthetic scene) ----------------
    lon, lat, bands, size, xy_min, res = make_scene(scale=args.scale)
    nb, h, w = bands.shape
    target_gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=TILE)
    source_gm = xrs.GridMapping.from_coords(lon, lat, "EPSG:4326", xy_res=res, xy_dim_names=("x", "y"))
    W_t, H_t = target_gm.size
    x_dev = _dev.to_device(lon)
    y_dev = _dev.to_device(lat)
    if world == 1:
        rows = (0, H_t)
    else:
        # Row bands of equal WORK, not equal height: a rotated swath leaves the top and bottom rows of
        # the target mostly empty.  The weights come from one full ij image (valid pixels per row plus
        # a constant for the fill writes); every rank derives the same partition.
        ij_full = xrect.RectifyPlan(target_gm, dev).ij(x_dev, y_dev)
        valid_per_row = (~torch.isnan(ij_full[0])).sum(dim=1).cpu().numpy().astype(np.float64)
        del ij_full
        rows = xbands.weighted_row_bands(valid_per_row + 0.3 * W_t, world, align=32)[rank]
    band_px = (rows[1] - rows[0]) * W_t

    # ---- device residency: full coordinates + the band's source footprint of all bands ----
    plan = xrect.RectifyPlan(target_gm, dev, rows=rows)
    boxes_host = _dev.to_host(plan.windows(x_dev, y_dev))
    fp = xbands.rectify_band_footprint(boxes_host, target_gm, rows, (w, h)) or (0, 0, w, min(h, 2))
    fi0, fj0, fi1, fj1 = 0, fp[1], w, fp[3]  # full-width rows of the footprint
    src_dev = _dev.to_device_pitched(bands[:, fj0:fj1, :])  # 128-byte row pitch -> TMA-staged gather
    outs = {m: torch.empty((nb, rows[1] - rows[0], W_t), dtype=torch.float32, device=dev) for m in METHODS}
    torch.cuda.synchronize()

    phase_ms = {"k0": 0.0, "k1": 0.0, "k2_nearest": 0.0, "k2_bilinear": 0.0}
    pending = []

    # one plan (tile tables + K0/K1 workspaces) per method: the nearest and the bilinear passes of a
    # step are independent, and at N > 1 they run as two concurrent chains (see step())
    plans = {m: (plan if k == 0 else xrect.RectifyPlan(target_gm, dev, rows=rows)) for k, m in enumerate(METHODS)}

    def method_pass(m, record):
        p = plans[m]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if record:
            evs[0].record()
        boxes = p.windows(x_dev, y_dev)
        if record:
            evs[1].record()
        if args.two_step:  # xrs_rectify_ij + xrs_gather_ij (the ij image goes through HBM)
            ij = p.ij(x_dev, y_dev, boxes)
            if record:
                evs[2].record()
            xrect.gather_ij(src_dev, ij, m, np.nan, out=outs[m], window_origin=(fi0, fj0), full_size=(w, h))
        else:              # xrs_rectify_gather: one variable per call, ij resolved in registers
            if record:
                evs[2].record()
            p.rectify_gather(x_dev, y_dev, src_dev, m, np.nan, out=outs[m], tile_boxes=boxes,
                             window_origin=(fi0, fj0), full_size=(w, h))
        if record:
            evs[3].record()
            pending.append((m, evs))

    chains = [torch.cuda.Stream(dev) for _ in METHODS]

    def step(record=False, concurrent=False):
        if not concurrent:
            for _scene in range(world):
                for m in METHODS:
                    method_pass(m, record)
            return
        # N > 1: a rank's band kernels are too small to fill the GPU one at a time (10-140 us, K0 and
        # the K1 scatter latency-bound), so the two methods' passes run as two chains on two streams,
        # forked from and joined to the current stream -- same work, same buffers per chain
        main = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(main)
        for m, chain in zip(METHODS, chains):
            chain.wait_event(fork)
            with torch.cuda.stream(chain):
                for _scene in range(world):
                    method_pass(m, False)
                joined = torch.cuda.Event()
                joined.record(chain)
            main.wait_event(joined)

    # ---- end to end through the public API with host buffers --------------------------
    # (This leg runs before the device-resident one, so that neither the clock sampler nor the
    # device-resident buffers' allocation is in flight while it is timed.)
    e2e = None
    if not args.no_e2e:
        # page-locked host inputs; this rank's band of the target grid as its own GridMapping
        lon_p, lat_p = _dev.pinned_empty(lon.shape, np.float64), _dev.pinned_empty(lat.shape, np.float64)
        lon_p[...] = lon
        lat_p[...] = lat
        bands_p = _dev.pinned_empty((nb, fj1 - fj0, w), np.float32)
        bands_p[...] = bands[:, fj0:fj1, :]
        e2e_steps = max(1, min(args.steps, 10))

        ds = xrs.Dataset(data_vars=dict(bands=(("band", "y", "x"), bands_p)),
                         coords=dict(lon=(("y", "x"), lon_p), lat=(("y", "x"), lat_p)))
        # the grid mapping of the end-to-end leg wraps the page-locked coordinate arrays (it is what
        # rectify_dataset uploads), so every host buffer of the call is pinned
        source_gm = xrs.GridMapping.from_coords(lon_p, lat_p, "EPSG:4326", xy_res=res, xy_dim_names=("x", "y"))

        def e2e_step():
            n = 0
            for _scene in range(world):
                for m in METHODS:
                    if world == 1:  # the call a user makes
                        out = xrs.rectify_dataset(ds, target_gm=target_gm, source_gm=source_gm,
                                                  interp_methods=m)["bands"].values
                    else:           # the same device pipeline on this rank's row band
                        out = xrect.rectify_band_host(lon_p, lat_p, bands_p, (fi0, fj0), (w, h), target_gm, rows,
                                                      m, np.nan)
                    n += out.size
            return n

        # Warm-up to steady state.  The first step page-locks the output buffers (seconds); for two to
        # three seconds after that, single steps were measured to take 1.5-4x longer at random on
        # this pool (host-side: the kernels and the copies of such a step are not slower when timed
        # alone).  Warm-up therefore runs for at least 4 s after the first step AND until three
        # consecutive steps agree within 3 % on every rank (at most 60 steps or 25 s).  The timed steps that
        # follow are consecutive and all counted; every warm-up and timed step time is reported.
        warm_ms = []
        t_first = None
        while len(warm_ms) < 60:
            ts = time.perf_counter()
            e2e_step()
            warm_ms.append(round((time.perf_counter() - ts) * 1e3, 2))
            if t_first is None:
                t_first = time.perf_counter()
            last = warm_ms[-3:]
            waited = time.perf_counter() - t_first
            steady = (len(warm_ms) >= 4 and max(last) <= 1.03 * min(last) and waited >= 4.0) or waited >= 25.0
            if max_over_ranks(0.0 if steady else 1.0) == 0.0:
                break
        barrier()
        t0 = time.perf_counter()
        n_units = 0
        step_ms = []
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            n_units += e2e_step()
            step_ms.append(round((time.perf_counter() - ts) * 1e3, 2))
        torch.cuda.synchronize()
        dt_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        n_units = sum_over_ranks(float(n_units))
        # bytes copied per step by the whole job: every rank handles `world` scenes x 2 methods
        h2d = sum_over_ranks(float(world * len(METHODS) * (lon_p.nbytes + lat_p.nbytes + bands_p.nbytes)))
        d2h = sum_over_ranks(float(world * len(METHODS) * nb * band_px * 4))
        e2e = {"value": n_units / (dt_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": dt_ms / e2e_steps,
               "rank0_step_ms": step_ms, "rank0_warmup_step_ms": warm_ms,
               "api": ("rectify_dataset(ds, target_gm, source_gm, interp_methods)" if world == 1 else
                       "rectify_band_host (rectify_dataset's device pipeline on this rank's row band)")
                      + ": pinned host arrays in, pinned host arrays out, host clock around synchronised calls"}

    # ---- device-resident timing ------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.profile_collect()
    use_graph = not args.no_graph
    # Region A (every N): K eager steps, one launch after the other on one stream, with libxrs's
    # per-launch CUDA events -- the per-kernel durations behind `roofline` and the step time of the
    # plain sequential form.
    launches0 = lib.xrs_launch_count()
    _lib.profile_enable(True)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(args.steps):
        step(record=True)
    a1.record()
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    kernel_times = _lib.profile_collect()
    launches = lib.xrs_launch_count() - launches0
    eager_ms = max_over_ranks(a0.elapsed_time(a1))
    e0, e1 = a0, a1
    if use_graph:
        # Region B, the one `value` is computed from: the same step captured once in a CUDA graph --
        # the nearest and the bilinear passes as two chains on two streams (K0 and the K1 scatter are
        # latency-bound and leave bandwidth for the other chain's gather; at N > 1 a rank's band
        # kernels are only 10-140 us) -- and replayed K times.
        launches0 = lib.xrs_launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step(concurrent=True)
        launches = (lib.xrs_launch_count() - launches0) * args.steps
        graph.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
    my_ms = e0.elapsed_time(e1)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(my_ms)
    units_rank = args.steps * world * len(METHODS) * nb * band_px
    units = sum_over_ranks(float(units_rank))
    value = units / (total_ms * 1e-3) / 1e6
    for m, evs in pending:
        phase_ms["k0"] += evs[0].elapsed_time(evs[1])
        phase_ms["k1"] += evs[1].elapsed_time(evs[2])
        phase_ms["k2_" + m] += evs[2].elapsed_time(evs[3])
    n_pass = args.steps * world
    phase_ms = {k: v / n_pass / (len(METHODS) if k in ("k0", "k1") else 1) for k, v in phase_ms.items()}

    # ---- roofline of the dominant kernel (largest share of the timed region) ------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    S = float(h * w)                      # source pixels (coordinates are resident in full)
    s_used = float((fj1 - fj0) * w)       # source pixels of the data bands resident for this row band
    T = float(band_px)
    # algorithmic (compulsory) bytes per launch, DESIGN.md "Kernels"
    if args.two_step:
        k2_index_bytes, k2_model = 16.0 * T, "16*T (ij) + 4*B*S_used (source once) + 4*B*T (output once)"
    else:
        k2_index_bytes = 4.0 * T + 16.0 * S
        k2_model = "4*T (claims) + 16*S (winning quads' vertices) + 4*B*S_used (source once) + 4*B*T (output once)"
    models = {
        "k0_tile_windows": (16.0 * S, "16*S (lon+lat fp64 read once)"),
        "k1_init_claims": (4.0 * T, "4*T (claim word per target pixel)"),
        "k1_scatter": (16.0 * S + 4.0 * T, "16*S (lon+lat fp64 read once) + 4*T (claim word per target pixel)"),
        "k1_scatter_slow": (0.0, "queue of border quads, no compulsory traffic of its own"),
        "k1_resolve": (4.0 * T + 16.0 * T + 16.0 * S, "4*T (claims) + 16*T (ij fp64 out) + 16*S (winning quads' vertices)"),
        "k2_gather_staged<nearest>": (k2_index_bytes + 4.0 * nb * s_used + 4.0 * nb * T, k2_model),
        "k2_gather_staged<bilinear>": (k2_index_bytes + 4.0 * nb * s_used + 4.0 * nb * T, k2_model),
    }
    total_kernel_ms = sum(v[0] for v in kernel_times.values()) or 1.0
    kernels = []
    for name, (ms_total, n_launch) in sorted(kernel_times.items(), key=lambda kv: -kv[1][0]):
        ms_launch = ms_total / max(n_launch, 1)
        nbytes, model = models.get(name, (0.0, "n/a"))
        gbs = nbytes / (ms_launch * 1e-3) / 1e9 if ms_launch > 0 else 0.0
        kernels.append({"kernel": name, "launches": int(n_launch), "ms_per_launch": ms_launch,
                        "share_of_kernel_time": ms_total / total_kernel_ms, "algorithmic_bytes_per_launch": nbytes,
                        "achieved_gbs": gbs, "frac": gbs / peak, "bytes_model": model})
    top = kernels[0] if kernels else {"kernel": "none", "ms_per_launch": 0.0, "achieved_gbs": 0.0, "frac": 0.0,
                                      "algorithmic_bytes_per_launch": 0.0, "bytes_model": "n/a",
                                      "share_of_kernel_time": 0.0}
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures of this
    # exact workload (profiles/r01_final_k0_k1_ncu_full.txt, r01_final_k2_fused_ncu_full.txt,
    # r01_final_k2_two_step_ncu_full.txt)
    if args.two_step:
        k2_traffic = {"k2_gather_staged<bilinear>": 5673.4e6, "k2_gather_staged<nearest>": 5692.7e6}
    else:
        k2_traffic = {"k2_gather_staged<bilinear>": 5520.4e6, "k2_gather_staged<nearest>": 5522.9e6}
    traffic = ({"k1_scatter": 525.0e6, "k0_tile_windows": 324.3e6, "k1_resolve": 1063.8e6} | k2_traffic).get(top["kernel"]) \
        if (world == 1 and args.scale == 1.0) else None
    roofline = {
        "bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_gbs"], "peak": peak,
        "peak_kind": peak_kind, "unit": "GB/s", "frac": top["frac"],
        "traffic": traffic,
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture of this "
                          "workload (profiles/r01_final_*_ncu_full.txt)" if traffic else None,
        "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"], "bytes_model": top["bytes_model"],
        "ms_per_launch": top["ms_per_launch"], "share_of_kernel_time": top["share_of_kernel_time"],
        "timing": "CUDA events recorded by libxrs around every launch on the launching stream over K eager, "
                  "sequential steps (ms_per_step_eager)" +
                  ("; `value` is timed over K CUDA-graph replays of the same step with the nearest and bilinear "
                   "passes as two concurrent chains" if use_graph else "; `value` is that region"),
        "ms_per_step_eager": eager_ms / args.steps,
        "kernels": kernels, "phase_ms_per_rectify": phase_ms,
    }

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if all_cpus:
            os.sched_setaffinity(0, all_cpus)  # the CPU arm uses every host core again
        v, cores, sample, _ms, _ = run_cpu(steps=5, warmup=1, scale=args.scale)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(w, h, nb, size, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Kernel timings at the BASELINE.json config sizes other than the bench.py headline (C2).

    python tools/bench_configs.py [c1] [c3] [c4] [c5] [--reps 5] [--one-shot] [--no-cpu]

For each configuration the device-resident kernel time (CUDA events on the launching stream,
median of ``reps`` after 2 warm-ups), output Mpix*band/s and the algorithmic-bytes bandwidth
(SURVEY.md 8d byte models) against MEASURED_PEAKS.json, plus a CPU figure for the same
computation on a bounded sample (oracle: scipy / numpy, what the reference calls; for the
projections a vectorised numpy restatement of the PROJ formulas -- not reference code).
``bench.py`` imports :func:`run_all` and puts the lines into its ``configs`` array (N=1).
``--one-shot`` runs every kernel exactly once (for ncu captures).
"""

import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import xcube_resampling_b200 as xrs  # noqa: E402
from xcube_resampling_b200 import _dev, affine, reproject  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        return 6650.0


def timed(fn, reps, one_shot):
    if one_shot:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def line(name, ms, units, bytes_, extra=None, cpu=None):
    ok = ms == ms and ms > 0
    out = {"config": name, "ms": ms, "Mpix_band_per_s": units / ms / 1e3 if ok else None,
           "algorithmic_bytes": bytes_, "GB_per_s": bytes_ / ms / 1e6 if ok else None,
           "frac": bytes_ / ms / 1e6 / peak() if ok else None, "cpu": cpu}
    out.update(extra or {})
    return out


def rand_dev(shape, dtype=torch.float32, seed=0, pitch_bytes=None):
    """Random device array; ``pitch_bytes`` pads the rows (a view of the padded buffer is returned) --
    a 128-byte pitch is what the entry points' upload pipeline produces and what the TMA-staged
    kernels need."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    w = shape[-1]
    if pitch_bytes:
        per = max(1, pitch_bytes // torch.empty((), dtype=dtype).element_size())
        shape = tuple(shape[:-1]) + (-(-w // per) * per,)
    if dtype == torch.uint8:
        t = torch.randint(0, 20, shape, dtype=torch.uint8, device="cuda", generator=g)
    else:
        t = torch.rand(shape, dtype=dtype, device="cuda", generator=g)
    return t[..., :w]


def cpu_timed(fn, units, threads, sample):
    """Mpix*band/s of ``fn()`` (one bounded sample processing ``units`` output pixels*bands)."""
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    return {"Mpix_band_per_s": units / dt / 1e6, "cores": threads, "sample": sample, "seconds": dt}


# ---------------------------------------------------------------------------
# C1 / C4: affine + coarsen
# ---------------------------------------------------------------------------
def c1(reps, one_shot, cpu, threads):
    """affine_transform_dataset: 2x bilinear downsample of 4096^2 float32 (identity gather + 2x2 mean)."""
    src = rand_dev((4096, 4096))
    fn = lambda: affine.affine_resample_dev(src, (1.0, 1.0), (0.0, 0.0), (2048, 2048), 1, float("nan"), "mean", (2, 2))  # noqa: E731
    ms = timed(fn, reps, one_shot)
    cpu_fig = None
    if cpu:
        from oracle import resample as ores

        from oracle import grid as ogrid

        a = np.random.default_rng(0).random((4096, 4096)).astype(np.float32)
        sg = ogrid.regular_grid((1024, 1024), (0, 0), 0.01)
        tg = ogrid.regular_grid((512, 512), (0, 0), 0.02)

        def chunk(k):  # one 1024^2 source chunk -> 512^2 output chunk (the reference's dask chunks)
            j, i = divmod(k, 4)
            return ores.affine_transform(a[j * 1024:(j + 1) * 1024, i * 1024:(i + 1) * 1024], sg, tg, interp=1)

        def run():
            with ThreadPoolExecutor(threads) as pool:
                list(pool.map(chunk, range(16)))

        cpu_fig = cpu_timed(run, 2048 * 2048, threads, "whole image, 16 chunks of 1024^2 (scipy affine_transform + nanmean)")
    return [line("C1 affine 2x bilinear downsample 4096^2 f32", ms, 2048 * 2048, 4096 * 4096 * 4 + 2048 * 2048 * 4,
                 cpu=cpu_fig)]


def c4(reps, one_shot, cpu, threads):
    """coarsen 20000^2 float32 and uint8 by 4 / 8."""
    n = 20000
    f32 = rand_dev((n, n))
    u8 = rand_dev((n, n), torch.uint8)
    out = []
    strip = 1600  # CPU sample: rows of the raster (divisible by 4, 8)
    host = {}
    if cpu:
        host["f32"] = f32[:strip].cpu().numpy()
        host["u8"] = u8[:strip].cpu().numpy()

    def cpu_fig(kind, f, agg):
        if not cpu:
            return None
        from oracle import resample as ores

        a = host[kind]
        parts = np.array_split(np.arange(strip // f), threads)

        def part(rows):
            if len(rows):
                ores.coarsen(a[rows[0] * f:(rows[-1] + 1) * f], f, f, agg)

        def run():
            with ThreadPoolExecutor(threads) as pool:
                list(pool.map(part, parts))

        return cpu_timed(run, (strip // f) * (n // f), threads, f"first {strip} rows (numpy reducer of coarsen.py per row strip)")

    for f in (4, 8):
        for agg in ("mean", "min", "max", "median"):
            ms = timed(lambda: affine.coarsen_dev(f32, (f, f), agg), reps, one_shot)
            out.append(line(f"C4 coarsen f32 20000^2 /{f} {agg}", ms, (n // f) ** 2, n * n * 4 + (n // f) ** 2 * 4,
                            cpu=cpu_fig("f32", f, agg)))
        for agg, out_b in (("mode", 8), ("min", 1), ("max", 1)):
            ms = timed(lambda: affine.coarsen_dev(u8, (f, f), agg), reps, one_shot)
            out.append(line(f"C4 coarsen u8 20000^2 /{f} {agg}", ms, (n // f) ** 2, n * n + (n // f) ** 2 * out_b,
                            cpu=cpu_fig("u8", f, agg)))
    return out


# ---------------------------------------------------------------------------
# C3 / C5: reproject
# ---------------------------------------------------------------------------
def _cpu_reproject_block(src_gm, tgt_gm, src_epsg, tgt_epsg, bands, rows, cols, threads, method="bilinear"):
    """Oracle (numpy PROJ-formula restatement + _reproject_block arithmetic) on target rows x cols of
    one reference tile, split over ``threads`` row strips."""
    from oracle import grid as ogrid
    from oracle import proj as oproj
    from oracle import reproject as orep

    g = ogrid.regular_grid(tgt_gm.size, (tgt_gm.x_min, tgt_gm.y_min), tgt_gm.xy_res, tile_size=tgt_gm.tile_size,
                           is_j_axis_up=tgt_gm.is_j_axis_up)
    tp, sp = oproj.from_epsg(tgt_epsg), oproj.from_epsg(src_epsg)
    xs, ys = src_gm.x_values, src_gm.y_values
    win = orep.source_windows(float(xs[0]), float(ys[0]), src_gm.x_res, src_gm.y_res, float(ys[1] - ys[0]),
                              src_gm.width, src_gm.height, g, tp, sp)
    ty, tx = rows[0] // g.tile_h, cols[0] // g.tile_w
    window = np.random.default_rng(0).random((bands, win["win_h"], win["win_w"])).astype(np.float32)
    xc, yc = ogrid.x_centres(g)[cols[0]:cols[1]], ogrid.y_centres(g)[rows[0]:rows[1]]
    parts = [p for p in np.array_split(np.arange(len(yc)), threads) if len(p)]

    def part(idx):
        xx, yy = np.meshgrid(xc, yc[idx])
        xx, yy = oproj.transform(tp, sp, xx, yy)
        orep.sample_window(xx, yy, window, win["x0"][ty, tx], win["y0"][ty, tx], src_gm.x_res, src_gm.y_res, method)

    def run():
        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(part, parts))

    return run, bands * (rows[1] - rows[0]) * (cols[1] - cols[0])


def _reproject_case(name, src_gm, tgt_gm, bands, reps, one_shot, cpu, threads, epsgs, methods=("bilinear", "nearest")):
    plan = reproject.ReprojectPlan(src_gm, tgt_gm)
    fp = plan.footprint()
    s_fp = (fp[2] - fp[0]) * (fp[3] - fp[1])
    src = rand_dev((bands, src_gm.height, src_gm.width), pitch_bytes=128)
    T = tgt_gm.width * tgt_gm.height
    out = []
    for method in methods:
        for od in ((None, np.float32) if method == "bilinear" else (None,)):
            out_b = 8 if (method == "bilinear" and od is None) else 4
            dst = _dev.empty((bands, tgt_gm.height, tgt_gm.width), np.float64 if out_b == 8 else np.float32)
            ms = timed(lambda: plan.run(src, method, float("nan"), out=dst, out_dtype=od), reps, one_shot)
            cpu_fig = None
            if cpu and od is None:
                th = tgt_gm.tile_height
                run, units = _cpu_reproject_block(src_gm, tgt_gm, epsgs[0], epsgs[1], bands, (th, th + 512), (th, th + 1024),
                                                  threads, method)
                cpu_fig = cpu_timed(run, units, threads, "512 x 1024 target pixels of one reference tile, all bands "
                                                         "(numpy restatement of the PROJ formulas + _reproject_block)")
            out.append(line(f"{name} {method} out={'f64' if out_b == 8 else 'f32'}", ms, bands * T,
                            4.0 * bands * s_fp + out_b * bands * T,
                            {"target": [tgt_gm.width, tgt_gm.height], "source": [src_gm.width, src_gm.height],
                             "bands": bands, "tile_window": [plan.windows.win_w, plan.windows.win_h]}, cpu=cpu_fig))
            del dst
    return out


def c3_grids():
    tgt = xrs.GridMapping.regular((10980, 10980), (399960.0, 990240.0), 10.0, "EPSG:32632", tile_size=2048)
    box = reproject.transform_bounds("EPSG:32632", "EPSG:4326", [tgt.xy_bbox])[0]
    res = 0.0001
    x_min = float(np.floor(box[0] / res) * res) - 4 * res
    y_min = float(np.floor(box[1] / res) * res) - 4 * res
    w = int(np.ceil((box[2] - x_min) / res)) + 4
    h = int(np.ceil((box[3] - y_min) / res)) + 4
    return xrs.GridMapping.regular((w, h), (x_min, y_min), res, "EPSG:4326"), tgt


def c3(reps, one_shot, cpu, threads):
    """reproject 0.0001-deg EPSG:4326 -> UTM 32N 10980^2 Sentinel-2 tile (low latitude), 13 bands."""
    src, tgt = c3_grids()
    return _reproject_case("C3 reproject 4326->UTM32N 10980^2 13 bands", src, tgt, 13, reps, one_shot, cpu, threads,
                           (4326, 32632))


def c5_grids():
    ext = 20037508.342789244
    tgt = xrs.GridMapping.regular((36000, 36000), (-ext, -ext), 2 * ext / 36000, "EPSG:3857", tile_size=4500)
    src = xrs.GridMapping.regular((36000, 18000), (-180.0, -90.0), 0.01, "EPSG:4326")
    return src, tgt


def c5(reps, one_shot, cpu, threads):
    """global 0.01-deg grid -> EPSG:3857, one of the 8 row bands (4500 target rows) of 8 variables."""
    src, tgt = c5_grids()
    windows = reproject.get_source_windows(src, tgt)
    out = []
    for band_index in (3, 0):  # an equatorial band and the northernmost one
        rows = (band_index * 4500, (band_index + 1) * 4500)
        plan = reproject.ReprojectPlan(src, tgt, rows=rows, windows=windows)
        i0, j0, i1, j1 = plan.footprint()
        nb = 8
        window = rand_dev((nb, j1 - j0, i1 - i0), pitch_bytes=128)
        dst = _dev.empty((nb, 4500, 36000), np.float32)
        ms = timed(lambda: plan.run(window, "bilinear", float("nan"), out=dst, out_dtype=np.float32,
                                    window_origin=(i0, j0)), reps, one_shot)
        cpu_fig = None
        if cpu and band_index == 3:
            run, units = _cpu_reproject_block(src, tgt, 4326, 3857, nb, (rows[0], rows[0] + 512), (4500, 4500 + 1024), threads)
            cpu_fig = cpu_timed(run, units, threads, "512 x 1024 target pixels of one reference tile, 8 variables "
                                                     "(numpy restatement of the PROJ formulas + _reproject_block)")
        out.append(line(f"C5 reproject global 0.01deg -> 3857, row band {band_index} of 8, 8 vars bilinear out=f32", ms,
                        nb * 4500 * 36000, 4.0 * nb * (j1 - j0) * (i1 - i0) + 4.0 * nb * 4500 * 36000,
                        {"footprint_rows": [j0, j1], "footprint_cols": [i0, i1]}, cpu=cpu_fig))
        del window, dst
    return out


def c5_across(rank, world, reps=3, e2e=True):
    """BASELINE.json configs[4]: the global 0.01-deg stack (8 float32 variables) reprojected to EPSG:3857,
    36000^2, partitioned over ``world`` GPUs by target row bands of 4500 rows (the reference tile rows):
    rank r computes bands r, r + world, ...  Returns this rank's figures (the caller reduces over
    ranks): device-resident kernel time per band and, with ``e2e``, the same bands through
    ``reproject_groups`` from page-locked host arrays (footprint-only upload, band download)."""
    from xcube_resampling_b200._pipeline import Target, group_by_buffer

    src, tgt = c5_grids()
    windows = reproject.get_source_windows(src, tgt)
    nb, n_bands = 8, 8
    mine = [b for b in range(n_bands) if b % world == rank]
    out = {"bands": mine, "kernel_ms": 0.0, "units": 0, "algorithmic_bytes": 0.0, "per_band_ms": [],
           "e2e_s": None, "h2d_bytes": 0, "d2h_bytes": 0}
    for b in mine:
        rows = (b * 4500, (b + 1) * 4500)
        plan = reproject.ReprojectPlan(src, tgt, rows=rows, windows=windows)
        i0, j0, i1, j1 = plan.footprint()
        window = rand_dev((nb, j1 - j0, i1 - i0), pitch_bytes=128)
        dst = _dev.empty((nb, 4500, 36000), np.float32)
        ms = timed(lambda: plan.run(window, "bilinear", float("nan"), out=dst, out_dtype=np.float32,
                                    window_origin=(i0, j0)), reps, False)
        out["kernel_ms"] += ms
        out["per_band_ms"].append(ms)
        out["units"] += nb * 4500 * 36000
        out["algorithmic_bytes"] += 4.0 * nb * (j1 - j0) * (i1 - i0) + 4.0 * nb * 4500 * 36000
        del window, dst
        torch.cuda.empty_cache()
    if e2e and mine:
        # one page-locked source plane shared by the 8 variables (zero band stride: the bytes that move
        # are real, the host memory stays at 2.6 GB per rank), band-only page-locked outputs
        plane = _dev.pinned_empty((src.height, src.width), np.float32)
        rng = np.random.default_rng(rank)
        for j in range(0, src.height, 1000):
            plane[j:j + 1000] = rng.random((min(1000, src.height - j), src.width), dtype=np.float32)
        values = np.lib.stride_tricks.as_strided(plane, shape=(nb,) + plane.shape, strides=(0,) + plane.strides)
        host_out = _dev.pinned_empty((nb, 4500, 36000), np.float32)
        dev = torch.device("cuda", torch.cuda.current_device())
        stats = []

        def run_all_bands():
            for b in mine:
                groups = group_by_buffer([(values, Target("v", "bilinear", float("nan"), host_out, row0=b * 4500))])
                reproject.reproject_groups(groups, src, tgt, [dev], band_edges=[b * 4500, (b + 1) * 4500],
                                           windows=windows, stats=stats)

        run_all_bands()  # warm-up (page-locking, allocator)
        stats.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_all_bands()
        out["e2e_s"] = time.perf_counter() - t0
        out["h2d_bytes"] = sum(st["h2d_bytes"] for st in stats)
        out["d2h_bytes"] = sum(st["d2h_bytes"] for st in stats)
    return out


CASES = {"c1": c1, "c3": c3, "c4": c4, "c5": c5}


def run_all(reps=3, cpu=True, cpu_threads=None, restore_affinity=None, names=("c1", "c3", "c4", "c5"), one_shot=False):
    """Lines of all configurations (list of dicts).  ``restore_affinity``: CPU set to run the CPU
    figures on (bench.py pins itself to the GPU's NUMA node while timing the GPU legs)."""
    threads = cpu_threads or len(os.sched_getaffinity(0))
    if restore_affinity:
        try:
            os.sched_setaffinity(0, restore_affinity)
        except OSError:
            pass
    out = []
    for name in names:
        try:
            out.extend(CASES[name](reps, one_shot, cpu, threads))
        except Exception as e:  # one configuration failing must not take the bench line down
            out.append({"config": name, "error": f"{type(e).__name__}: {e}"})
        torch.cuda.empty_cache()
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 5
    if "--reps" in sys.argv:
        args = [a for a in args if a != sys.argv[sys.argv.index("--reps") + 1]]
    torch.cuda.set_device(0)
    for ln in run_all(reps=reps, cpu="--no-cpu" not in sys.argv, names=args or ("c1", "c3", "c4", "c5"),
                      one_shot="--one-shot" in sys.argv):
        print(json.dumps(ln), flush=True)


if __name__ == "__main__":
    main()

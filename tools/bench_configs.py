#!/usr/bin/env python
"""Kernel timings at the BASELINE.json config sizes other than the bench.py headline (C2).

    python tools/bench_configs.py [c1] [c3] [c4] [c5] [--reps 5] [--one-shot]

For each configuration the device-resident kernel time (CUDA events on the launching stream,
median of ``reps`` after 2 warm-ups), output Mpix*band/s and the algorithmic-bytes bandwidth
(SURVEY.md 8d byte models) against MEASURED_PEAKS.json are printed as one JSON line each.
``--one-shot`` runs every kernel exactly once (for ncu captures).  Not a driver contract -- bench.py
is; this feeds DESIGN.md / profiles/.
"""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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


def report(name, ms, units, bytes_, extra=None):
    line = {"config": name, "ms": ms, "Mpix_band_per_s": units / ms / 1e3 if ms == ms else None,
            "algorithmic_GB": bytes_ / 1e9, "GB_per_s": bytes_ / ms / 1e6 if ms == ms else None,
            "frac_of_measured_peak": bytes_ / ms / 1e6 / peak() if ms == ms else None}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def rand_dev(shape, dtype=torch.float32, seed=0):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    if dtype == torch.uint8:
        return torch.randint(0, 20, shape, dtype=torch.uint8, device="cuda", generator=g)
    return torch.rand(shape, dtype=dtype, device="cuda", generator=g)


def c1(reps, one_shot):
    """affine_transform_dataset: 2x bilinear downsample of 4096^2 float32 (identity gather + 2x2 mean)."""
    src = rand_dev((4096, 4096))
    fn = lambda: affine.affine_resample_dev(src, (1.0, 1.0), (0.0, 0.0), (2048, 2048), 1, float("nan"), "mean", (2, 2))  # noqa: E731
    ms = timed(fn, reps, one_shot)
    report("C1 affine 2x bilinear downsample 4096^2 f32", ms, 2048 * 2048, 4096 * 4096 * 4 + 2048 * 2048 * 4)


def c4(reps, one_shot):
    """coarsen 20000^2 float32 and uint8 by 4 / 8."""
    n = 20000
    f32 = rand_dev((n, n))
    u8 = rand_dev((n, n), torch.uint8)
    for f in (4, 8):
        for agg in ("mean", "min", "max", "median"):
            ms = timed(lambda: affine.coarsen_dev(f32, (f, f), agg), reps, one_shot)
            report(f"C4 coarsen f32 20000^2 /{f} {agg}", ms, (n // f) ** 2, n * n * 4 + (n // f) ** 2 * 4)
        for agg, out_b in (("mode", 8), ("min", 1), ("max", 1)):
            ms = timed(lambda: affine.coarsen_dev(u8, (f, f), agg), reps, one_shot)
            report(f"C4 coarsen u8 20000^2 /{f} {agg}", ms, (n // f) ** 2, n * n + (n // f) ** 2 * out_b)


def _reproject_case(name, src_gm, tgt_gm, bands, reps, one_shot, methods=("bilinear", "nearest"), out_dtype=None):
    plan = reproject.ReprojectPlan(src_gm, tgt_gm)
    fp = plan.footprint()
    s_fp = (fp[2] - fp[0]) * (fp[3] - fp[1])
    src = rand_dev((bands, src_gm.height, src_gm.width))
    T = tgt_gm.width * tgt_gm.height
    for method in methods:
        for od in ((None, np.float32) if method == "bilinear" else (None,)):
            out_b = 8 if (method == "bilinear" and od is None) else 4
            out = _dev.empty((bands, tgt_gm.height, tgt_gm.width), np.float64 if out_b == 8 else np.float32)
            ms = timed(lambda: plan.run(src, method, float("nan"), out=out, out_dtype=od), reps, one_shot)
            report(f"{name} {method} out={'f64' if out_b == 8 else 'f32'}", ms, bands * T,
                   4.0 * bands * s_fp + out_b * bands * T,
                   {"target": [tgt_gm.width, tgt_gm.height], "source": [src_gm.width, src_gm.height], "bands": bands,
                    "tile_window": [plan.windows.win_w, plan.windows.win_h]})
            del out


def c3(reps, one_shot):
    """reproject 0.0001-deg EPSG:4326 -> UTM 32N 10980^2 Sentinel-2 tile (low latitude), 13 bands."""
    tgt = xrs.GridMapping.regular((10980, 10980), (399960.0, 990240.0), 10.0, "EPSG:32632", tile_size=2048)
    box = reproject.transform_bounds("EPSG:32632", "EPSG:4326", [tgt.xy_bbox])[0]
    res = 0.0001
    x_min = float(np.floor(box[0] / res) * res) - 4 * res
    y_min = float(np.floor(box[1] / res) * res) - 4 * res
    w = int(np.ceil((box[2] - x_min) / res)) + 4
    h = int(np.ceil((box[3] - y_min) / res)) + 4
    src = xrs.GridMapping.regular((w, h), (x_min, y_min), res, "EPSG:4326")
    _reproject_case("C3 reproject 4326->UTM32N 10980^2 13 bands", src, tgt, 13, reps, one_shot)


def c5(reps, one_shot):
    """global 0.01-deg grid -> EPSG:3857, one of the 8 row bands (4500 target rows) of 8 variables."""
    ext = 20037508.342789244
    tgt = xrs.GridMapping.regular((36000, 36000), (-ext, -ext), 2 * ext / 36000, "EPSG:3857", tile_size=4500)
    src = xrs.GridMapping.regular((36000, 18000), (-180.0, -90.0), 0.01, "EPSG:4326")
    windows = reproject.get_source_windows(src, tgt)
    for band_index in (3, 0):  # an equatorial band and the northernmost one
        rows = (band_index * 4500, (band_index + 1) * 4500)
        plan = reproject.ReprojectPlan(src, tgt, rows=rows, windows=windows)
        i0, j0, i1, j1 = plan.footprint()
        nb = 8
        window = rand_dev((nb, j1 - j0, i1 - i0))
        out = _dev.empty((nb, 4500, 36000), np.float32)
        ms = timed(lambda: plan.run(window, "bilinear", float("nan"), out=out, out_dtype=np.float32,
                                    window_origin=(i0, j0)), reps, one_shot)
        report(f"C5 reproject global 0.01deg -> 3857, row band {band_index} of 8, 8 vars bilinear out=f32", ms,
               nb * 4500 * 36000, 4.0 * nb * (j1 - j0) * (i1 - i0) + 4.0 * nb * 4500 * 36000,
               {"footprint_rows": [j0, j1], "footprint_cols": [i0, i1]})
        del window, out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 5
    one_shot = "--one-shot" in sys.argv
    todo = args or ["c1", "c3", "c4", "c5"]
    torch.cuda.set_device(0)
    for name in todo:
        {"c1": c1, "c3": c3, "c4": c4, "c5": c5}[name](reps, one_shot)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

"""GPU tests of the multi-GPU product path (``devices=[...]``), run on ONE device: the participants
are host threads that may share a GPU, so ``devices=[0, 0, 0]`` exercises the slab scans, the table
exchange, the ragged footprint uploads, K1 restricted to a footprint and the band downloads into one
output array.  Everything is compared bit for bit with the single-device path and the oracle.
"""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import rectify as orect

from .helpers import INT32_MAX, assert_same, covering_grid_args, quad_footprints_np, swath

pytestmark = pytest.mark.gpu
nan = np.nan


@pytest.fixture(scope="module")
def xrs():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import xcube_resampling_b200 as pkg
    from xcube_resampling_b200 import _dev, multigpu, rectify, reproject

    pkg.dev, pkg.rect, pkg.rep, pkg.mg = _dev, rectify, reproject, multigpu
    return pkg


def _scene(w, h, theta, seed, res=0.0027, tile=128, holes=True):
    x, y = swath(w, h, theta=theta, seed=seed)
    if holes:
        x[h // 3:h // 3 + 3, w // 4:w // 2] = nan
        y[h // 2, w // 2] = np.inf
    size, xy_min = covering_grid_args(x, y, res)
    return x, y, size, xy_min, res, tile


def test_slab_scans_merge_to_the_whole_swath_tables(xrs):
    """K0 over row slabs + MIN merge == K0 over the whole swath; the footprint kernel == its numpy
    restatement; both for an uneven number of slabs."""
    import torch

    x, y, size, xy_min, res, tile = _scene(411, 333, 21.0, 3)
    h, w = x.shape
    gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=tile)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile)
    xd, yd = xrs.dev.to_device(x), xrs.dev.to_device(y)
    plan = xrs.rect.RectifyPlan(gm, xd.device)
    want_boxes = xrs.dev.to_host(plan.windows(xd, yd)).copy()
    assert_same(want_boxes, orect.source_windows(x, y, g), "K0 whole swath")
    lib = plan.lib
    group = int(lib.xrs_quad_row_group())
    n_groups = -(-(h - 1) // group)
    n_tiles = plan.ntx * plan.nty
    for n in (1, 3, 5):
        edges = xrs.mg.default_band_edges(gm.height, n)
        parts = []
        for (s0, s1) in xrs.mg.source_slabs(h, n, group):
            table = xrs.dev.empty((4 * n_tiles + 2 * n * n_groups,), np.int32)
            assert lib.xrs_minform_init(xrs.dev.ptr(table), table.numel(), xrs.dev.stream_ptr()) == 0
            if s1 > s0:
                s1v = min(h, s1 + 1)
                plan.scan_slab(xd[s0:s1v], yd[s0:s1v], s0, s1 - s0, h, w, edges, table)
            parts.append(xrs.dev.to_host(table))
        merged = xrs.mg.merge_minform_host(parts)
        boxes = xrs.dev.to_host(plan.finalize_windows(xrs.dev.to_device(merged), w, h))
        assert_same(boxes, want_boxes, f"K0 merged from {n} slabs")
        fp = merged[4 * n_tiles:].reshape(n, n_groups, 2)
        assert_same(fp, quad_footprints_np(x, y, g, edges, group), f"quad footprints, {n} bands")
        assert (fp[:, :, 0] != INT32_MAX).any()
    torch.cuda.synchronize()


@pytest.mark.parametrize("theta,n_dev,dtype", [(12.0, 3, np.float32), (-35.0, 2, np.uint8), (77.0, 4, np.int16),
                                               (0.0, 5, np.float64)])
def test_rectify_dataset_on_several_devices_equals_one(xrs, theta, n_dev, dtype):
    """rectify_dataset(devices=[0]*n) fills one output array band by band from footprint-only uploads:
    bit-identical to the single-device call and to the oracle."""
    x, y, size, xy_min, res, tile = _scene(380, 300, theta, 7)
    h, w = x.shape
    rng = np.random.default_rng(1)
    data = (rng.random((6, h, w)) * 200).astype(dtype)
    single = (rng.random((h, w)) * 100).astype(np.float32)
    ds = xrs.Dataset(data_vars=dict(a_nearest=(("band", "y", "x"), data), a_bilinear=(("band", "y", "x"), data),
                                    flat=(("y", "x"), single)),
                     coords=dict(lon=(("y", "x"), x), lat=(("y", "x"), y)))
    src_gm = xrs.GridMapping.from_coords(x, y, "EPSG:4326", xy_res=res, xy_dim_names=("x", "y"))
    tgt_gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=tile)
    interp = {"a_nearest": "nearest", "a_bilinear": "bilinear", "flat": "triangular"}
    one = xrs.rectify_dataset(ds, target_gm=tgt_gm, source_gm=src_gm, interp_methods=interp)
    many = xrs.rectify_dataset(ds, target_gm=tgt_gm, source_gm=src_gm, interp_methods=interp, devices=[0] * n_dev)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile)
    ij = orect.rectify_ij(x, y, g)
    fill = nan if np.issubdtype(dtype, np.floating) else 255 if dtype == np.uint8 else -1
    for name, src, f in (("a_nearest", data, fill), ("a_bilinear", data, fill), ("flat", single, nan)):
        want = orect.gather(src, ij, interp[name], f)
        assert_same(one[name].values, want, f"{name} single device vs oracle")
        assert_same(many[name].values, want, f"{name} {n_dev} bands vs oracle")


def test_rectify_band_uploads_only_the_footprint(xrs):
    """Byte counts of a band call: the ragged footprint of a rotated swath is a fraction of the
    scene, and results land in band-only output arrays (``Target.row0``)."""
    from xcube_resampling_b200._pipeline import Target, group_by_buffer

    x, y, size, xy_min, res, tile = _scene(900, 700, 12.0, 9, tile=256, holes=False)
    h, w = x.shape
    tgt_gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=tile)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile)
    data = np.random.default_rng(2).random((5, h, w)).astype(np.float32)
    n = 4
    edges = xrs.mg.default_band_edges(tgt_gm.height, n)
    full = orect.gather(data, orect.rectify_ij(x, y, g), "bilinear", nan)
    total_px = 0

    def worker(k, dev, exchange):
        nonlocal total_px
        r0, r1 = edges[k], edges[k + 1]
        out = np.full((5, r1 - r0, tgt_gm.width), -1.0, dtype=np.float32)
        groups = group_by_buffer([(data, Target("v", "bilinear", nan, out, row0=r0))])
        st = xrs.mg.RectifyBandStats()
        xrs.mg.rectify_band(x, y, groups, tgt_gm, edges, k, exchange, device=dev, stats=st)
        assert_same(out, full[:, r0:r1], f"band {k}")
        assert st.src_px < 0.55 * h * w, (k, st.src_px / (h * w))
        assert st.d2h_bytes == out.nbytes
        total_px += st.src_px

    xrs.mg.run_on_devices([0] * n, worker)
    assert total_px < 1.8 * h * w, total_px / (h * w)  # all bands together: not much more than the scene once


def test_reproject_dataset_on_several_devices_equals_one(xrs):
    src_gm = xrs.GridMapping.regular((400, 360), (9.0, 47.0), 0.01, "EPSG:4326")
    tgt_gm = xrs.GridMapping.regular((560, 760), (4180000.0, 2660000.0), 450.0, "EPSG:3035", tile_size=96)  # finer than the source: no pre-downscale
    rng = np.random.default_rng(3)
    data = rng.random((5, src_gm.height, src_gm.width)).astype(np.float32)
    cls = rng.integers(0, 9, (src_gm.height, src_gm.width)).astype(np.uint8)
    ds = xrs.Dataset(data_vars=dict(v=xrs.DataArray(data, dims=("band", "lat", "lon")),
                                    v_nn=xrs.DataArray(data, dims=("band", "lat", "lon")),
                                    cls=xrs.DataArray(cls, dims=("lat", "lon"))),
                     coords=dict(lon=xrs.DataArray(src_gm.x_values, dims="lon"),
                                 lat=xrs.DataArray(src_gm.y_values, dims="lat")))
    interp = {"v": "bilinear", "v_nn": "nearest", "cls": "nearest"}
    one = xrs.reproject_dataset(ds, tgt_gm, source_gm=src_gm, interp_methods=interp)
    many = xrs.reproject_dataset(ds, tgt_gm, source_gm=src_gm, interp_methods=interp, devices=[0, 0, 0])
    plan = xrs.rep.ReprojectPlan(src_gm, tgt_gm)
    direct = xrs.dev.to_host(plan.run(xrs.dev.to_device(data), "bilinear", nan))
    assert one["v"].values.dtype == np.float64
    assert_same(one["v"].values, direct, "pipeline vs direct kernel call")
    for name in interp:
        assert_same(many[name].values, one[name].values, f"{name}: 3 row bands vs one device")
    assert np.isfinite(one["v"].values).mean() > 0.3


@pytest.mark.parametrize("n_bands", [21, 24, 25])
def test_gather_band_counts_around_the_launch_split(xrs, n_bands):
    """B = 21 (the benchmark's stack: the mbarrier ring of the staged kernel wraps five times), 24 (one
    full launch) and 25 (split into two launches) against the oracle -- staged and direct kernels,
    two-step and fused forms."""
    x, y, size, xy_min, res, tile = _scene(300, 240, 12.0, 5, tile=128)
    h, w = x.shape
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile)
    gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=tile)
    data = np.random.default_rng(n_bands).random((n_bands, h, w)).astype(np.float32)
    data[3, 50:60, 50:80] = nan
    ij_ref = orect.rectify_ij(x, y, g)
    xd, yd = xrs.dev.to_device(x), xrs.dev.to_device(y)
    plan = xrs.rect.RectifyPlan(gm, xd.device)
    ij = plan.ij(xd, yd)
    assert_same(xrs.dev.to_host(ij), ij_ref, "ij")
    for pitched in (True, False):
        sd = xrs.dev.to_device_pitched(data) if pitched else xrs.dev.to_device(data)
        for method in ("nearest", "bilinear"):
            want = orect.gather(data, ij_ref, method, nan)
            assert_same(xrs.dev.to_host(xrs.rect.gather_ij(sd, ij, method, nan)), want,
                        f"two-step {method} B={n_bands} pitched={pitched}")
            assert_same(xrs.dev.to_host(plan.rectify_gather(xd, yd, sd, method, nan)), want,
                        f"fused {method} B={n_bands} pitched={pitched}")


def test_from_device_coords_equals_from_coords(xrs):
    """GridMapping derived from device-resident coordinate images (one statistics pass + edge rows /
    columns) equals the host derivation of coords.py:99-337."""
    for theta, crs in ((12.0, "EPSG:4326"), (-40.0, "EPSG:4326")):
        x, y = swath(260, 190, theta=theta, seed=2)
        x[0, 5] = nan  # NaN in the first row: nanmin / nanmax for the bounding box
        host = xrs.GridMapping.from_coords(x, y, crs, xy_dim_names=("x", "y"))
        dev = xrs.GridMapping.from_device_coords(xrs.dev.to_device(x), xrs.dev.to_device(y), crs, xy_dim_names=("x", "y"))
        assert dev.xy_res == host.xy_res and dev.size == host.size
        assert dev.xy_bbox == host.xy_bbox
        assert (dev.is_regular, dev.is_j_axis_up, dev.is_lon_360) == (host.is_regular, host.is_j_axis_up, host.is_lon_360)
        assert_same(dev.x_values, x, "x fetched on demand")
    # metre coordinates, explicit resolution, antimeridian crossing
    lon, lat = swath(120, 100, theta=5.0, seed=1, lon0=179.9)
    lon = np.where(lon > 180.0, lon - 360.0, lon)
    host = xrs.GridMapping.from_coords(lon, lat, "EPSG:4326", xy_dim_names=("x", "y"))
    dev = xrs.GridMapping.from_device_coords(xrs.dev.to_device(lon), xrs.dev.to_device(lat), "EPSG:4326",
                                             xy_dim_names=("x", "y"))
    assert host.is_lon_360 and dev.is_lon_360
    assert dev.xy_bbox == host.xy_bbox and dev.xy_res == host.xy_res
    assert_same(dev.x_values, host.x_values, "to_lon_360 on the device")


def test_cross_crs_rectify_keeps_coordinates_on_the_device(xrs, monkeypatch):
    """rectify_dataset with differing CRSs: the transformed 2-D coordinates replace the originals
    (rectify.py:214-229), never travel to the host, and the original lon/lat are not carried into the
    result."""
    from xcube_resampling_b200 import _dev

    lon, lat = swath(200, 160, theta=10.0, seed=4, lon0=9.0, lat0=2.0)
    data = np.random.default_rng(4).random((2, 160, 200)).astype(np.float32)
    ds = xrs.Dataset(data_vars=dict(rad=(("band", "y", "x"), data)),
                     coords=dict(lon=(("y", "x"), lon), lat=(("y", "x"), lat)))
    tx, ty = xrs.rep.transform_points(lon, lat, "EPSG:4326", "EPSG:32632")
    res = 300.0
    size, xy_min = covering_grid_args(tx, ty, res)
    tgt_gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:32632", tile_size=128)
    big = []
    real_to_host = _dev.to_host

    def counting_to_host(t):
        if tuple(t.shape[-2:]) == (160, 200):  # anything of the source's shape: coordinate images
            big.append(tuple(t.shape))
        return real_to_host(t)

    monkeypatch.setattr(_dev, "to_host", counting_to_host)
    out = xrs.rectify_dataset(ds, target_gm=tgt_gm, interp_methods="bilinear")
    monkeypatch.undo()
    assert not big, f"coordinate-sized device->host copies inside rectify_dataset: {big}"
    assert "lon" not in out.coords and "lat" not in out.coords and "transformed_x" not in out.coords
    assert out["rad"].dims == ("band", "y", "x") and out["x"].shape == (size[0],)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=128)
    want = orect.gather(data, orect.rectify_ij(tx, ty, g), "bilinear", nan)
    assert_same(out["rad"].values, want, "cross-CRS rectify vs oracle on the transformed coordinates")


def test_unsigned_sum_and_prod_are_uint64(xrs):
    """np.nansum / np.nanprod of unsigned data give uint64 (coarsen.py:50-90); ADVICE r1."""
    from xcube_resampling_b200 import affine

    a = np.random.default_rng(0).integers(0, 250, (8, 12)).astype(np.uint8)
    ds = xrs.Dataset(data_vars=dict(v=(("y", "x"), a)), coords=dict(x=np.arange(12) + 0.5, y=8 - (np.arange(8) + 0.5)))
    for agg in ("sum", "prod"):
        out = affine.resample_dataset(ds, ((2, 0, 0), (0, 2, 0)), ("y", "x"), (6, 4), None, interp_methods=1,
                                      agg_methods=agg)["v"].values
        assert out.dtype == np.uint64, (agg, out.dtype)


def test_rectify_band_stream_pipelines_scenes(xrs):
    """A rank working through a sequence of scenes (prologue of scene s+1 on a side stream while
    scene s streams its data): every scene's band equals the oracle."""
    from xcube_resampling_b200._pipeline import Target, group_by_buffer

    n, n_scenes = 2, 3
    scenes = []
    for s in range(n_scenes):
        x, y = swath(420, 330, theta=12.0 + 9.0 * s, seed=20 + s)
        data = np.random.default_rng(s).random((5, 330, 420)).astype(np.float32)
        scenes.append((x, y, data))
    # one target grid for all scenes (the bench's situation: scene after scene onto the same grid)
    xs = np.concatenate([sc[0].ravel() for sc in scenes])
    ys = np.concatenate([sc[1].ravel() for sc in scenes])
    res = 0.0027
    size, xy_min = covering_grid_args(xs, ys, res)
    tgt_gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=128)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=128)
    edges = xrs.mg.default_band_edges(tgt_gm.height, n)
    outs = [{m: np.full((5, tgt_gm.height, tgt_gm.width), -5.0, dtype=np.float32) for m in ("nearest", "bilinear")}
            for _ in range(n_scenes)]

    def worker(k, dev, exchange):
        jobs = []
        for s, (x, y, data) in enumerate(scenes):
            groups = group_by_buffer([(data, Target(m, m, nan, outs[s][m])) for m in ("nearest", "bilinear")])
            jobs.append((x, y, groups))
        stats = xrs.mg.rectify_band_stream(jobs, tgt_gm, edges, k, exchange, device=dev)
        assert len(stats) == n_scenes and all(st.d2h_bytes > 0 for st in stats)

    xrs.mg.run_on_devices([0] * n, worker)
    for s, (x, y, data) in enumerate(scenes):
        ij = orect.rectify_ij(x, y, g)
        for m in ("nearest", "bilinear"):
            assert_same(outs[s][m], orect.gather(data, ij, m, nan), f"scene {s} {m}")


def test_lazy_zarr_variables_stream_through_the_pipeline(xrs, tmp_path):
    """The I/O edge: variables that live in an uncompressed Zarr-v2 store are read band chunk by band
    chunk into page-locked staging buffers while the previous chunk uploads (io.py); same bytes as the
    in-memory dataset, on one device and on row bands."""
    from xcube_resampling_b200.io import LazyDataArray, open_zarr_dataset, write_zarr_array

    x, y, size, xy_min, res, tile = _scene(300, 240, 15.0, 11, holes=False)  # (no source_gm given below: the
    x[80:83, 75:150] = nan                                                  # resolution estimate must stay finite)
    h, w = x.shape
    rng = np.random.default_rng(6)
    rad = rng.random((7, h, w)).astype(np.float32)
    cls = rng.integers(0, 30, (h, w)).astype(np.uint8)
    store = tmp_path / "scene.zarr"
    write_zarr_array(str(store / "lon"), x, (64, 64), ("y", "x"))
    write_zarr_array(str(store / "lat"), y, (64, 64), ("y", "x"))
    write_zarr_array(str(store / "rad"), rad, (2, 100, 128), ("band", "y", "x"))
    write_zarr_array(str(store / "cls"), cls, (64, 300), ("y", "x"))
    lazy = open_zarr_dataset(str(store))
    assert isinstance(lazy["rad"], LazyDataArray) and isinstance(lazy["cls"], LazyDataArray)
    eager = xrs.Dataset(data_vars=dict(rad=(("band", "y", "x"), rad), cls=(("y", "x"), cls)),
                        coords=dict(lon=(("y", "x"), x), lat=(("y", "x"), y)))
    tgt_gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=tile)
    interp = {"rad": "bilinear", "cls": "nearest"}
    # (explicit source resolution: the estimate from cell areas, 0.0025, would trigger the pre-downscale)
    src_gm = xrs.GridMapping.from_coords(x, y, "EPSG:4326", xy_res=res, xy_dim_names=("x", "y"))
    want = xrs.rectify_dataset(eager, target_gm=tgt_gm, source_gm=src_gm, interp_methods=interp)
    for devices in (None, [0, 0]):
        got = xrs.rectify_dataset(lazy, target_gm=tgt_gm, source_gm=src_gm, interp_methods=interp, devices=devices)
        for name in interp:
            assert_same(got[name].values, want[name].values, f"{name} lazy store, devices={devices}")

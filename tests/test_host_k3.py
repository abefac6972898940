"""K3 as kernels -- k3_lattice_nodes and k3_reproject<T, OUT, METHOD, SEP> (csrc/reproject.cu) on top of
proj.cuh, compiled for the host (tests/hostmath.build_k3; the 256-byte L2 load hint and the L2 prefetch are
compiled out, everything else is the product's text: make_proj_consts, the lattice pre-kernel and its checks, the
separable / row-block / lattice / exact forms, the reference-tile windows and the blends) -- against the oracle's
reprojection (reproject.py:268-530 + the PROJ formulas), without a GPU.  The cases and tolerances are those of
tests/test_reproject_gpu.py: bit-exact where the transform is the identity, 1e-6 relative (nearest: rounding-tie
mismatches < 1e-3) for real projections, where libm stands in for CUDA's math library.  The per-tile source
windows come from the product's own host code with the point transform replaced by the oracle's formulas (as in
test_host_reproject_windows.py)."""

import numpy as np
import pytest

import xcube_resampling_b200 as xrs
from oracle import grid as ogrid
from oracle import proj as oproj
from oracle import reproject as orep

from .helpers import assert_same
nan = np.nan


def _fake_transform_points(x, y, from_crs, to_crs, device=None):
    """The device point transform replaced by the oracle's formulas plus PROJ's longitude wrap."""
    def proj_of(crs):
        return oproj.from_epsg(crs.epsg if crs.epsg is not None else 4326)  # OGC:CRS84: same formulas as EPSG:4326

    ox, oy = oproj.transform(proj_of(from_crs), proj_of(to_crs), np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))
    if to_crs.is_geographic:
        ox = (ox + 180.0) % 360.0 - 180.0
    return ox, oy


@pytest.fixture(scope="module")
def k3_so(tmp_path_factory):
    from . import hostmath

    try:
        return hostmath.build_k3(str(tmp_path_factory.mktemp("k3host")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise


@pytest.fixture()
def rep(monkeypatch):
    import xcube_resampling_b200.reproject as rep

    monkeypatch.setattr(rep, "transform_points", _fake_transform_points)
    return rep


def _oracle(src_gm, tgt_gm, data, method, fill, src_epsg, tgt_epsg):
    g = ogrid.regular_grid(tgt_gm.size, (tgt_gm.x_min, tgt_gm.y_min), tgt_gm.xy_res, tile_size=tgt_gm.tile_size,
                           is_j_axis_up=tgt_gm.is_j_axis_up)
    xs, ys = src_gm.x_values, src_gm.y_values
    return orep.reproject(data, float(xs[0]), float(ys[0]), src_gm.x_res, src_gm.y_res, float(ys[1] - ys[0]), g,
                          oproj.from_epsg(tgt_epsg), oproj.from_epsg(src_epsg), method, fill)


def _source(bands, h, w, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(bands, h, w)).astype(np.float32)
    a[0, h // 3, w // 4] = nan
    return a


@pytest.mark.parametrize("method", ["nearest", "bilinear", "triangular"])
@pytest.mark.parametrize("tile,j_up", [(None, False), ((32, 16), False), ((16, 16), True)])
def test_identity_transform_is_bit_exact(k3_so, rep, method, tile, j_up):
    """Geographic source, shifted / rescaled geographic target that pokes out of the source on two sides: the
    separable kernel (row-block form for bilinear), padding, reference-tile borders inside a CTA tile."""
    from . import hostmath

    src_gm = xrs.GridMapping.regular((90, 64), (10.0, 50.0), 0.01, "EPSG:4326")
    tgt_gm = xrs.GridMapping.regular((70, 53), (10.0 + 0.013 - 0.05, 50.0 - 0.021 - 0.03), 0.0113, "OGC:CRS84", tile_size=tile,
                                     is_j_axis_up=j_up)
    data = _source(3, src_gm.height, src_gm.width)
    w = rep.get_source_windows(src_gm, tgt_gm)
    want = _oracle(src_gm, tgt_gm, data, method, nan, 4326, 4326)
    got, (plan, sep, nodes) = hostmath.k3_reproject(k3_so, data, src_gm, tgt_gm, w, method, nan)
    assert sep and not nodes and got.dtype == want.dtype
    assert_same(got, want, f"{method}/{tile}/{j_up}")
    if method == "bilinear":  # cast once to the source dtype instead of the reference's float64
        got32, _ = hostmath.k3_reproject(k3_so, data, src_gm, tgt_gm, w, method, nan, out_f64=False)
        assert_same(got32, want.astype(np.float32), "bilinear, float32 out")


def test_many_bands_row_bands_and_resident_window(k3_so, rep):
    from . import hostmath

    src_gm = xrs.GridMapping.regular((90, 64), (10.0, 50.0), 0.01, "EPSG:4326")
    tgt_gm = xrs.GridMapping.regular((70, 53), (9.963, 49.949), 0.0113, "OGC:CRS84", tile_size=(32, 16))
    data = _source(27, src_gm.height, src_gm.width, seed=5)   # two launches: 24 + 3 bands
    w = rep.get_source_windows(src_gm, tgt_gm)
    want = _oracle(src_gm, tgt_gm, data, "nearest", nan, 4326, 4326)
    got, _ = hostmath.k3_reproject(k3_so, data, src_gm, tgt_gm, w, "nearest", nan)
    assert_same(got, want, "27 bands")
    want = _oracle(src_gm, tgt_gm, data[:4], "bilinear", nan, 4326, 4326)
    from xcube_resampling_b200.bands import reproject_band_footprint

    for rows in ((0, 20), (20, 41), (41, 53)):   # CTA tiles are anchored at absolute rows: bands cut through them
        got, _ = hostmath.k3_reproject(k3_so, data[:4], src_gm, tgt_gm, w, "bilinear", nan, rows=rows)
        assert_same(got, want[:, rows[0]:rows[1]], f"rows {rows}")
        fp = reproject_band_footprint(w.i0, w.j0, w.win_w, w.win_h, tgt_gm, rows, (src_gm.width, src_gm.height))
        poisoned = np.full_like(data[:4], 1e30)
        poisoned[:, fp[1]:fp[3], fp[0]:fp[2]] = data[:4, fp[1]:fp[3], fp[0]:fp[2]]
        got, _ = hostmath.k3_reproject(k3_so, poisoned, src_gm, tgt_gm, w, "bilinear", nan, rows=rows, window=fp)
        assert_same(got, want[:, rows[0]:rows[1]], f"rows {rows}, footprint {fp} resident only")


def _case_utm_from_geographic(n=150, tile=64, lat_origin=990240.0):
    """Scaled-down C3: 0.0001 deg geographic source -> 10 m UTM 32N target near 9 deg N."""
    tgt = xrs.GridMapping.regular((n, n), (399960.0, lat_origin), 10.0, "EPSG:32632", tile_size=tile)
    box = oproj.transform_bounds(oproj.from_epsg(32632), oproj.from_epsg(4326), *tgt.xy_bbox)
    res = 0.0001
    x_min = float(np.floor(box[0] / res) * res) - 4 * res
    y_min = float(np.floor(box[1] / res) * res) - 4 * res
    w = int(np.ceil((box[2] - x_min) / res)) + 4
    h = int(np.ceil((box[3] - y_min) / res)) + 4
    return xrs.GridMapping.regular((w, h), (x_min, y_min), res, "EPSG:4326"), tgt


@pytest.mark.parametrize("method", ["nearest", "bilinear", "triangular"])
def test_geographic_to_utm_lattice_and_exact_forms(k3_so, rep, method):
    from . import hostmath

    src_gm, tgt_gm = _case_utm_from_geographic()
    data = np.random.default_rng(1).random((3, src_gm.height, src_gm.width)).astype(np.float32)
    w = rep.get_source_windows(src_gm, tgt_gm)
    want = _oracle(src_gm, tgt_gm, data, method, nan, 4326, 32632)
    results = {}
    for name, kw in (("lattice from the pre-kernel", {}), ("lattice evaluated by the CTAs", dict(use_nodes=False)),
                     ("exact per pixel", dict(exact_only=True))):
        got, (plan, sep, nodes) = hostmath.k3_reproject(k3_so, data, src_gm, tgt_gm, w, method, nan, **kw)
        assert not sep and nodes is (name == "lattice from the pre-kernel") and got.dtype == want.dtype
        results[name] = got
        if method == "nearest":
            assert float(np.mean(got != want)) < 1e-3, name
        else:
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9, equal_nan=True, err_msg=name)
    assert_same(results["lattice from the pre-kernel"], results["lattice evaluated by the CTAs"], "the two lattice sources")
    if method == "bilinear":  # the bicubic lattice against the per-pixel formulas (1e-8 px of the source)
        np.testing.assert_allclose(results["lattice from the pre-kernel"], results["exact per pixel"], rtol=0, atol=1e-7,
                                   equal_nan=True)


@pytest.mark.parametrize("src_epsg,tgt_epsg,tgt_args", [
    (32632, 3035, dict(size=(120, 90), xy_min=(4320000.0, 3380000.0), xy_res=25.0)),
    (3035, 4326, dict(size=(100, 100), xy_min=(6.0, 48.0), xy_res=0.002)),
    (4326, 3857, dict(size=(128, 96), xy_min=(1100000.0, 6100000.0), xy_res=150.0)),
    (3857, 32632, dict(size=(96, 128), xy_min=(560000.0, 5930000.0), xy_res=30.0)),
])
def test_other_projection_pairs(k3_so, rep, src_epsg, tgt_epsg, tgt_args):
    from . import hostmath

    tgt = xrs.GridMapping.regular(crs=f"EPSG:{tgt_epsg}", tile_size=(64, 48), **tgt_args)
    box = np.array(oproj.transform_bounds(oproj.from_epsg(tgt_epsg), oproj.from_epsg(src_epsg), *tgt.xy_bbox))
    w, h = 140, 120
    res = max((box[2] - box[0]) / (w - 10), (box[3] - box[1]) / (h - 10))
    src = xrs.GridMapping.regular((w, h), (box[0] - 5 * res, box[1] - 5 * res), float(res), f"EPSG:{src_epsg}")
    data = np.random.default_rng(2).random((2, h, w)).astype(np.float32)
    win = rep.get_source_windows(src, tgt)
    for method in ("nearest", "bilinear"):
        want = _oracle(src, tgt, data, method, nan, src_epsg, tgt_epsg)
        got, (plan, sep, nodes) = hostmath.k3_reproject(k3_so, data, src, tgt, win, method, nan)
        assert sep is ({src_epsg, tgt_epsg} <= {4326, 3857})
        if method == "nearest":
            assert float(np.mean(got != want)) < 1e-3
        else:
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9, equal_nan=True)

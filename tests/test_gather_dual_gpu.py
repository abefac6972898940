"""GPU parity of the two-method gather (``xrs_gather_ij2``, ``k2_gather_dual``): nearest AND bilinear /
triangular samples of the same bands from one pass over the ij image, against the CPU oracle
(``_compute_var_image`` restated, rectify.py:579-734) and against two single-method gathers.

Everything is bit-exact: the nearest sample is one of the four taps of the interpolated one.
"""

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import rectify as orect

from .helpers import assert_same, covering_grid_args, hand_made_ij, load_golden, swath

pytestmark = pytest.mark.gpu
nan = np.nan


@pytest.fixture(scope="module")
def xrs():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import xcube_resampling_b200 as pkg
    from xcube_resampling_b200 import _dev, _lib, rectify

    pkg.dev, pkg.rect, pkg.lib = _dev, rectify, _lib
    return pkg


def _fill_for(dtype):
    return nan if np.issubdtype(dtype, np.floating) else 255 if np.dtype(dtype).kind == "u" else -1


def _scene(w, h, theta, seed, res, tile):
    x, y = swath(w, h, theta=theta, seed=seed)
    x[h // 3:h // 3 + 3, w // 4:w // 2] = nan  # a hole: pixels without a source inside covered tiles
    size, xy_min = covering_grid_args(x, y, res)
    return x, y, size, xy_min, ogrid.regular_grid(size, xy_min, res, tile_size=tile)


def _golden_cases():
    return [str(c) for c in load_golden("rectify.npz")["cases"]]


@pytest.mark.parametrize("case", _golden_cases())
def test_pair_against_the_reference_goldens(xrs, case):
    """Outputs of the reference's own kernels (tests/golden/rectify.npz): one launch gives the nearest AND
    the bilinear / triangular golden of every variable from the golden ij image."""
    z = load_golden("rectify.npz")
    ij = xrs.dev.to_device(np.ascontiguousarray(z[f"{case}/ij"]))
    for vname, fill in (("f32", nan), ("u8", 255), ("i16", -1), ("f64", nan)):
        src = xrs.dev.to_device_pitched(z[f"{case}/src_{vname}"])
        for method in ("bilinear", "triangular"):
            got_i, got_n = xrs.rect.gather_ij_pair(src, ij, method, fill, fill)
            assert_same(xrs.dev.to_host(got_i), z[f"{case}/out_{vname}_{method}"], f"{case} {vname}/{method}")
            assert_same(xrs.dev.to_host(got_n), z[f"{case}/out_{vname}_nearest"], f"{case} {vname}/nearest")


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8, np.int16, np.int64])
@pytest.mark.parametrize("n_bands", [1, 5])
def test_pair_against_oracle_all_dtypes(xrs, dtype, n_bands):
    """Both results of one launch vs the oracle -- staged (TMA) kernel on a pitched source, the
    two-pass fallback on an unpitched one, bilinear and triangular beside nearest, different fills."""
    x, y, size, xy_min, g = _scene(333, 260, -22.0, 21, 0.0027, 96)
    h, w = x.shape
    gm = xrs.GridMapping.regular(size, xy_min, 0.0027, "EPSG:4326", tile_size=96)
    rng = np.random.default_rng(n_bands)
    src = (rng.random((n_bands, h, w)) * 200).astype(dtype)
    if np.issubdtype(dtype, np.floating):
        src[0, 50:60, 50:80] = nan
        src[-1, 100, 100] = np.inf
    fill_i, fill_n = _fill_for(dtype), (nan if np.issubdtype(dtype, np.floating) else 7)
    ij_ref = orect.rectify_ij(x, y, g)
    xd, yd = xrs.dev.to_device(x), xrs.dev.to_device(y)
    ij = xrs.rect.RectifyPlan(gm, xd.device).ij(xd, yd)
    assert_same(xrs.dev.to_host(ij), ij_ref, "ij")
    for pitched in (True, False):
        sd = xrs.dev.to_device_pitched(src) if pitched else xrs.dev.to_device(src)
        for method in ("bilinear", "triangular"):
            got_i, got_n = xrs.rect.gather_ij_pair(sd, ij, method, fill_i, fill_n)
            what = f"{np.dtype(dtype).name} B={n_bands} {method} pitched={pitched}"
            assert_same(xrs.dev.to_host(got_i), orect.gather(src, ij_ref, method, fill_i), f"pair/{method} {what}")
            assert_same(xrs.dev.to_host(got_n), orect.gather(src, ij_ref, "nearest", fill_n), f"pair/nearest {what}")


@pytest.mark.parametrize("n_bands", [4, 21, 24, 25, 49])
def test_pair_band_counts_ring_wrap_and_launch_split(xrs, n_bands):
    """B = 4 (one ring pass), 21 (the benchmark's stack: the mbarrier ring wraps five times), 24 (one full
    launch), 25 and 49 (two and three launches) vs the oracle and vs two single-method gathers."""
    x, y, size, xy_min, g = _scene(300, 240, 12.0, 5, 0.0027, 128)
    h, w = x.shape
    gm = xrs.GridMapping.regular(size, xy_min, 0.0027, "EPSG:4326", tile_size=128)
    data = np.random.default_rng(n_bands).random((n_bands, h, w)).astype(np.float32)
    data[min(3, n_bands - 1), 50:60, 50:80] = nan
    ij_ref = orect.rectify_ij(x, y, g)
    xd, yd = xrs.dev.to_device(x), xrs.dev.to_device(y)
    ij = xrs.rect.RectifyPlan(gm, xd.device).ij(xd, yd)
    sd = xrs.dev.to_device_pitched(data)
    got_i, got_n = xrs.rect.gather_ij_pair(sd, ij, "bilinear", nan, nan)
    assert_same(xrs.dev.to_host(got_i), orect.gather(data, ij_ref, "bilinear", nan), f"pair/bilinear B={n_bands}")
    assert_same(xrs.dev.to_host(got_n), orect.gather(data, ij_ref, "nearest", nan), f"pair/nearest B={n_bands}")
    assert_same(xrs.dev.to_host(got_i), xrs.dev.to_host(xrs.rect.gather_ij(sd, ij, "bilinear", nan)), "vs gather_ij")
    assert_same(xrs.dev.to_host(got_n), xrs.dev.to_host(xrs.rect.gather_ij(sd, ij, "nearest", nan)), "vs gather_ij")


@pytest.mark.parametrize("ratio", [2.5, 6.0])
def test_pair_coarse_target_takes_the_global_tap_path(xrs, ratio):
    """Target pixels several source pixels wide: a 32x32 target tile reaches more than the 64x48 staged
    box, so the CTA reads its taps from global memory (same kernel, other branch)."""
    res = 0.0027 * ratio
    x, y, size, xy_min, g = _scene(420, 330, 17.0, 9, res, 64)
    gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=64)
    h, w = x.shape
    data = (np.random.default_rng(0).random((6, h, w)) * 1000).astype(np.float32)
    ij_ref = orect.rectify_ij(x, y, g)
    xd, yd = xrs.dev.to_device(x), xrs.dev.to_device(y)
    ij = xrs.rect.RectifyPlan(gm, xd.device).ij(xd, yd)
    assert_same(xrs.dev.to_host(ij), ij_ref, "ij")
    sd = xrs.dev.to_device_pitched(data)
    for method in ("bilinear", "triangular"):
        got_i, got_n = xrs.rect.gather_ij_pair(sd, ij, method, nan, -1.0)
        assert_same(xrs.dev.to_host(got_i), orect.gather(data, ij_ref, method, nan), f"pair/{method} ratio={ratio}")
        assert_same(xrs.dev.to_host(got_n), orect.gather(data, ij_ref, "nearest", -1.0), f"pair/nearest ratio={ratio}")


@pytest.mark.parametrize("smooth", [True, False])
@pytest.mark.parametrize("dtype", [np.float32, np.uint16])
def test_pair_image_edges_and_half_pixel_ties(xrs, smooth, dtype):
    h, w = 40, 70
    ij_ref = hand_made_ij(smooth, h, w, 64, 96)
    data = (np.random.default_rng(6).random((6, h, w)) * 100).astype(dtype)
    sd = xrs.dev.to_device_pitched(data)
    ij = xrs.dev.to_device(ij_ref)
    fill = _fill_for(dtype)
    for method in ("bilinear", "triangular"):
        got_i, got_n = xrs.rect.gather_ij_pair(sd, ij, method, fill, fill)
        assert_same(xrs.dev.to_host(got_i), orect.gather(data, ij_ref, method, fill), f"pair/{method} smooth={smooth}")
        assert_same(xrs.dev.to_host(got_n), orect.gather(data, ij_ref, "nearest", fill), f"pair/nearest smooth={smooth}")
        # and the single-method kernels on the same hand-made ij
        for m in (method, "nearest"):
            assert_same(xrs.dev.to_host(xrs.rect.gather_ij(sd, ij, m, fill)), orect.gather(data, ij_ref, m, fill), m)


def test_pair_row_band_window(xrs):
    """Only a row window of the source resident (multi-GPU footprints): window_origin / full_size."""
    x, y, size, xy_min, g = _scene(300, 260, 8.0, 3, 0.0027, 128)
    h, w = x.shape
    gm = xrs.GridMapping.regular(size, xy_min, 0.0027, "EPSG:4326", tile_size=128)
    data = (np.random.default_rng(1).random((7, h, w)) * 50).astype(np.float32)
    ij_ref = orect.rectify_ij(x, y, g)
    rows = (gm.height // 3, gm.height // 3 + 97)
    band_ij = ij_ref[:, rows[0]:rows[1]]
    jj = band_ij[1][~np.isnan(band_ij[1])]
    j0, j1 = max(int(jj.min()) - 1, 0), min(int(jj.max()) + 3, h)
    xd, yd = xrs.dev.to_device(x), xrs.dev.to_device(y)
    ij = xrs.rect.RectifyPlan(gm, xd.device, rows=rows).ij(xd, yd)
    assert_same(xrs.dev.to_host(ij), band_ij, "band ij")
    sd = xrs.dev.to_device_pitched(data[:, j0:j1])
    got_i, got_n = xrs.rect.gather_ij_pair(sd, ij, "bilinear", nan, nan, window_origin=(0, j0), full_size=(w, h))
    assert_same(xrs.dev.to_host(got_i), orect.gather(data, band_ij, "bilinear", nan), "pair/bilinear band window")
    assert_same(xrs.dev.to_host(got_n), orect.gather(data, band_ij, "nearest", nan), "pair/nearest band window")


def test_rectify_dataset_pairs_the_methods_of_one_array(xrs, monkeypatch):
    """rectify_dataset with the same host array wanted nearest + bilinear: ONE xrs_gather_ij2 per band
    chunk instead of two xrs_gather_ij; identical results with the pairing switched off; a third
    variable of another dtype keeps its own launch."""
    x, y, size, xy_min, g = _scene(360, 280, 12.0, 2, 0.0027, 128)
    h, w = x.shape
    rng = np.random.default_rng(5)
    data = rng.random((9, h, w)).astype(np.float32)
    mask = (rng.random((h, w)) * 4).astype(np.uint8)
    ds = xrs.Dataset(data_vars=dict(rad_nearest=(("band", "y", "x"), data), rad_bilinear=(("band", "y", "x"), data),
                                    mask=(("y", "x"), mask)),
                     coords=dict(lon=(("y", "x"), x), lat=(("y", "x"), y)))
    src_gm = xrs.GridMapping.from_coords(x, y, "EPSG:4326", xy_res=0.0027, xy_dim_names=("x", "y"))
    tgt_gm = xrs.GridMapping.regular(size, xy_min, 0.0027, "EPSG:4326", tile_size=128)
    interp = {"rad_nearest": "nearest", "rad_bilinear": "bilinear", "mask": "nearest"}
    fills = {"rad_nearest": -5.0, "rad_bilinear": nan, "mask": 255}
    calls = {"pair": 0, "single": 0}
    pair, single = xrs.rect.gather_ij_pair, xrs.rect.gather_ij

    def count_pair(*a, **k):
        calls["pair"] += 1
        return pair(*a, **k)

    def count_single(*a, **k):
        calls["single"] += 1
        return single(*a, **k)

    monkeypatch.setattr(xrs.rect, "gather_ij_pair", count_pair)
    monkeypatch.setattr(xrs.rect, "gather_ij", count_single)
    monkeypatch.setattr(xrs.rect, "_PIPELINE_MIN_BYTES", 0)
    monkeypatch.setattr(xrs.rect, "DUAL_GATHER", True)
    paired = xrs.rectify_dataset(ds, target_gm=tgt_gm, source_gm=src_gm, interp_methods=interp, fill_values=fills)
    n_pair, n_single = calls["pair"], calls["single"]
    assert n_pair >= 1 and n_single >= 1, calls  # the radiances pair up, the mask goes alone
    calls.update(pair=0, single=0)
    monkeypatch.setattr(xrs.rect, "DUAL_GATHER", False)
    plain = xrs.rectify_dataset(ds, target_gm=tgt_gm, source_gm=src_gm, interp_methods=interp, fill_values=fills)
    assert calls["pair"] == 0 and calls["single"] == 2 * n_pair + n_single, calls
    ij_ref = orect.rectify_ij(x, y, g)
    for name, src in (("rad_nearest", data), ("rad_bilinear", data), ("mask", mask)):
        want = orect.gather(src, ij_ref, interp[name], fills[name])
        assert_same(paired[name].values, want, f"{name}: paired vs oracle")
        assert_same(plain[name].values, want, f"{name}: per-method vs oracle")


def test_c_abi_rejects_bad_arguments(xrs):
    import ctypes

    lib = xrs.lib.load()
    arr = (ctypes.c_void_p * 1)(16)
    ok = ctypes.c_void_p(16)
    args = [arr, arr, arr, 1, 0, 8, 8, 8, 0, 0, 8, 8, ok, 8, 8, 1, 0.0, 0.0, None]

    def call(**over):
        a = list(args)
        for k, v in over.items():
            a[int(k[1:])] = v
        return lib.xrs_gather_ij2(*a)

    assert call(_15=0) != 0  # nearest is not a valid partner method
    assert b"bilinear" in lib.xrs_last_error()
    assert call(_3=0) != 0   # no bands
    assert call(_12=None) != 0  # no ij
    assert call(_10=9) != 0  # window wider than the image

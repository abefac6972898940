"""Host test (no GPU) of ``resample_in_space``'s decision rules (reference ``spatial.py:82-131``): which
of rectify / affine / reproject a (source grid mapping, target grid mapping) pair is routed to, and the two
cases in which the source dataset comes back untouched.  The three back ends are replaced by recorders."""

import logging

import numpy as np
import pytest

import xcube_resampling_b200 as xrs
from xcube_resampling_b200 import affine, rectify, reproject


@pytest.fixture
def calls(monkeypatch):
    seen = []

    def recorder(name):
        def fn(source_ds, target_gm=None, **kw):
            seen.append((name, target_gm, kw))
            return name

        return fn

    monkeypatch.setattr(rectify, "rectify_dataset", recorder("rectify"))
    monkeypatch.setattr(affine, "affine_transform_dataset", recorder("affine"))
    monkeypatch.setattr(reproject, "reproject_dataset", recorder("reproject"))
    return seen


def _regular_ds(w=8, h=6, res=0.5, x0=10.0, y0=50.0):
    lon = x0 + res * (np.arange(w) + 0.5)
    lat = y0 + res * (h - np.arange(h) - 0.5)
    return xrs.Dataset(data_vars=dict(rad=(("lat", "lon"), np.ones((h, w), np.float32))),
                       coords=dict(lon=("lon", lon), lat=("lat", lat)))


def _swath_ds():
    lon = np.array([[1.0, 6.0], [0.0, 2.0]])
    lat = np.array([[56.0, 53.0], [52.0, 50.0]])
    return xrs.Dataset(data_vars=dict(rad=(("y", "x"), np.ones((2, 2)))), coords=dict(lon=(("y", "x"), lon), lat=(("y", "x"), lat)))


def test_irregular_source_goes_to_rectify_with_all_options(calls):
    ds = _swath_ds()
    tgt = xrs.GridMapping.regular((4, 4), (-1, 49), 2, "EPSG:4326")
    out = xrs.resample_in_space(ds, target_gm=tgt, variables="rad", interp_methods="nearest", agg_methods="mean",
                                recover_nans=True, fill_values=-1, tile_size=3)
    assert out == "rectify" and len(calls) == 1
    name, target_gm, kw = calls[0]
    assert target_gm is tgt and kw["tile_size"] == 3 and kw["variables"] == "rad" and kw["interp_methods"] == "nearest"
    assert kw["agg_methods"] == "mean" and kw["recover_nans"] is True and kw["fill_values"] == -1
    assert not kw["source_gm"].is_regular
    # no target grid mapping: rectify derives one (rectify.py:113-114)
    assert xrs.resample_in_space(ds) == "rectify" and calls[1][1] is None


def test_regular_source_without_target_returns_the_source_with_a_warning(calls, caplog):
    ds = _regular_ds()
    with caplog.at_level(logging.WARNING):
        assert xrs.resample_in_space(ds) is ds
    assert "`target_gm` must be given" in caplog.text and not calls


def test_close_grids_return_the_source(calls):
    ds = _regular_ds()
    src = xrs.GridMapping.from_dataset(ds)
    same = xrs.GridMapping.regular(src.size, (src.x_min, src.y_min), src.xy_res, src.crs)
    assert xrs.resample_in_space(ds, target_gm=same) is ds and not calls


def test_same_crs_goes_to_affine_other_crs_to_reproject(calls):
    ds = _regular_ds()
    coarser = xrs.GridMapping.regular((4, 3), (10.0, 50.0), 1.0, "EPSG:4326")
    assert xrs.resample_in_space(ds, target_gm=coarser, interp_methods=1) == "affine"
    assert calls[-1][2]["interp_methods"] == 1 and "tile_size" not in calls[-1][2]
    # both geographic counts as "same" for the affine rule (utils.py:_can_apply_affine_transform)
    crs84 = xrs.GridMapping.regular((4, 3), (10.0, 50.0), 1.0, "OGC:CRS84")
    assert xrs.resample_in_space(ds, target_gm=crs84) == "affine"
    utm = xrs.GridMapping.regular((10, 10), (500000.0, 5540000.0), 10000.0, "EPSG:32632")
    assert xrs.resample_in_space(ds, target_gm=utm) == "reproject"
    assert [c[0] for c in calls] == ["affine", "affine", "reproject"]


def test_target_must_be_regular(calls):
    ds = _regular_ds()
    irregular = xrs.GridMapping.from_dataset(_swath_ds())
    with pytest.raises(ValueError, match="target_gm"):
        xrs.resample_in_space(ds, target_gm=irregular)
    assert not calls

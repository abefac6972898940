"""Host logic of the reprojection windows (reproject.py:385-469) without a GPU: the device point
transform is replaced by the oracle's formulas plus PROJ's longitude wrap."""

import numpy as np

from oracle import grid as ogrid
from oracle import proj as oproj
from oracle import reproject as orep


def _fake_transform_points(x, y, from_crs, to_crs, device=None):
    fp, tp = oproj.from_epsg(from_crs.epsg), oproj.from_epsg(to_crs.epsg)
    ox, oy = oproj.transform(fp, tp, np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))
    if to_crs.is_geographic:  # PROJ reduces inverse-projected longitudes to [-180, 180]
        ox = (ox + 180.0) % 360.0 - 180.0
    return ox, oy


def test_unwrap_longitudes():
    from xcube_resampling_b200.reproject import _unwrap_longitudes

    lon = np.array([[170.0, 175.0, 179.9, -179.9, -175.0, 179.0, 170.0],
                    [-170.0, -179.0, 179.5, 178.0, np.nan, 178.0, -178.0]])
    out = _unwrap_longitudes(lon)
    np.testing.assert_allclose(out[0], [170.0, 175.0, 179.9, 180.1, 185.0, 179.0, 170.0])
    np.testing.assert_allclose(out[1][:4], [-170.0, -179.0, -180.5, -182.0])
    assert np.isnan(out[1][4])
    same = np.array([[10.0, 11.0, 12.0, 11.0]])
    assert np.array_equal(_unwrap_longitudes(same), same)


def test_global_web_mercator_windows_at_the_antimeridian(monkeypatch):
    """BASELINE config 5: the easternmost tiles end beyond x = +20037508.34 (GridMapping.regular
    rounds the bounds like the reference, helpers.py:39-48), where the transform wraps to -180.
    The tile windows must keep their real extent (4502 columns), not span the globe."""
    import xcube_resampling_b200.reproject as rep
    from xcube_resampling_b200.gridmapping import GridMapping

    monkeypatch.setattr(rep, "transform_points", _fake_transform_points)
    ext = 20037508.342789244
    tgt = GridMapping.regular((36000, 36000), (-ext, -ext), 2 * ext / 36000, "EPSG:3857", tile_size=4500)
    src = GridMapping.regular((36000, 18000), (-180.0, -90.0), 0.01, "EPSG:4326")
    assert tgt.x_max > ext  # the rounded bound lies beyond the antimeridian
    w = rep.get_source_windows(src, tgt)
    g = ogrid.regular_grid(tgt.size, (tgt.x_min, tgt.y_min), tgt.xy_res, tile_size=tgt.tile_size,
                           is_j_axis_up=tgt.is_j_axis_up)
    xs, ys = src.x_values, src.y_values
    win = orep.source_windows(float(xs[0]), float(ys[0]), src.x_res, src.y_res, float(ys[1] - ys[0]), src.width,
                              src.height, g, oproj.from_epsg(3857), oproj.from_epsg(4326))  # oracle: no wrap at all
    assert (w.win_w, w.win_h) == (win["win_w"], win["win_h"]) == (4502, 4100)
    assert np.array_equal(w.i0, win["i0"]) and np.array_equal(w.j0, win["j0"])
    assert np.array_equal(w.x0, win["x0"]) and np.array_equal(w.y0, win["y0"])
    assert w.x0.dtype == np.float32


def test_gridmapping_transform_and_to_regular(monkeypatch):
    """tests/gridmapping/test_base.py:326-404 with the device transform replaced by the oracle's."""
    import xcube_resampling_b200.reproject as rep
    from xcube_resampling_b200.crs import CRS
    from xcube_resampling_b200.gridmapping import GridMapping

    monkeypatch.setattr(rep, "transform_points", _fake_transform_points)
    gm = GridMapping.regular((400, 200), (20, 56), 0.01, "EPSG:4326")
    t = gm.transform("EPSG:32633")
    assert t is not gm and t.crs == CRS.from_epsg(32633) and t.is_regular is False
    assert (t.size, t.tile_size, t.is_j_axis_up) == ((400, 200), (400, 200), False)
    assert t.xy_var_names == ("transformed_x", "transformed_y") and t.xy_dim_names == ("lon", "lat")

    t = gm.derive(tile_size=(200, 200)).transform("EPSG:32633", xy_res=1000)
    assert (t.size, t.tile_size, t.xy_res, t.is_j_axis_up) == ((400, 200), (200, 200), (1000, 1000), False)
    r = t.to_regular()
    assert r.is_regular and r.crs == CRS.from_epsg(32633)
    assert (r.size, r.tile_size, r.xy_res, r.is_j_axis_up) == ((267, 249), (200, 200), (1000, 1000), False)
    assert r.xy_var_names == ("x", "y") and r.xy_dim_names == ("x", "y")

    r = GridMapping.regular((1000, 1000), (9.6, 47.6), 0.0002, "EPSG:4326").transform("EPSG:32633").to_regular()
    assert (r.size, r.tile_size, r.is_j_axis_up, r.is_lon_360) == ((827, 1163), (1000, 1000), False, False)

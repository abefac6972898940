"""The multi-GPU rectification prologue (multigpu.RectifyBandJob.prologue) with every kernel compiled for the
host (tests/hostmath: K0 in min-form, KB quad footprints, K1) and the product's own host logic
(multigpu.source_slabs, footprint_segments, rectify._separable_axes), without a GPU:

  participant k scans ITS slab of the swath (K0 tile windows + KB footprints into a min-form table),
  the tables are merged with MIN, windows are finalised, and K1 runs on band k's rows reading only the
  band's footprint.

"Only the footprint is resident" is made literal: outside the rectangles `footprint_segments` would upload,
the coordinates (and the data) handed to K1 (and to the gather) are replaced by finite garbage.  The band's rows
must still equal the oracle's -- i.e. every quad that can claim a pixel of the band, and every tap a gather
through the band's ij can read, lies inside the uploaded rectangles."""

import numpy as np
import pytest

import xcube_resampling_b200 as xrs
from oracle import grid as ogrid
from oracle import rectify as orect
from xcube_resampling_b200 import multigpu
from xcube_resampling_b200.rectify import _separable_axes, _xy_border

from .helpers import assert_same, covering_grid_args, quad_footprints_np, swath


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    from . import hostmath

    d = str(tmp_path_factory.mktemp("bandhost"))
    try:
        return hostmath.build_k0(d), hostmath.build_kb(d), hostmath.build_k1(d)
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise


@pytest.mark.parametrize("shape,theta,n,tile,j_up", [
    ((150, 170), 30.0, 3, 48, False),   # rotated swath: diagonal footprints
    ((120, 140), -12.0, 4, 64, False),
    ((90, 200), 60.0, 2, (40, 24), True),
])
def test_band_prologue_on_the_host(libs, shape, theta, n, tile, j_up):
    from . import hostmath

    k0_so, kb_so, k1_so = libs
    w, h = shape
    x, y = swath(w, h, theta=theta, seed=7 * w + h)
    x[h // 2, w // 3:w // 3 + 9] = np.nan
    res = 0.0027
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile, is_j_axis_up=j_up)
    gm = xrs.GridMapping.regular(size, xy_min, res, "EPSG:4326", tile_size=tile, is_j_axis_up=j_up)
    want_windows = orect.source_windows(x, y, g)
    want_ij = orect.rectify_ij(x, y, g, windows=want_windows)

    group = 32
    edges = multigpu.default_band_edges(g.height, n, align=8)
    slabs = multigpu.source_slabs(h, n, group)
    assert slabs[0][0] == 0 and slabs[-1][1] == h and all(s[0] % group == 0 for s in slabs)

    # 1. + 2. slab scans of all participants, merged with MIN (hostmath scans slab after slab into one table,
    # which is what the element-wise minimum of the partial tables gives)
    x_lo, x_hi, y_lo, y_hi, _, _ = _separable_axes(gm.xy_bboxes, _xy_border(gm))
    windows = None
    for s0, s1 in slabs:  # K0 form 2 takes equal slab heights: scan each participant's slab by its own call
        part = hostmath.k0(k0_so, x[s0:s1], y[s0:s1], x_lo, x_hi, y_lo, y_hi, ij_border=0, form=2, slab_rows=s1 - s0)
        part = np.where(part[:, :1] == -1, part, part + np.array([0, s0, 0, s0]))
        if windows is None:
            windows = part
        else:
            both, only_new = (windows[:, 0] != -1) & (part[:, 0] != -1), (windows[:, 0] == -1)
            merged = np.where(both[:, None], np.stack([np.minimum(windows[:, 0], part[:, 0]), np.minimum(windows[:, 1], part[:, 1]),
                                                       np.maximum(windows[:, 2], part[:, 2]), np.maximum(windows[:, 3], part[:, 3])], 1),
                              np.where(only_new[:, None], part, windows))
            windows = merged
    grown = windows + np.array([-1, -1, 1, 1])  # ij_border = 1 and the clip of bboxes.py:90-106
    grown[:, 0:2] = np.maximum(grown[:, 0:2], 0)
    grown[:, 2] = np.minimum(grown[:, 2], w)
    grown[:, 3] = np.minimum(grown[:, 3], h)
    windows = np.where(windows[:, :1] == -1, windows, grown)
    assert_same(windows, want_windows, "tile windows from the participants' slab scans")

    fp = hostmath.band_quad_footprints(kb_so, x, y, g, edges, slabs=[(s0, min(h, s1 + 1)) for s0, s1 in slabs if s1 > s0])
    assert_same(fp, quad_footprints_np(x, y, g, edges, group=group), "KB (host build of bands.cu) vs its numpy restatement")

    rng = np.random.default_rng(5)
    src = rng.random((2, h, w)).astype(np.float32)
    for k in range(n):
        rows = (edges[k], edges[k + 1])
        if rows[1] <= rows[0]:
            continue
        window, segments, n_px = multigpu.footprint_segments(fp[k], h, w, group)
        if window is None:
            assert np.isnan(want_ij[:, rows[0]:rows[1]]).all()
            continue
        # 3. only the footprint is resident: everything else is garbage
        resident = np.zeros((h, w), dtype=bool)
        for j0, j1, i0, i1 in segments:
            resident[j0:j1, i0:i1] = True
        assert resident.sum() == n_px and window == (segments[0][0], segments[-1][1])
        xg = np.where(resident, x, rng.uniform(np.nanmin(x), np.nanmax(x), x.shape))
        yg = np.where(resident, y, rng.uniform(np.nanmin(y), np.nanmax(y), y.shape))
        ij, _, _ = hostmath.k1(k1_so, xg, yg, windows, g, rows=rows, fp_cols=fp[k])
        assert_same(ij, want_ij[:, rows[0]:rows[1]], f"band {k}: K1 on the footprint alone")
        # ... and so are the data: the gather through the band's ij reads resident pixels only
        srcg = np.where(resident, src, np.float32(1e30))
        for method in ("nearest", "bilinear", "triangular"):
            assert_same(orect.gather(srcg, ij, method, np.nan), orect.gather(src, ij, method, np.nan), f"band {k}: {method}")
    # less than the whole image for every participant -- even at these sizes, where the 32-row groups and the
    # 32-column alignment of the rectangles weigh heavily (config C2 at N = 8: 1.28 x one image in total)
    total = sum(multigpu.footprint_segments(fp[k], h, w, group)[2] for k in range(n))
    assert total < n * h * w

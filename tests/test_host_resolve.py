"""K1's resolve step -- the text of csrc/rectify_common.cuh (resolve_row / resolve_pixel with their
multiply-high divisions) compiled for the HOST (tests/hostmath) -- against the oracle's ij image, bit for bit,
without a GPU.

The claim words (which source quad and which of its two triangles owns a target pixel) come from a plain
Python walk over the quads in the reference's order (rectify.py:458-576: tile by tile, quads row-major inside
the tile's source window, first writer wins); that walk also computes ij itself and must reproduce the oracle
before its claims are trusted."""

import math

import numpy as np
import pytest

from oracle import grid as ogrid
from oracle import rectify as orect

from .helpers import assert_same, covering_grid_args, swath

UV_DELTA = 1e-3
NOCLAIM = 0xFFFFFFFF


def _det(ax, ay, bx, by, cx, cy):
    return (ax - bx) * (ay - cy) - (ax - cx) * (ay - by)


def _u(px, py, ax, ay, cx, cy):
    return (ax - px) * (ay - cy) - (ay - py) * (ax - cx)


def _v(px, py, ax, ay, bx, by):
    return (ay - py) * (ax - bx) - (ax - px) * (ay - by)


def _clamp01(t):
    return 0.0 if t < 0.0 else (1.0 if t > 1.0 else t)


def _floor_i64(v):
    return None if not math.isfinite(v) else math.floor(v)


def claims_and_ij(x, y, g, windows):
    """First-writer-wins scatter of rectify.py:458-576 in plain Python: (claims uint32 (H, W), ij (2, H, W))."""
    h, w = x.shape
    nqi = w - 1
    claims = np.full((g.height, g.width), NOCLAIM, dtype=np.uint32)
    ij = np.full((2, g.height, g.width), np.nan)
    lo, hi = -UV_DELTA, 1.0 + 2 * UV_DELTA
    nty, ntx = g.n_tiles
    for ty in range(nty):
        for tx in range(ntx):
            bb = windows[ty * ntx + tx]
            if bb[0] == -1:
                continue
            r0, c0 = ty * g.tile_h, tx * g.tile_w
            th, tw = min(g.tile_h, g.height - r0), min(g.tile_w, g.width - c0)
            x_off = g.x_min + float(c0) * g.x_res
            y_off = (g.y_min + float(r0) * g.y_res) if g.is_j_axis_up else (g.y_max - float(r0) * g.y_res)
            x_scale, y_scale = g.x_res, (g.y_res if g.is_j_axis_up else -g.y_res)
            j_end, i_end = min(bb[3] + 1, h), min(bb[2] + 1, w)
            for j0 in range(bb[1], j_end - 1):
                for i0 in range(bb[0], i_end - 1):
                    qx = (x[j0, i0], x[j0, i0 + 1], x[j0 + 1, i0], x[j0 + 1, i0 + 1])
                    qy = (y[j0, i0], y[j0, i0 + 1], y[j0 + 1, i0], y[j0 + 1, i0 + 1])
                    pis = [_floor_i64((qx[k] - x_off) / x_scale) for k in range(4)]
                    pjs = [_floor_i64((qy[k] - y_off) / y_scale) for k in range(4)]
                    if None in pis or None in pjs:
                        continue  # this test's swaths are finite; non-finite vertices are the GPU suite's business
                    ci_lo, ci_hi, cj_lo, cj_hi = min(pis), max(pis), min(pjs), max(pjs)
                    if ci_hi < 0 or cj_hi < 0 or ci_lo >= tw or cj_lo >= th:
                        continue
                    ci_lo, ci_hi, cj_lo, cj_hi = max(ci_lo, 0), min(ci_hi, tw - 1), max(cj_lo, 0), min(cj_hi, th - 1)
                    det_a = _det(qx[0], qy[0], qx[1], qy[1], qx[2], qy[2])
                    det_b = _det(qx[3], qy[3], qx[2], qy[2], qx[1], qy[1])
                    if det_a == 0.0 and det_b == 0.0:
                        continue
                    for dj in range(cj_lo, cj_hi + 1):
                        py = y_off + (float(dj) + 0.5) * y_scale
                        for di in range(ci_lo, ci_hi + 1):
                            if claims[r0 + dj, c0 + di] != NOCLAIM:
                                continue
                            px = x_off + (float(di) + 0.5) * x_scale
                            tri = None
                            if det_a != 0.0:
                                u = _u(px, py, qx[0], qy[0], qx[2], qy[2]) / det_a
                                v = _v(px, py, qx[0], qy[0], qx[1], qy[1]) / det_a
                                if u >= lo and v >= lo and u + v <= hi:
                                    tri, si, sj = 0, float(i0 - bb[0]) + _clamp01(u), float(j0 - bb[1]) + _clamp01(v)
                            if tri is None and det_b != 0.0:
                                u = _u(px, py, qx[3], qy[3], qx[1], qy[1]) / det_b
                                v = _v(px, py, qx[3], qy[3], qx[2], qy[2]) / det_b
                                if u >= lo and v >= lo and u + v <= hi:
                                    tri = 1
                                    si, sj = float(i0 - bb[0] + 1) - _clamp01(u), float(j0 - bb[1] + 1) - _clamp01(v)
                            if tri is not None:
                                claims[r0 + dj, c0 + di] = 2 * (j0 * nqi + i0) + tri
                                ij[0, r0 + dj, c0 + di] = float(bb[0]) + si
                                ij[1, r0 + dj, c0 + di] = float(bb[1]) + sj
    return claims, ij


@pytest.fixture(scope="module")
def resolve_so(tmp_path_factory):
    from . import hostmath

    try:
        return hostmath.build_resolve(str(tmp_path_factory.mktemp("resolvehost")))
    except RuntimeError as e:
        if "g++ not available" in str(e):
            pytest.skip(str(e))
        raise


@pytest.mark.parametrize("shape,theta,res_factor,tile,j_up", [
    ((46, 38), 12.0, 1.0, 16, False),        # several small reference tiles, quads ~1 px
    ((40, 33), -35.0, 0.6, (23, 9), False),  # finer target (quads span 2-3 px), ragged non-square tiles
    ((52, 30), 77.0, 1.7, None, False),      # coarser target (several quads per pixel), one tile
    ((37, 41), 5.0, 1.0, 12, True),          # j axis up
])
def test_resolve_step_reproduces_the_oracle_ij(resolve_so, shape, theta, res_factor, tile, j_up):
    from . import hostmath

    w, h = shape
    x, y = swath(w, h, theta=theta, seed=w * h)
    res = 0.0027 * res_factor
    size, xy_min = covering_grid_args(x, y, res)
    g = ogrid.regular_grid(size, xy_min, res, tile_size=tile, is_j_axis_up=j_up)
    windows = orect.source_windows(x, y, g)
    want = orect.rectify_ij(x, y, g, windows=windows)
    claims, ij_walk = claims_and_ij(x, y, g, windows)
    assert_same(ij_walk, want, "the Python walk's own ij vs the oracle (validates the claim words)")
    claimed = claims != NOCLAIM
    assert np.array_equal(claimed, ~np.isnan(want[0])) and 0.3 < claimed.mean() < 0.95
    assert (claims[claimed] & 1).any() and not (claims[claimed] & 1).all()  # both triangles occur
    got = hostmath.resolve(resolve_so, x, y, windows, claims, g)
    assert_same(got, want, "resolve_pixel (host build of rectify_common.cuh) vs the oracle")

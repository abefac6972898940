"""fp64 restatement of the PROJ formulas the resampling path needs (test infrastructure only).

The reference delegates every CRS transform to ``pyproj.Transformer`` -> the PROJ C library
(``reproject.py:124,347,398,483``; ``rectify.py:196-204``; ``gridmapping/transform.py:77-91``).
Neither pyproj nor PROJ is present (``pyproject.toml:49`` pins only ``pyproj>=3.0``), so the
published algorithms are restated here with numpy, always in ``always_xy=True`` axis order:

* geographic (EPSG:4326, OGC:CRS84, EPSG:4258): degrees, identity;
* transverse Mercator / UTM: Krueger series to n^6 as published by Karney (2011), which is what
  PROJ's default ("exact" / Poder-Engsager) ``tmerc`` evaluates;
* EPSG:3857 web Mercator: spherical formulas on the WGS84 semi-major axis;
* Lambert azimuthal equal-area, ellipsoidal (EPSG:3035): Snyder (1987) eqs. 24-*, authalic latitude
  inverted with Newton's method.
* datum shift WGS84 <-> ETRS89: identity (PROJ's "ballpark" behaviour for these pairs).

Pinned against: the CRS84 -> EPSG:32632 known-answer of the reference
(``tests/gridmapping/test_transform.py:46-65``, 7 decimals), the EPSG Guidance Note 7-2 LAEA
worked example, and -- through ``oracle.reproject`` -- the nearest-neighbour goldens of
``tests/test_reproject.py:21-201``.  EPSG:3857 has no golden in the reference: parity unpinned.
"""

import numpy as np

GEOGRAPHIC, TMERC, WEBMERC, LAEA = 0, 1, 2, 3

WGS84_A = 6378137.0
WGS84_INV_F = 298.257223563
GRS80_INV_F = 298.257222101


class Proj:
    """kind + parameters (same fields as ``struct xrs_proj`` in include/xrs.h)."""

    def __init__(self, kind, a=WGS84_A, inv_f=WGS84_INV_F, lon0=0.0, lat0=0.0, k0=1.0, fe=0.0, fn=0.0):
        self.kind, self.a, self.inv_f = kind, a, inv_f
        self.lon0, self.lat0, self.k0, self.fe, self.fn = lon0, lat0, k0, fe, fn

    @property
    def f(self):
        return 1.0 / self.inv_f if self.inv_f else 0.0


def from_epsg(code: int) -> Proj:
    code = int(code)
    if code in (4326, 84):
        return Proj(GEOGRAPHIC)
    if code == 4258:
        return Proj(GEOGRAPHIC, inv_f=GRS80_INV_F)
    if 32601 <= code <= 32660 or 32701 <= code <= 32760:
        zone = code % 100
        return Proj(TMERC, lon0=6.0 * zone - 183.0, k0=0.9996, fe=500000.0, fn=10000000.0 if code >= 32700 else 0.0)
    if 25828 <= code <= 25838:
        return Proj(TMERC, inv_f=GRS80_INV_F, lon0=6.0 * (code % 100) - 183.0, k0=0.9996, fe=500000.0)
    if code == 3857:
        return Proj(WEBMERC)
    if code == 3035:
        return Proj(LAEA, inv_f=GRS80_INV_F, lon0=10.0, lat0=52.0, fe=4321000.0, fn=3210000.0)
    raise ValueError(f"EPSG:{code} not supported")


# ---------------------------------------------------------------------------
# transverse Mercator (Karney 2011, eqs. 7-12, 35-36)
# ---------------------------------------------------------------------------
def _kruger(f):
    n = f / (2.0 - f)
    n2, n3, n4, n5, n6 = n * n, n ** 3, n ** 4, n ** 5, n ** 6
    alpha = np.array([
        n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800,
        13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360,
        61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440,
        49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600,
        34729 * n5 / 80640 - 3418889 * n6 / 1995840,
        212378941 * n6 / 319334400,
    ])
    beta = np.array([
        n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800,
        n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720,
        17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720,
        4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600,
        4583 * n5 / 161280 - 108847 * n6 / 3991680,
        20648693 * n6 / 638668800,
    ])
    a_scale = (1 + n2 / 4 + n4 / 64 + n6 / 256) / (1 + n)  # A / a
    return alpha, beta, a_scale


def _taup(tau, e):
    """conformal-latitude tangent from geodetic-latitude tangent."""
    sigma = np.sinh(e * np.arctanh(e * tau / np.sqrt(1 + tau * tau)))
    return tau * np.sqrt(1 + sigma * sigma) - sigma * np.sqrt(1 + tau * tau)


def _tau_from_taup(taup, e):
    e2m = 1 - e * e
    tau = taup / e2m
    for _ in range(6):
        tp = _taup(tau, e)
        dtau = (taup - tp) / np.sqrt(1 + tp * tp) * (1 + e2m * tau * tau) / (e2m * np.sqrt(1 + tau * tau))
        tau = tau + dtau
    return tau


def _tmerc_xi0(p: Proj, alpha, e):
    """xi of the latitude of origin on the central meridian."""
    if p.lat0 == 0.0:
        return 0.0
    taup = _taup(np.tan(np.deg2rad(p.lat0)), e)
    xi_p = np.arctan2(taup, 1.0)
    j = np.arange(1, 7)
    return float(xi_p + np.sum(alpha * np.sin(2 * j * xi_p)))


def tmerc_forward(p: Proj, lon, lat):
    f = p.f
    e = np.sqrt(f * (2 - f))
    alpha, _, a_scale = _kruger(f)
    lam = np.deg2rad(np.asarray(lon, dtype=np.float64) - p.lon0)
    phi = np.deg2rad(np.asarray(lat, dtype=np.float64))
    taup = _taup(np.tan(phi), e)
    xi_p = np.arctan2(taup, np.cos(lam))
    eta_p = np.arcsinh(np.sin(lam) / np.hypot(taup, np.cos(lam)))
    xi, eta = xi_p.copy(), eta_p.copy()
    for j in range(1, 7):
        xi = xi + alpha[j - 1] * np.sin(2 * j * xi_p) * np.cosh(2 * j * eta_p)
        eta = eta + alpha[j - 1] * np.cos(2 * j * xi_p) * np.sinh(2 * j * eta_p)
    ka = p.k0 * p.a * a_scale
    return p.fe + ka * eta, p.fn + ka * (xi - _tmerc_xi0(p, alpha, e))


def tmerc_inverse(p: Proj, x, y):
    f = p.f
    e = np.sqrt(f * (2 - f))
    alpha, beta, a_scale = _kruger(f)
    ka = p.k0 * p.a * a_scale
    xi = (np.asarray(y, dtype=np.float64) - p.fn) / ka + _tmerc_xi0(p, alpha, e)
    eta = (np.asarray(x, dtype=np.float64) - p.fe) / ka
    xi_p, eta_p = xi.copy(), eta.copy()
    for j in range(1, 7):
        xi_p = xi_p - beta[j - 1] * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
        eta_p = eta_p - beta[j - 1] * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
    taup = np.sin(xi_p) / np.hypot(np.sinh(eta_p), np.cos(xi_p))
    lam = np.arctan2(np.sinh(eta_p), np.cos(xi_p))
    tau = _tau_from_taup(taup, e)
    return np.rad2deg(lam) + p.lon0, np.rad2deg(np.arctan(tau))


# ---------------------------------------------------------------------------
# web Mercator
# ---------------------------------------------------------------------------
def webmerc_forward(p: Proj, lon, lat):
    lam = np.deg2rad(np.asarray(lon, dtype=np.float64) - p.lon0)
    phi = np.deg2rad(np.asarray(lat, dtype=np.float64))
    return p.fe + p.a * lam, p.fn + p.a * np.arcsinh(np.tan(phi))


def webmerc_inverse(p: Proj, x, y):
    lam = (np.asarray(x, dtype=np.float64) - p.fe) / p.a
    phi = np.arctan(np.sinh((np.asarray(y, dtype=np.float64) - p.fn) / p.a))
    return np.rad2deg(lam) + p.lon0, np.rad2deg(phi)


# ---------------------------------------------------------------------------
# Lambert azimuthal equal-area, ellipsoidal, oblique aspect (Snyder 1987, pp. 187-190)
# ---------------------------------------------------------------------------
def _q(sin_phi, e):
    es = e * sin_phi
    return (1 - e * e) * (sin_phi / (1 - es * es) - (0.5 / e) * np.log((1 - es) / (1 + es)))


def _laea_consts(p: Proj):
    f = p.f
    e = np.sqrt(f * (2 - f))
    phi1 = np.deg2rad(p.lat0)
    qp = _q(1.0, e)
    q1 = _q(np.sin(phi1), e)
    beta1 = np.arcsin(q1 / qp)
    rq = p.a * np.sqrt(qp / 2)
    m1 = np.cos(phi1) / np.sqrt(1 - (e * np.sin(phi1)) ** 2)
    d = p.a * m1 / (rq * np.cos(beta1))
    return e, qp, beta1, rq, d


def laea_forward(p: Proj, lon, lat):
    e, qp, beta1, rq, d = _laea_consts(p)
    lam = np.deg2rad(np.asarray(lon, dtype=np.float64) - p.lon0)
    phi = np.deg2rad(np.asarray(lat, dtype=np.float64))
    beta = np.arcsin(np.clip(_q(np.sin(phi), e) / qp, -1.0, 1.0))
    b = rq * np.sqrt(2 / (1 + np.sin(beta1) * np.sin(beta) + np.cos(beta1) * np.cos(beta) * np.cos(lam)))
    x = p.fe + b * d * np.cos(beta) * np.sin(lam)
    y = p.fn + (b / d) * (np.cos(beta1) * np.sin(beta) - np.sin(beta1) * np.cos(beta) * np.cos(lam))
    return x, y


def laea_inverse(p: Proj, x, y):
    e, qp, beta1, rq, d = _laea_consts(p)
    xx = (np.asarray(x, dtype=np.float64) - p.fe) / d
    yy = (np.asarray(y, dtype=np.float64) - p.fn) * d
    rho = np.hypot(xx, yy)
    ce = 2 * np.arcsin(np.clip(rho / (2 * rq), -1.0, 1.0))
    with np.errstate(invalid="ignore", divide="ignore"):
        sin_beta = np.where(rho == 0, np.sin(beta1),
                            np.cos(ce) * np.sin(beta1) + yy * np.sin(ce) * np.cos(beta1) / rho)
    lam = np.arctan2(xx * np.sin(ce), rho * np.cos(beta1) * np.cos(ce) - yy * np.sin(beta1) * np.sin(ce))
    q = qp * sin_beta
    # authalic -> geodetic latitude: Newton on q(phi) (Snyder eq. 3-16)
    phi = np.arcsin(np.clip(q / 2, -1.0, 1.0))
    for _ in range(8):
        s, c = np.sin(phi), np.cos(phi)
        es = e * s
        with np.errstate(invalid="ignore", divide="ignore"):
            dphi = (1 - es * es) ** 2 / (2 * c) * (q / (1 - e * e) - s / (1 - es * es)
                                                   + (0.5 / e) * np.log((1 - es) / (1 + es)))
        phi = phi + np.where(np.isfinite(dphi), dphi, 0.0)
    return np.rad2deg(lam) + p.lon0, np.rad2deg(phi)


# ---------------------------------------------------------------------------
# generic
# ---------------------------------------------------------------------------
def forward(p: Proj, lon, lat):
    """geographic degrees -> CRS coordinates."""
    if p.kind == GEOGRAPHIC:
        return np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64)
    return {TMERC: tmerc_forward, WEBMERC: webmerc_forward, LAEA: laea_forward}[p.kind](p, lon, lat)


def inverse(p: Proj, x, y):
    """CRS coordinates -> geographic degrees."""
    if p.kind == GEOGRAPHIC:
        return np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    return {TMERC: tmerc_inverse, WEBMERC: webmerc_inverse, LAEA: laea_inverse}[p.kind](p, x, y)


def transform(src: Proj, dst: Proj, x, y):
    """``pyproj.Transformer.from_crs(src, dst, always_xy=True).transform(x, y)``."""
    lon, lat = inverse(src, x, y)
    return forward(dst, lon, lat)


def transform_bounds(src: Proj, dst: Proj, left, bottom, right, top, densify_pts=21):
    """``Transformer.transform_bounds``: densify each edge, transform, min / max.

    PROJ (proj_trans_bounds) puts ``densify_pts`` extra points on every edge (from memory --
    source not in container); its antimeridian / pole special cases for geographic output are
    not restated.  Only feeds integer window indices and the scale test of the reference
    (``reproject.py:347-352, 398-402``): parity unpinned beyond the reference's goldens.
    """
    n = densify_pts + 2
    t = np.linspace(0.0, 1.0, n)
    xs = np.concatenate([left + (right - left) * t, np.full(n, right), right + (left - right) * t, np.full(n, left)])
    ys = np.concatenate([np.full(n, bottom), bottom + (top - bottom) * t, np.full(n, top), top + (bottom - top) * t])
    tx, ty = transform(src, dst, xs, ys)
    return float(np.min(tx)), float(np.min(ty)), float(np.max(tx)), float(np.max(ty))

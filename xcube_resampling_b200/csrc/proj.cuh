// proj.cuh -- fp64 map-projection math for the reprojection kernels (sm_100a).
//
// The reference hands every CRS transform to pyproj.Transformer -> the PROJ C library
// (reproject.py:124,347,398,483; rectify.py:196-204; gridmapping/transform.py:77-91).  PROJ is not
// part of this build; the four projection families the path needs are evaluated here from their
// published closed forms, always in always_xy axis order:
//
//   geographic    degrees, identity
//   tmerc / UTM   Krueger n-series to n^6 (Karney 2011 eqs. 7-12, 35-36) -- the algorithm behind
//                 PROJ's default tmerc; series summed with Clenshaw recurrences
//   web Mercator  spherical formulas on the semi-major axis (EPSG:3857)
//   LAEA          Snyder (1987) pp. 187-190, ellipsoidal oblique / equatorial aspect (EPSG:3035)
//
// The auxiliary-latitude conversions (geodetic <-> conformal, authalic -> geodetic) are 6-term
// trigonometric series whose coefficients are NOT typed in from tables: make_proj_consts() derives
// them on the host from the exact closed forms by a discrete sine transform in long double, so
// they are exact to double rounding for any ellipsoid.
//
// Host side: make_proj_consts(xrs_proj) -> ProjC (plain struct passed to kernels by value).
// Device side: proj_inverse(ProjC, x, y) -> (lam, phi) radians; proj_forward(ProjC, lam, phi).
#pragma once

#include "common.cuh"

namespace xrs {

constexpr int PROJ_TERMS = 6;
constexpr double PROJ_PI = 3.14159265358979323846;
constexpr double PROJ_DEG2RAD = PROJ_PI / 180.0;
constexpr double PROJ_RAD2DEG = 180.0 / PROJ_PI;
constexpr double TMERC_ETA_MAX = 2.623395162778;  // PROJ's domain limit (|lon - lon0| ~ 90 deg)

struct ProjC {
    int kind;
    double a, e, es, one_es;
    double lon0;  // radians
    double fe, fn;
    // transverse Mercator
    double Qn;               // k0 * a * (A / a)
    double xi0;              // xi of the latitude of origin
    double alpha[PROJ_TERMS];  // Gaussian sphere -> projected plane
    double beta[PROJ_TERMS];   // projected plane -> Gaussian sphere
    double cbg[PROJ_TERMS];    // geodetic -> conformal latitude
    double cgb[PROJ_TERMS];    // conformal -> geodetic latitude
    // Lambert azimuthal equal-area
    double qp, rq, dd, sinb1, cosb1, lat0;
    double apa[PROJ_TERMS];    // authalic -> geodetic latitude
};

int make_proj_consts(const xrs_proj *p, ProjC *out);  // reproject.cu (host)

// ---------------------------------------------------------------------------
// device math
// ---------------------------------------------------------------------------
// sum_{k=1..6} c[k-1] * sin(2 k t), given s2 = sin 2t, c2 = cos 2t (Clenshaw)
__device__ __forceinline__ double clenshaw_sin(const double *c, double s2, double c2) {
    const double r = 2.0 * c2;
    double h1 = c[PROJ_TERMS - 1], h2 = 0.0;
#pragma unroll
    for (int k = PROJ_TERMS - 2; k >= 0; --k) {
        const double h = r * h1 - h2 + c[k];
        h2 = h1;
        h1 = h;
    }
    return s2 * h1;
}

// sum_k c[k-1] * sin(2 k (xi + i eta)) -> (d_xi, d_eta), complex Clenshaw
__device__ __forceinline__ void clenshaw_complex(const double *c, double sin2xi, double cos2xi, double sinh2eta,
                                                 double cosh2eta, double &d_xi, double &d_eta) {
    const double r = 2.0 * cos2xi * cosh2eta, i = -2.0 * sin2xi * sinh2eta;
    double hr1 = c[PROJ_TERMS - 1], hi1 = 0.0, hr2 = 0.0, hi2 = 0.0;
#pragma unroll
    for (int k = PROJ_TERMS - 2; k >= 0; --k) {
        const double hr = -hr2 + r * hr1 - i * hi1 + c[k];
        const double hi = -hi2 + i * hr1 + r * hi1;
        hr2 = hr1; hi2 = hi1;
        hr1 = hr; hi1 = hi;
    }
    const double sr = sin2xi * cosh2eta, si = cos2xi * sinh2eta;
    d_xi = sr * hr1 - si * hi1;
    d_eta = sr * hi1 + si * hr1;
}

// PROJ's adjlon: longitudes are brought back into [-pi, pi], but a value that overshoots pi by
// rounding only (x = -20037508.342789244 / a is -pi * (1 + 1 ulp)) is left alone, so that points ON
// the antimeridian do not flip sign (adjlon.c lets lon overshoot by 1e-12).
__device__ __forceinline__ double wrap_pi(double lam) {
    if (fabs(lam) > PROJ_PI + 1e-12) lam -= 2.0 * PROJ_PI * rint(lam / (2.0 * PROJ_PI));
    return lam;
}

__device__ __forceinline__ double laea_q(const ProjC &P, double sinphi) {
    const double es_ = P.e * sinphi;
    return P.one_es * (sinphi / (1.0 - es_ * es_) - (0.5 / P.e) * log((1.0 - es_) / (1.0 + es_)));
}

// CRS coordinates -> geographic (lam, phi) in radians.  false = outside the projection's domain.
__device__ __forceinline__ bool proj_inverse(const ProjC &P, double x, double y, double &lam, double &phi) {
    switch (P.kind) {
    case XRS_PROJ_GEOGRAPHIC:
        lam = x * PROJ_DEG2RAD;
        phi = y * PROJ_DEG2RAD;
        return true;
    case XRS_PROJ_WEBMERC:
        lam = wrap_pi(P.lon0 + (x - P.fe) / P.a);
        phi = atan(sinh((y - P.fn) / P.a));
        return true;
    case XRS_PROJ_TMERC: {
        const double xi = (y - P.fn) / P.Qn + P.xi0, eta = (x - P.fe) / P.Qn;
        if (!(fabs(eta) <= TMERC_ETA_MAX)) return false;
        double s2, c2;
        sincos(2.0 * xi, &s2, &c2);
        const double e2 = exp(2.0 * eta), ie2 = 1.0 / e2;
        double dxi, deta;
        clenshaw_complex(P.beta, s2, c2, 0.5 * (e2 - ie2), 0.5 * (e2 + ie2), dxi, deta);
        const double xip = xi - dxi, etap = eta - deta;
        double sx, cx;
        sincos(xip, &sx, &cx);
        const double sh = sinh(etap);
        lam = wrap_pi(P.lon0 + atan2(sh, cx));
        const double hyp = sqrt(sh * sh + cx * cx);  // cos(chi) * cosh(eta')
        const double chi = atan2(sx, hyp);
        const double inv = 1.0 / (sx * sx + hyp * hyp);
        phi = chi + clenshaw_sin(P.cgb, 2.0 * sx * hyp * inv, (hyp * hyp - sx * sx) * inv);
        return true;
    }
    case XRS_PROJ_LAEA: {
        const double xx = (x - P.fe) / P.dd, yy = (y - P.fn) * P.dd;
        const double rho = sqrt(xx * xx + yy * yy);
        if (rho < 1e-10) {
            lam = P.lon0;
            phi = P.lat0;
            return true;
        }
        const double sce = rho / (2.0 * P.rq);
        if (!(sce <= 1.0)) return false;
        const double sin_ce = 2.0 * sce * sqrt(1.0 - sce * sce), cos_ce = 1.0 - 2.0 * sce * sce;
        double sinb = cos_ce * P.sinb1 + yy * sin_ce * P.cosb1 / rho;
        sinb = fmin(1.0, fmax(-1.0, sinb));
        lam = wrap_pi(P.lon0 + atan2(xx * sin_ce, rho * P.cosb1 * cos_ce - yy * P.sinb1 * sin_ce));
        const double cosb = sqrt(1.0 - sinb * sinb);
        phi = asin(sinb) + clenshaw_sin(P.apa, 2.0 * sinb * cosb, 1.0 - 2.0 * sinb * sinb);
        return true;
    }
    }
    return false;
}

// geographic (lam, phi) radians -> CRS coordinates
__device__ __forceinline__ bool proj_forward(const ProjC &P, double lam, double phi, double &x, double &y) {
    switch (P.kind) {
    case XRS_PROJ_GEOGRAPHIC:
        x = lam * PROJ_RAD2DEG;
        y = phi * PROJ_RAD2DEG;
        return true;
    case XRS_PROJ_WEBMERC: {
        if (!(fabs(phi) < 0.5 * PROJ_PI)) return false;
        x = P.fe + P.a * wrap_pi(lam - P.lon0);
        y = P.fn + P.a * asinh(tan(phi));
        return true;
    }
    case XRS_PROJ_TMERC: {
        if (!(fabs(phi) <= 0.5 * PROJ_PI)) return false;
        const double dl = wrap_pi(lam - P.lon0);
        double s2, c2;
        sincos(2.0 * phi, &s2, &c2);
        const double chi = phi + clenshaw_sin(P.cbg, s2, c2);
        double sc, cc, sl, cl;
        sincos(chi, &sc, &cc);
        sincos(dl, &sl, &cl);
        const double ccl = cc * cl;
        const double xip = atan2(sc, ccl);
        const double inv = 1.0 / sqrt(sc * sc + ccl * ccl);  // cosh(eta')
        const double tan_ce = sl * cc * inv;                  // sinh(eta')
        const double etap = asinh(tan_ce);
        if (!(fabs(etap) <= TMERC_ETA_MAX)) return false;
        const double two_inv = 2.0 * inv, two_inv_sq = two_inv * inv, tmp = ccl * two_inv_sq;
        double dxi, deta;
        clenshaw_complex(P.alpha, sc * tmp, ccl * tmp - 1.0, tan_ce * two_inv, two_inv_sq - 1.0, dxi, deta);
        x = P.fe + P.Qn * (etap + deta);
        y = P.fn + P.Qn * (xip + dxi - P.xi0);
        return true;
    }
    case XRS_PROJ_LAEA: {
        if (!(fabs(phi) <= 0.5 * PROJ_PI)) return false;
        const double dl = wrap_pi(lam - P.lon0);
        double sl, cl;
        sincos(dl, &sl, &cl);
        double sinb = laea_q(P, sin(phi)) / P.qp;
        sinb = fmin(1.0, fmax(-1.0, sinb));
        const double cosb = sqrt(1.0 - sinb * sinb);
        double b = 1.0 + P.sinb1 * sinb + P.cosb1 * cosb * cl;
        if (!(b > 1e-10)) return false;  // antipode of the projection centre
        b = P.rq * sqrt(2.0 / b);
        x = P.fe + b * P.dd * cosb * sl;
        y = P.fn + (b / P.dd) * (P.cosb1 * sinb - P.sinb1 * cosb * cl);
        return true;
    }
    }
    return false;
}

// ---------------------------------------------------------------------------
// Separable forms for REGULAR grids.  On a regular grid in a transverse-Mercator CRS xi depends on
// the row only and eta on the column only; on a geographic or web-Mercator grid longitude and
// latitude themselves are separable.  The reprojection kernel evaluates the row-only and
// column-only transcendental functions once per tile row / column ("terms") and finishes each
// pixel with the "tail" below -- the same formulas as proj_inverse / proj_forward, regrouped.
// ---------------------------------------------------------------------------
struct Terms4 {
    double a, b, c, d;
};

// sin / cos and sinh / cosh of a small argument (|x| <= 0.1): Taylor series, truncation < 1e-22.
// The coefficients live in constant memory so that they are read as instruction operands instead
// of being rebuilt from immediates in every loop iteration.
static __constant__ double PROJ_INV_FACT[13] = {
    1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
    1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600};
__device__ __forceinline__ void sincos_small(double x, double &s, double &c) {
    const double z = -(x * x);
    const double *f = PROJ_INV_FACT;
    s = x * (1.0 + z * (f[3] + z * (f[5] + z * (f[7] + z * (f[9] + z * f[11])))));
    c = 1.0 + z * (f[2] + z * (f[4] + z * (f[6] + z * (f[8] + z * (f[10] + z * f[12])))));
}
__device__ __forceinline__ void sinhcosh_small(double x, double &s, double &c) {
    const double z = x * x;
    const double *f = PROJ_INV_FACT;
    s = x * (1.0 + z * (f[3] + z * (f[5] + z * (f[7] + z * (f[9] + z * f[11])))));
    c = 1.0 + z * (f[2] + z * (f[4] + z * (f[6] + z * (f[8] + z * (f[10] + z * f[12])))));
}

// transverse Mercator inverse: row terms (sin 2xi, cos 2xi, sin xi, cos xi) from y
__device__ __forceinline__ Terms4 tmerc_inv_row_terms(const ProjC &P, double y) {
    const double xi = (y - P.fn) / P.Qn + P.xi0;
    Terms4 t;
    sincos(xi, &t.c, &t.d);
    t.a = 2.0 * t.c * t.d;
    t.b = 1.0 - 2.0 * t.c * t.c;
    return t;
}
// column terms (sinh 2eta, cosh 2eta, sinh eta, cosh eta) from x; NaN outside the domain
__device__ __forceinline__ Terms4 tmerc_inv_col_terms(const ProjC &P, double x) {
    const double eta = (x - P.fe) / P.Qn;
    Terms4 t;
    if (!(fabs(eta) <= TMERC_ETA_MAX)) {
        t.a = t.b = t.c = t.d = NAN;
        return t;
    }
    const double e1 = exp(eta), i1 = 1.0 / e1;
    t.c = fabs(eta) < 1e-3 ? eta * (1.0 + eta * eta * (1.0 / 6)) : 0.5 * (e1 - i1);
    t.d = 0.5 * (e1 + i1);
    t.a = 2.0 * t.c * t.d;
    t.b = 1.0 + 2.0 * t.c * t.c;
    return t;
}
// per pixel: (lam, phi) from the terms.  false = corrections too large for the small-angle forms
// (far outside any sensible zone): the caller evaluates proj_inverse instead.
__device__ __forceinline__ bool tmerc_inv_tail(const ProjC &P, const Terms4 &r, const Terms4 &c, double &lam,
                                               double &phi) {
    double dxi, deta;
    clenshaw_complex(P.beta, r.a, r.b, c.a, c.b, dxi, deta);
    if (c.a != c.a) {  // column outside the projection's domain
        lam = phi = NAN;
        return true;
    }
    if (!(fabs(dxi) <= 0.1 && fabs(deta) <= 0.1)) return false;
    double sd, cd, shd, chd;
    sincos_small(dxi, sd, cd);
    sinhcosh_small(deta, shd, chd);
    const double sx = r.c * cd - r.d * sd, cx = r.d * cd + r.c * sd;  // sin, cos of xi' = xi - dxi
    const double sh = c.c * chd - c.d * shd;                          // sinh of eta' = eta - deta
    if (!(cx > 1e-9)) return false;  // at or beyond a pole: exact path
    lam = wrap_pi(P.lon0 + atan(sh / cx));
    const double hyp = sqrt(sh * sh + cx * cx);
    const double chi = atan(sx / hyp);
    const double inv = 1.0 / (sx * sx + hyp * hyp);
    phi = chi + clenshaw_sin(P.cgb, 2.0 * sx * hyp * inv, (hyp * hyp - sx * sx) * inv);
    return true;
}

// forward projections from separable geographic coordinates: row terms from phi, column terms
// from lam, per-pixel tail.  kind GEOGRAPHIC / WEBMERC are fully separable (a = coordinate).
__device__ __forceinline__ Terms4 fwd_row_terms(const ProjC &P, double phi, bool ok) {
    Terms4 t;
    t.a = t.b = t.c = t.d = NAN;
    if (!ok || !(fabs(phi) <= 0.5 * PROJ_PI)) return t;
    switch (P.kind) {
    case XRS_PROJ_GEOGRAPHIC: t.a = phi * PROJ_RAD2DEG; break;
    case XRS_PROJ_WEBMERC:
        if (fabs(phi) < 0.5 * PROJ_PI) t.a = P.fn + P.a * asinh(tan(phi));
        break;
    case XRS_PROJ_TMERC: {
        double s2, c2;
        sincos(2.0 * phi, &s2, &c2);
        sincos(phi + clenshaw_sin(P.cbg, s2, c2), &t.a, &t.b);  // sin, cos of the conformal latitude
        break;
    }
    case XRS_PROJ_LAEA: {
        double sinb = laea_q(P, sin(phi)) / P.qp;
        sinb = fmin(1.0, fmax(-1.0, sinb));
        t.a = sinb;
        t.b = sqrt(1.0 - sinb * sinb);
        break;
    }
    }
    return t;
}
__device__ __forceinline__ Terms4 fwd_col_terms(const ProjC &P, double lam) {
    Terms4 t;
    t.a = t.b = t.c = t.d = NAN;
    switch (P.kind) {
    case XRS_PROJ_GEOGRAPHIC: t.a = lam * PROJ_RAD2DEG; break;
    case XRS_PROJ_WEBMERC: t.a = P.fe + P.a * wrap_pi(lam - P.lon0); break;
    case XRS_PROJ_TMERC:
    case XRS_PROJ_LAEA: sincos(wrap_pi(lam - P.lon0), &t.a, &t.b); break;
    }
    return t;
}
__device__ __forceinline__ void fwd_tail(const ProjC &P, const Terms4 &r, const Terms4 &c, double &x, double &y) {
    switch (P.kind) {
    case XRS_PROJ_GEOGRAPHIC:
    case XRS_PROJ_WEBMERC:
        x = c.a;
        y = r.a;
        return;
    case XRS_PROJ_TMERC: {
        const double sc = r.a, cc = r.b, sl = c.a, cl = c.b;
        const double ccl = cc * cl;
        const double xip = atan2(sc, ccl);
        const double inv = 1.0 / sqrt(sc * sc + ccl * ccl);
        const double tan_ce = sl * cc * inv;
        const double etap = asinh(tan_ce);
        if (!(fabs(etap) <= TMERC_ETA_MAX)) {
            x = y = NAN;
            return;
        }
        const double two_inv = 2.0 * inv, two_inv_sq = two_inv * inv, tmp = ccl * two_inv_sq;
        double dxi, deta;
        clenshaw_complex(P.alpha, sc * tmp, ccl * tmp - 1.0, tan_ce * two_inv, two_inv_sq - 1.0, dxi, deta);
        x = P.fe + P.Qn * (etap + deta);
        y = P.fn + P.Qn * (xip + dxi - P.xi0);
        return;
    }
    case XRS_PROJ_LAEA: {
        const double sinb = r.a, cosb = r.b, sl = c.a, cl = c.b;
        double b = 1.0 + P.sinb1 * sinb + P.cosb1 * cosb * cl;
        if (!(b > 1e-10)) {
            x = y = NAN;
            return;
        }
        b = P.rq * sqrt(2.0 / b);
        x = P.fe + b * P.dd * cosb * sl;
        y = P.fn + (b / P.dd) * (P.cosb1 * sinb - P.sinb1 * cosb * cl);
        return;
    }
    }
    x = y = NAN;
}

// pyproj.Transformer.from_crs(from, to, always_xy=True).transform(x, y); NaN when not transformable
// (PROJ reports inf there).  Two geographic CRSs pass the coordinates through untouched (the datum
// shift WGS84 <-> ETRS89 is PROJ's "ballpark" identity).
__device__ __forceinline__ void proj_transform(const ProjC &from, const ProjC &to, double x, double y, double &ox,
                                               double &oy) {
    if (from.kind == XRS_PROJ_GEOGRAPHIC && to.kind == XRS_PROJ_GEOGRAPHIC) {
        ox = x;
        oy = y;
        return;
    }
    double lam, phi;
    bool ok = proj_inverse(from, x, y, lam, phi);
    if (ok && from.kind == XRS_PROJ_GEOGRAPHIC && !(fabs(y) <= 90.0)) ok = false;
    ok = ok && proj_forward(to, lam, phi, ox, oy);
    if (!ok) ox = oy = NAN;
}

}  // namespace xrs

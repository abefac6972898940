// rectify_ij.cu -- K1: the source-index (ij) image of a regular target grid.
//
//   xrs_rectify_ij   rectify.py:312-576 (_compute_target_source_ij*), 737-768 (_fdet/_fu/_fv/_fclamp)
//
// The reference scatters every source quad's two triangles onto the target tile sequentially and
// lets the FIRST writer win.  Whether a quad accepts a pixel never depends on earlier writes, so
// the result equals "per pixel, the accepting quad with the smallest row-major index wins".  That
// form is order-free and runs as four kernels:
//
//   k1_init_claims   claim words <- "none"; source quad rows visible to the requested target rows.
//   k1_scatter       one lane per source quad, in target pixel space.  A warp marches down a strip
//                    of 31 quad columns (vertex rows kept in registers / prefetched, every
//                    coordinate loaded once, coalesced); the quad's pixel box is scanned with
//                    stepped edge functions and sign-bit decisions; accepted pixels receive
//                    atomicMin(claim, 2 * quad index + triangle).
//   k1_scatter_slow  the generic path in the reference's exact arithmetic for queued quads
//                    (non-finite vertices, degenerate or huge quads, pixels within rounding
//                    distance of an edge).
//   k1_resolve       one thread per target pixel: the winning triangle recomputes its barycentric
//                    coordinates with the reference's exact expressions and writes (i, j).
//
// All parity-critical arithmetic uses round-to-nearest intrinsics in the reference's expression
// order (no FMA contraction) and the reference TILE-LOCAL offsets (rectify.py:402-416), so the ij
// image is bit-identical to the numba kernel's for any reference tile size.
//
// (History: a gather form -- CTA per target tile, source window staged in shared memory by bulk
// copies -- measured 3.0 ms on the OLCI scene; a CRS-space scatter 2.1 ms; this form 1.45 ms.)
#include "rectify_common.cuh"

namespace xrs {

// np.floor(v).astype(np.int64) reduced to int32: non-finite / out-of-range -> sentinel
// (x86 gives INT64_MIN for those), everything else clamped to +-2^30, which preserves
// every comparison against pixel ranges.
__device__ __forceinline__ int floor_px(double v) {
    const double f = floor(v);
    if (!(f >= -9223372036854775808.0 && f < 9223372036854775808.0)) return K1_SENTINEL;
    return static_cast<int>(fmin(fmax(f, -1073741824.0), 1073741824.0));
}

// floor_px(num / den) where the quotient is first approximated by num * fl(1/den) (error
// < 4e-16 relative): only when that lands within 1e-9 (relative) of an integer can the floor
// differ from the exactly rounded quotient's, and only then is the division carried out.
__device__ __forceinline__ int floor_px_div(double num, double den, double inv_den) {
    double q = num * inv_den;
    if (fabs(q - rint(q)) <= 1e-9 * fmax(1.0, fabs(q))) q = ddiv(num, den);
    return floor_px(q);
}

// rectify.py:558-573 acceptance of one triangle, exact form.
__device__ __forceinline__ bool tri_accepts_exact(double nu, double nv, double det, double lo, double hi) {
    const double u = ddiv(nu, det), v = ddiv(nv, det);
    return u >= lo && v >= lo && dadd(u, v) <= hi;
}

// The same decision without divisions.  With s = sign(det), ad = |det| the three conditions
// u >= lo, v >= lo, u + v <= hi become half-plane tests s*nu >= lo*ad, s*nv >= lo*ad,
// s*(nu+nv) <= hi*ad.  The rounded quotients the reference compares differ from the real ones
// by < 1e-15 (relative to ad) whenever the outcome is not already decided by another condition,
// so outside a band of 1e-12*ad around the thresholds the decision is certain; inside the band
// (and for NaN / infinite operands, where every comparison below is false) the exact form runs.
struct TriTest {
    double det, sgn, t_lo_m, t_lo_p, t_hi_m, t_hi_p;
    bool live;
};
__device__ __forceinline__ TriTest make_tri_test(double det, double lo, double hi) {
    TriTest t;
    t.det = det;
    t.live = det != 0.0;
    t.sgn = det < 0.0 ? -1.0 : 1.0;
    const double ad = fabs(det), m = 1e-12 * ad;
    t.t_lo_m = lo * ad - m; t.t_lo_p = lo * ad + m;
    t.t_hi_m = hi * ad - m; t.t_hi_p = hi * ad + m;
    return t;
}
__device__ __forceinline__ bool tri_accepts(const TriTest &t, double nu, double nv, double lo, double hi) {
    if (!t.live) return false;
    const double a = t.sgn * nu, b = t.sgn * nv, c = a + b;
    if (a < t.t_lo_m || b < t.t_lo_m || c > t.t_hi_p) return false;
    if (a > t.t_lo_p && b > t.t_lo_p && c < t.t_hi_m) return true;
    return tri_accepts_exact(nu, nv, t.det, lo, hi);
}

// reference tile geometry a lane keeps cached while its quads stay in the same tile
struct TileCtx {
    int id;                    // ty * ntx + tx, -1 = nothing cached
    int r0, c0, th, tw;        // tile origin and clipped size
    int dj_lo, dj_hi;          // tile-local rows inside [row_begin, row_end)
    int qi_lo, qi_hi, qj_lo, qj_hi;  // quads inside the tile's source window (rectify.py:397-399)
    double x_off, y_off;
    bool has_window;
};

__device__ __forceinline__ void load_tile_ctx(const IjGeom &g, int ty, int tx, TileCtx &t) {
    t.id = ty * g.ntx + tx;
    t.r0 = ty * g.tile_h; t.c0 = tx * g.tile_w;
    t.th = static_cast<int>(min(static_cast<int64_t>(g.tile_h), g.dst_h - t.r0));
    t.tw = static_cast<int>(min(static_cast<int64_t>(g.tile_w), g.dst_w - t.c0));
    t.dj_lo = static_cast<int>(max(int64_t(0), g.row_begin - t.r0));
    t.dj_hi = static_cast<int>(min(static_cast<int64_t>(t.th), g.row_end - t.r0)) - 1;
    const int64_t *bb = g.tile_boxes + 4 * static_cast<int64_t>(t.id);
    const int64_t b0 = __ldg(bb), b1 = __ldg(bb + 1), b2 = __ldg(bb + 2), b3 = __ldg(bb + 3);
    t.has_window = b0 != -1;
    t.qi_lo = static_cast<int>(b0); t.qj_lo = static_cast<int>(b1);
    t.qi_hi = static_cast<int>(min(b2 + 1, g.src_w)) - 2;  // vertex slice [b0, min(b2+1, w)) -> last quad column
    t.qj_hi = static_cast<int>(min(b3 + 1, g.src_h)) - 2;
    // rectify.py:402-406
    t.x_off = dadd(g.x_min, dmul(static_cast<double>(t.c0), g.x_res));
    t.y_off = g.j_up ? dadd(g.y_min, dmul(static_cast<double>(t.r0), g.y_res))
                     : dsub(g.y_max, dmul(static_cast<double>(t.r0), g.y_res));
}

// Claim words <- "no claim"; block 0 also resets the slow-quad counter and reduces the source quad
// rows the requested target rows can see: the union of the source windows of the reference tiles
// that intersect [row_begin, row_end) (rectify.py:397-399 -- a tile only ever looks at the quads of
// its own window).  k1_scatter skips every quad row outside that range, which is what makes a
// row-band call cost its footprint instead of the whole swath.
__global__ void k1_init_claims(uint4 *claims, int64_t n_vec, unsigned int *slow_count, IjGeom g) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n_vec) claims[i] = make_uint4(K1_NOCLAIM, K1_NOCLAIM, K1_NOCLAIM, K1_NOCLAIM);
    if (blockIdx.x != 0) return;
    int *range = reinterpret_cast<int *>(slow_count) + 1;  // [qj_min, qj_max]
    if (threadIdx.x == 0) {
        *slow_count = 0u;
        range[0] = INT32_MAX;
        range[1] = -1;
    }
    __syncthreads();
    const int ty0 = static_cast<int>(g.row_begin / g.tile_h), ty1 = static_cast<int>((g.row_end - 1) / g.tile_h);
    int lo = INT32_MAX, hi = -1;
    for (int t = ty0 * g.ntx + threadIdx.x; t < (ty1 + 1) * g.ntx; t += blockDim.x) {
        const int64_t *bb = g.tile_boxes + 4 * static_cast<int64_t>(t);
        if (__ldg(bb) == -1) continue;
        lo = min(lo, static_cast<int>(__ldg(bb + 1)));
        hi = max(hi, static_cast<int>(min(__ldg(bb + 3) + 1, g.src_h)) - 2);
    }
    if (lo <= hi) {
        atomicMin(range, lo);
        atomicMax(range + 1, hi);
    }
    // ... and, per block of K1S_ROWS quad rows, the quad COLUMNS those windows cover: for a rotated
    // swath a target row band maps to a diagonal strip of the source, so most of every source row
    // lies outside it
    int *col_range = range + 3;  // after [qj_min, qj_max, pad]
    const int n_rb = static_cast<int>((g.src_h - 1 + K1S_ROWS - 1) / K1S_ROWS);
    for (int rb = threadIdx.x; rb < n_rb; rb += blockDim.x) {
        const int rb_lo = rb * K1S_ROWS, rb_hi = rb_lo + K1S_ROWS - 1;
        int clo = INT32_MAX, chi = -1;
        for (int t = ty0 * g.ntx; t < (ty1 + 1) * g.ntx; ++t) {
            const int64_t *bb = g.tile_boxes + 4 * static_cast<int64_t>(t);
            const int64_t b0 = __ldg(bb);
            if (b0 == -1) continue;
            const int qj_lo = static_cast<int>(__ldg(bb + 1));
            const int qj_hi = static_cast<int>(min(__ldg(bb + 3) + 1, g.src_h)) - 2;
            if (qj_lo > rb_hi || qj_hi < rb_lo) continue;
            clo = min(clo, static_cast<int>(b0));
            chi = max(chi, static_cast<int>(min(__ldg(bb + 2) + 1, g.src_w)) - 2);
        }
        if (g.fp_cols) {  // caller-supplied footprint of the row band (min-form: c_min, -c_max)
            clo = max(clo, static_cast<int>(__ldg(g.fp_cols + 2 * rb)));
            const int neg_hi = __ldg(g.fp_cols + 2 * rb + 1);
            chi = neg_hi == INT32_MAX ? -1 : min(chi, -neg_hi);
        }
        col_range[2 * rb] = clo;
        col_range[2 * rb + 1] = chi;
    }
}

// ---------------------------------------------------------------------------
// k1_scatter
// ---------------------------------------------------------------------------
struct ScatterConst {
    double x_scale, y_scale, inv_xs, inv_ys, uv_lo, uv_hi;
};

// rectify.py:516-576 for one quad inside one reference tile, given the quad's tile-local pixel
// box [i_lo, i_hi] x [j_lo, j_hi] (already clipped to the tile and to the requested rows).
__device__ __forceinline__ void claim_pixels(const IjGeom &g, const TileCtx &tc, const ScatterConst &k, double x0,
                                             double y0, double x1, double y1, double x2, double y2, double x3,
                                             double y3, int i_lo, int i_hi, int j_lo, int j_hi, uint32_t qkey) {
    // rectify.py:528-542
    double det_a = tri_det(x0, y0, x1, y1, x2, y2);
    if (det_a != det_a) det_a = 0.0;
    double det_b = tri_det(x3, y3, x2, y2, x1, y1);
    if (det_b != det_b) det_b = 0.0;
    if (det_a == 0.0 && det_b == 0.0) return;
    const TriTest ta = make_tri_test(det_a, k.uv_lo, k.uv_hi), tb = make_tri_test(det_b, k.uv_lo, k.uv_hi);
    uint32_t *claim_row = g.claims + (static_cast<int64_t>(tc.r0) + j_lo - g.row_begin) * g.dst_w + tc.c0;
    for (int dj = j_lo; dj <= j_hi; ++dj, claim_row += g.dst_w) {
        const double py = dadd(tc.y_off, dmul(dadd(static_cast<double>(dj), 0.5), k.y_scale));
        for (int di = i_lo; di <= i_hi; ++di) {
            const double px = dadd(tc.x_off, dmul(dadd(static_cast<double>(di), 0.5), k.x_scale));
            const bool acc_a = tri_accepts(ta, tri_u(px, py, x0, y0, x2, y2), tri_v(px, py, x0, y0, x1, y1), k.uv_lo, k.uv_hi);
            const bool acc_b = !acc_a && tri_accepts(tb, tri_u(px, py, x3, y3, x1, y1), tri_v(px, py, x3, y3, x2, y2),
                                                     k.uv_lo, k.uv_hi);
            // claim word: quad index and, in bit 0, which triangle accepted (A is tried first)
            if (acc_a || acc_b) atomicMin(claim_row + di, (qkey << 1) | (acc_b ? 1u : 0u));
        }
    }
}

// Quads near a reference-tile border, with non-finite vertices or reaching outside the target:
// visit every tile a conservative pixel box touches and redo the tile-local arithmetic there.
__device__ __forceinline__ void scatter_quad_generic(const IjGeom &g, const ScatterConst &k, double x0, double y0,
                                                  double x1, double y1, double x2, double y2, double x3, double y3,
                                                  int qi, int qj, uint32_t qkey) {
    const double inv_xr = 1.0 / g.x_res, inv_yr = 1.0 / g.y_res;
    const double W = static_cast<double>(g.dst_w);
    const double R0 = static_cast<double>(g.row_begin), R1 = static_cast<double>(g.row_end);
    const double fx0 = (x0 - g.x_min) * inv_xr, fx1 = (x1 - g.x_min) * inv_xr;
    const double fx2 = (x2 - g.x_min) * inv_xr, fx3 = (x3 - g.x_min) * inv_xr;
    const double fy0 = g.j_up ? (y0 - g.y_min) * inv_yr : (g.y_max - y0) * inv_yr;
    const double fy1 = g.j_up ? (y1 - g.y_min) * inv_yr : (g.y_max - y1) * inv_yr;
    const double fy2 = g.j_up ? (y2 - g.y_min) * inv_yr : (g.y_max - y2) * inv_yr;
    const double fy3 = g.j_up ? (y3 - g.y_min) * inv_yr : (g.y_max - y3) * inv_yr;
    const bool f0 = isfinite(fx0) && isfinite(fy0), f1 = isfinite(fx1) && isfinite(fy1);
    const bool f2 = isfinite(fx2) && isfinite(fy2), f3 = isfinite(fx3) && isfinite(fy3);
    double lo_x = INFINITY, hi_x = -INFINITY, lo_y = INFINITY, hi_y = -INFINITY;
    if (f0 && f1 && f2 && f3) {
        // the reference scans [min floor(p), max floor(p)] (rectify.py:500-526); +-1 px absorbs the
        // difference between this global and the tile-local pixel arithmetic
        lo_x = fmin(fmin(fx0, fx1), fmin(fx2, fx3)) - 1.0; hi_x = fmax(fmax(fx0, fx1), fmax(fx2, fx3)) + 1.0;
        lo_y = fmin(fmin(fy0, fy1), fmin(fy2, fy3)) - 1.0; hi_y = fmax(fmax(fy0, fy1), fmax(fy2, fy3)) + 1.0;
    } else {
        // A triangle with a non-finite vertex never accepts a pixel (its u or v is NaN): only
        // all-finite triangles count; their acceptance region is the triangle grown by the uv
        // tolerance, bounded by a few % of its extent + 1 px.
        if (f0 && f1 && f2) {
            lo_x = fmin(fx0, fmin(fx1, fx2)); hi_x = fmax(fx0, fmax(fx1, fx2));
            lo_y = fmin(fy0, fmin(fy1, fy2)); hi_y = fmax(fy0, fmax(fy1, fy2));
        }
        if (f3 && f2 && f1) {
            lo_x = fmin(lo_x, fmin(fx3, fmin(fx1, fx2))); hi_x = fmax(hi_x, fmax(fx3, fmax(fx1, fx2)));
            lo_y = fmin(lo_y, fmin(fy3, fmin(fy1, fy2))); hi_y = fmax(hi_y, fmax(fy3, fmax(fy1, fy2)));
        }
        const double margin_k = 0.01 + 4.0 * g.uv_delta;
        const double mx = margin_k * (hi_x - lo_x) + 1.0, my = margin_k * (hi_y - lo_y) + 1.0;
        lo_x -= mx; hi_x += mx; lo_y -= my; hi_y += my;
    }
    if (!(lo_x <= hi_x) || hi_x < 0.0 || hi_y < R0 || lo_x >= W || lo_y >= R1) return;
    const int gx0 = static_cast<int>(fmax(lo_x, 0.0)), gx1 = static_cast<int>(fmin(hi_x, W - 1.0));
    const int gy0 = static_cast<int>(fmax(lo_y, R0)), gy1 = static_cast<int>(fmin(hi_y, R1 - 1.0));
    const int tx_a = gx0 / g.tile_w, tx_b = gx1 / g.tile_w, ty_a = gy0 / g.tile_h, ty_b = gy1 / g.tile_h;
    TileCtx tc;
    for (int ty = ty_a; ty <= ty_b; ++ty)
        for (int tx = tx_a; tx <= tx_b; ++tx) {
            load_tile_ctx(g, ty, tx, tc);
            if (!tc.has_window || qi < tc.qi_lo || qi > tc.qi_hi || qj < tc.qj_lo || qj > tc.qj_hi) continue;
            // rectify.py:500-526: tile-local pixel box of the quad
            const int pi0 = floor_px_div(dsub(x0, tc.x_off), k.x_scale, k.inv_xs);
            const int pi1 = floor_px_div(dsub(x1, tc.x_off), k.x_scale, k.inv_xs);
            const int pi2 = floor_px_div(dsub(x2, tc.x_off), k.x_scale, k.inv_xs);
            const int pi3 = floor_px_div(dsub(x3, tc.x_off), k.x_scale, k.inv_xs);
            const int pj0 = floor_px_div(dsub(y0, tc.y_off), k.y_scale, k.inv_ys);
            const int pj1 = floor_px_div(dsub(y1, tc.y_off), k.y_scale, k.inv_ys);
            const int pj2 = floor_px_div(dsub(y2, tc.y_off), k.y_scale, k.inv_ys);
            const int pj3 = floor_px_div(dsub(y3, tc.y_off), k.y_scale, k.inv_ys);
            int i_lo = min(min(pi0, pi1), min(pi2, pi3)), i_hi = max(max(pi0, pi1), max(pi2, pi3));
            int j_lo = min(min(pj0, pj1), min(pj2, pj3)), j_hi = max(max(pj0, pj1), max(pj2, pj3));
            if (i_hi < 0 || j_hi < 0 || i_lo >= tc.tw || j_lo >= tc.th) continue;
            i_lo = max(i_lo, 0); i_hi = min(i_hi, tc.tw - 1);
            j_lo = max(j_lo, tc.dj_lo); j_hi = min(j_hi, tc.dj_hi);  // tile clip + requested rows
            if (j_lo > j_hi) continue;
            claim_pixels(g, tc, k, x0, y0, x1, y1, x2, y2, x3, y3, i_lo, i_hi, j_lo, j_hi, qkey);
        }
}

// ---------------------------------------------------------------------------
// k1_scatter, pixel-space form
//
// Barycentric coordinates are invariant under the affine map CRS -> target pixel coordinates, so
// the acceptance conditions u >= lo, v >= lo, u + v <= hi can be evaluated with the vertices
// expressed in (fractional) GLOBAL pixel coordinates f = (x - x_min) / x_res: no tile-local
// offsets, no exact floor divisions, and the edge functions step by plain vertex differences.
// The reference's own arithmetic (tile-local offsets, CRS units, divided form) differs from this
// by rounding only (< 1e-11 relative to |det|); a pixel whose edge functions do not all clear a
// margin well above that sends its quad to the generic kernel, which IS the reference's arithmetic.  The
// reference's pixel box needs no separate test: a pixel centre inside the triangles grown by the
// uv tolerance lies inside [floor(min f), floor(max f)] as long as uv_delta * extent < 0.5 px
// (quads larger than K1_MAX_EXTENT pixels take the generic path).
// ---------------------------------------------------------------------------
constexpr double K1_MAX_EXTENT = 64.0;
constexpr double K1_TIGHT_EXTENT = 8.0;  // quads up to this many pixels get the centre-tight candidate box

struct TileWin {
    int id;                          // ty * ntx + tx, -1 = nothing cached
    int qi_lo, qi_hi, qj_lo, qj_hi;  // quads inside the tile's source window (rectify.py:397-399)
    bool has;
};

__device__ __forceinline__ void load_tile_win(const IjGeom &g, int ty, int tx, TileWin &t) {
    t.id = ty * g.ntx + tx;
    const int64_t *bb = g.tile_boxes + 4 * static_cast<int64_t>(t.id);
    const int64_t b0 = __ldg(bb), b1 = __ldg(bb + 1), b2 = __ldg(bb + 2), b3 = __ldg(bb + 3);
    t.has = b0 != -1;
    t.qi_lo = static_cast<int>(b0); t.qj_lo = static_cast<int>(b1);
    t.qi_hi = static_cast<int>(min(b2 + 1, g.src_w)) - 2;
    t.qj_hi = static_cast<int>(min(b3 + 1, g.src_h)) - 2;
}

// Edge functions of one triangle in pixel space, referred to the centre of pixel (i_ref, j_ref):
//   e1 = s*nu - lo*|det| - m,  e2 = s*nv - lo*|det| - m  (m = decision margin), stepped per pixel;
//   the third condition hi*|det| - s*(nu + nv) - m > 0 is e1 + e2 < k3.
struct PxEdges {
    double e1, e2;
    double dx1, dx2;  // per pixel column
    double dy1, dy2;  // per pixel row
    double k3, two_m;
    bool degenerate;
};

// (ax, ay) is the triangle's origin vertex, (bx, by) the u-direction vertex, (cx, cy) the v-direction
// vertex (_fdet / _fu / _fv of rectify.py:737-757 with pixel coordinates).
__device__ __forceinline__ PxEdges make_px_edges(double ax, double ay, double bx, double by, double cx, double cy,
                                                 double pcx, double pcy, double lo, double hi, double coord_bound) {
    PxEdges f;
    const double abx = ax - bx, aby = ay - by, acx = ax - cx, acy = ay - cy, apx = ax - pcx, apy = ay - pcy;
    const double det = abx * acy - acx * aby;
    const double nu = apx * acy - apy * acx;
    const double nv = apy * abx - apx * aby;
    const double s = det < 0.0 ? -1.0 : 1.0, ad = fabs(det);
    f.dx1 = -s * acy; f.dy1 = s * acx;   // gradient of s*nu
    f.dx2 = s * aby;  f.dy2 = -s * abx;  // gradient of s*nv
    const double m = 1e-9 * ad + coord_bound * (fabs(acy) + fabs(acx) + fabs(aby) + fabs(abx));
    f.e1 = s * nu - lo * ad - m;
    f.e2 = s * nv - lo * ad - m;
    f.k3 = (hi - 2.0 * lo) * ad - 3.0 * m;
    f.two_m = 2.0 * m;
    // (nearly) collapsed triangle, in absolute terms or as a sliver of its own edges: the reference's
    // u = nu / det is then dominated by rounding, so the generic kernel (the reference's arithmetic) decides
    const double edge_sum = fabs(acy) + fabs(acx) + fabs(aby) + fabs(abx);
    f.degenerate = !(ad > 1e-6) || ad < 1e-4 * edge_sum * edge_sum;
    return f;
}

// The same verdict from the vertices alone (quads that need no edge functions).
__device__ __forceinline__ bool px_degenerate(double ax, double ay, double bx, double by, double cx, double cy) {
    const double abx = ax - bx, aby = ay - by, acx = ax - cx, acy = ay - cy;
    const double ad = fabs(abx * acy - acx * aby);
    const double edge_sum = fabs(acy) + fabs(acx) + fabs(aby) + fabs(abx);
    return !(ad > 1e-6) || ad < 1e-4 * edge_sum * edge_sum;
}

// The per-pixel decisions read only the SIGN / high word of the three condition values
// (a1, a2, k3 - (a1 + a2)), i.e. integer instructions instead of fp64 compares:
//   certain accept  all three non-negative (their high words OR-ed have the sign bit clear);
//   certain reject  one of them below -4m: for negative doubles the high word, read as unsigned,
//                   grows with the magnitude, so `hi > hi(-4m)` is that test to within 2^-20 relative,
//                   which the factor 2 between the 2m the argument needs and the 4m used here absorbs.
__device__ __forceinline__ bool px_accepts(double a1, double a2, double t3) {
    return (__double2hiint(a1) | __double2hiint(a2) | __double2hiint(t3)) >= 0;
}
__device__ __forceinline__ bool px_rejects(uint32_t neg_thresh_hi, double a1, double a2, double t3) {
    return static_cast<uint32_t>(__double2hiint(a1)) > neg_thresh_hi ||
           static_cast<uint32_t>(__double2hiint(a2)) > neg_thresh_hi ||
           static_cast<uint32_t>(__double2hiint(t3)) > neg_thresh_hi;
}

#ifndef XRS_K1_MINBLOCKS
#define XRS_K1_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(K1S_WARPS * 32, XRS_K1_MINBLOCKS) k1_scatter(const __grid_constant__ IjGeom g) {
    const int64_t nqi = g.src_w - 1, nqj = g.src_h - 1;
    const int lane = threadIdx.x & 31;
    const int64_t strip = static_cast<int64_t>(blockIdx.x) * K1S_WARPS + (threadIdx.x >> 5);
    if (strip * 31 >= nqi) return;  // whole warp out of range
    const int64_t col = strip * 31 + lane;  // vertex column of this lane
    const int *qj_range = reinterpret_cast<const int *>(g.slow_count) + 1;  // from k1_init_claims
    const int64_t j_begin = max(static_cast<int64_t>(blockIdx.y) * K1S_ROWS, static_cast<int64_t>(qj_range[0]));
    const int64_t j_end = min(min(static_cast<int64_t>(blockIdx.y + 1) * K1S_ROWS, nqj),
                              static_cast<int64_t>(qj_range[1]) + 1);  // quad rows [j_begin, j_end)
    if (j_begin >= j_end) return;
    // quad columns visible to the requested target rows in this block of quad rows: quads outside
    // lie in no source window of the band's tiles (and outside the caller's footprint, whose
    // coordinates need not even be resident), so neither they nor their vertices are touched
    const int *col_range = qj_range + 3 + 2 * blockIdx.y;
    const int64_t c_lo = col_range[0], c_hi = col_range[1];
    if (strip * 31 + 30 < c_lo || strip * 31 > c_hi) return;  // warp-uniform
    const bool col_ok = col < g.src_w && col >= c_lo && col <= c_hi + 1;
    const bool quad_lane = lane < 31 && col < nqi && col >= c_lo && col <= c_hi;

    const double inv_xr = 1.0 / g.x_res, inv_yr = 1.0 / g.y_res;
    const double uv_lo = -g.uv_delta, uv_hi = dadd(1.0, dmul(2.0, g.uv_delta));
    const int W = static_cast<int>(g.dst_w), R0 = static_cast<int>(g.row_begin), R1 = static_cast<int>(g.row_end);
    const double clamp_hi = static_cast<double>(max(g.dst_w, g.dst_h)) + 8.0;
    // rounding of a pixel coordinate seen through an edge function's gradient, with a factor 8 to
    // spare.  Two sources: this kernel's own pixel coordinates (two roundings of a value up to the
    // image size) and the reference's pixel centres x_off + (i + 0.5) * res (rectify.py:516-523), whose
    // rounding error is an ulp of the CRS coordinate itself -- |x| / res pixels, which for metre grids
    // (UTM northings of 5e6 m at 10 m) is far larger than the image size.
    const double crs_mag = fmax(fmax(fabs(g.x_min), fabs(g.x_min + static_cast<double>(g.dst_w) * g.x_res)) * inv_xr,
                                fmax(fabs(g.y_min), fabs(g.y_max)) * inv_yr);
    const double coord_bound = 8.0 * 2.3e-16 * fmax(clamp_hi, crs_mag);
    const bool small_tolerance = g.uv_delta * (K1_MAX_EXTENT + 2.0) < 0.25;
    const int qi = static_cast<int>(col);

    TileWin tw;
    tw.id = -1;
    // A vertex in pixel space: fractional coordinates and their floors (pixel indices, clamped to a
    // band around the image so that they fit an int).  `ok` is false for non-finite coordinates.
    // Besides the floor, a vertex carries two flags per axis (packed below the index: (floor << 2) | up << 1 | dn):
    // `up`: its fraction lies above 0.5 + t, so no pixel CENTRE of column `floor` is at or beyond it;
    // `dn`: its fraction lies below 0.5 - t, the same towards the other side.  A pixel is accepted only if
    // its centre lies inside the triangle grown by the uv tolerance, i.e. within t = 3 * uv_delta * extent
    // (+ rounding) pixels of the quad's coordinate box, so the centres worth testing are columns
    //   min over vertices (floor + up) ... max over vertices (floor - dn)
    // -- about extent^2 instead of (extent + 1)^2 candidates.  t is taken for quads of up to K1_TIGHT_EXTENT
    // pixels; larger quads (and tolerances too large for t < 0.45) keep the plain floor box.
    const double tight_t = 3.0 * g.uv_delta * (K1_TIGHT_EXTENT + 2.0) + 1e-5;
    const bool tight = tight_t < 0.45;
    const double frac_up = tight ? 0.5 + tight_t : 2.0, frac_dn = tight ? 0.5 - tight_t : -2.0;
    auto to_px = [&](double vx, double vy, double &fx, double &fy, int &pi, int &pj) {
        fx = (vx - g.x_min) * inv_xr;
        fy = g.j_up ? (vy - g.y_min) * inv_yr : (g.y_max - vy) * inv_yr;
        const double cx = fmin(fmax(fx, -8.0), clamp_hi), cy = fmin(fmax(fy, -8.0), clamp_hi);
        const int ix = __double2int_rd(cx), iy = __double2int_rd(cy);
        const double rx = cx - static_cast<double>(ix), ry = cy - static_cast<double>(iy);
        pi = (ix << 2) | (rx > frac_up ? 2 : 0) | (rx < frac_dn ? 1 : 0);
        pj = (iy << 2) | (ry > frac_up ? 2 : 0) | (ry < frac_dn ? 1 : 0);
        return fabs(fx) < 1e9 && fabs(fy) < 1e9;  // false for NaN / inf
    };
    auto px_floor = [](int packed) { return packed >> 2; };
    auto px_first = [](int packed) { return (packed >> 2) + ((packed >> 1) & 1); };  // first column whose centre can count
    auto px_last = [](int packed) { return (packed >> 2) - (packed & 1); };          // last such column
    double fx0, fy0;
    int pi0, pj0;
    bool ok0 = to_px(col_ok ? __ldg(g.x + j_begin * g.src_pitch + col) : NAN,
                     col_ok ? __ldg(g.y + j_begin * g.src_pitch + col) : NAN, fx0, fy0, pi0, pj0);
    double fx1 = __shfl_down_sync(0xffffffffu, fx0, 1), fy1 = __shfl_down_sync(0xffffffffu, fy0, 1);
    int pi1 = __shfl_down_sync(0xffffffffu, pi0, 1), pj1 = __shfl_down_sync(0xffffffffu, pj0, 1);
    bool ok1 = (__ballot_sync(0xffffffffu, ok0) >> ((lane + 1) & 31)) & 1u;

    // the next vertex row is fetched one iteration ahead so that its latency hides behind the
    // pixel scans of the current row
    double xn = col_ok ? __ldg(g.x + (j_begin + 1) * g.src_pitch + col) : NAN;
    double yn = col_ok ? __ldg(g.y + (j_begin + 1) * g.src_pitch + col) : NAN;
    for (int64_t j = j_begin; j < j_end; ++j) {
        double fx2, fy2;
        int pi2, pj2;
        const bool ok2 = to_px(xn, yn, fx2, fy2, pi2, pj2);
        if (j + 1 < j_end) {
            xn = col_ok ? __ldg(g.x + (j + 2) * g.src_pitch + col) : NAN;
            yn = col_ok ? __ldg(g.y + (j + 2) * g.src_pitch + col) : NAN;
        }
        const double fx3 = __shfl_down_sync(0xffffffffu, fx2, 1), fy3 = __shfl_down_sync(0xffffffffu, fy2, 1);
        const int pi3 = __shfl_down_sync(0xffffffffu, pi2, 1), pj3 = __shfl_down_sync(0xffffffffu, pj2, 1);
        const bool ok3 = (__ballot_sync(0xffffffffu, ok2) >> ((lane + 1) & 31)) & 1u;
        bool slow = false;
        if (quad_lane) {
            // (the packed values order like their floors, so min / max commute with the unpacking)
            const int ni_lo = min(min(pi0, pi1), min(pi2, pi3)), ni_hi = max(max(pi0, pi1), max(pi2, pi3));
            const int nj_lo = min(min(pj0, pj1), min(pj2, pj3)), nj_hi = max(max(pj0, pj1), max(pj2, pj3));
            const int bi_lo = px_floor(ni_lo), bi_hi = px_floor(ni_hi), bj_lo = px_floor(nj_lo), bj_hi = px_floor(nj_hi);
            if (!(ok0 && ok1 && ok2 && ok3) || !small_tolerance) {
                slow = true;  // non-finite vertices: generic path (it also rejects far-away quads)
            } else if (!(bi_hi < 0 || bi_lo >= W || bj_hi < R0 || bj_lo >= R1)) {
                if (bi_hi - bi_lo >= static_cast<int>(K1_MAX_EXTENT) || bj_hi - bj_lo >= static_cast<int>(K1_MAX_EXTENT)) {
                    slow = true;
                } else {
                    int i_lo = max(bi_lo, 0), i_hi = min(bi_hi, W - 1), j_lo = max(bj_lo, R0), j_hi = min(bj_hi, R1 - 1);
                    if (bi_hi - bi_lo < static_cast<int>(K1_TIGHT_EXTENT) && bj_hi - bj_lo < static_cast<int>(K1_TIGHT_EXTENT)) {
                        i_lo = max(i_lo, min(min(px_first(pi0), px_first(pi1)), min(px_first(pi2), px_first(pi3))));
                        i_hi = min(i_hi, max(max(px_last(pi0), px_last(pi1)), max(px_last(pi2), px_last(pi3))));
                        j_lo = max(j_lo, min(min(px_first(pj0), px_first(pj1)), min(px_first(pj2), px_first(pj3))));
                        j_hi = min(j_hi, max(max(px_last(pj0), px_last(pj1)), max(px_last(pj2), px_last(pj3))));
                    }
                    if (i_lo > i_hi || j_lo > j_hi) {
                        // no pixel centre inside the grown quad: nothing to claim -- unless a collapsed triangle
                        // lets rounding noise accept something, which the generic kernel finds out
                        slow = px_degenerate(fx0, fy0, fx1, fy1, fx2, fy2) || px_degenerate(fx3, fy3, fx2, fy2, fx1, fy1);
                    } else {
                        const uint32_t qkey = static_cast<uint32_t>(j * nqi + col);
                        const int qj = static_cast<int>(j);
                        const double pcx = static_cast<double>(i_lo) + 0.5, pcy = static_cast<double>(j_lo) + 0.5;
                        // triangle A: origin p0, u towards p1, v towards p2; triangle B: origin p3, u towards p2, v towards p1
                        const PxEdges fa = make_px_edges(fx0, fy0, fx1, fy1, fx2, fy2, pcx, pcy, uv_lo, uv_hi, coord_bound);
                        const PxEdges fb = make_px_edges(fx3, fy3, fx2, fy2, fx1, fy1, pcx, pcy, uv_lo, uv_hi, coord_bound);
                        if (fa.degenerate || fb.degenerate) slow = true;
                        const double dx3_a = -(fa.dx1 + fa.dx2), dx3_b = -(fb.dx1 + fb.dx2);
                        const uint32_t rej_hi_a = static_cast<uint32_t>(__double2hiint(-2.0 * fa.two_m));
                        const uint32_t rej_hi_b = static_cast<uint32_t>(__double2hiint(-2.0 * fb.two_m));
                        // The pixel box of an ordinary quad lies inside ONE reference tile (a box of a few pixels
                        // against tiles of hundreds); the rare quad that straddles a tile border takes the generic
                        // kernel, which visits every tile its box touches.
                        const int tx = static_cast<int>(div_magic(static_cast<uint32_t>(i_lo), g.magic_tw));
                        const int ty = static_cast<int>(div_magic(static_cast<uint32_t>(j_lo), g.magic_th));
                        if (i_hi >= (tx + 1) * g.tile_w || j_hi >= (ty + 1) * g.tile_h) slow = true;
                        if (!slow) {
                            if (tw.id != ty * g.ntx + tx) load_tile_win(g, ty, tx, tw);
                            if (tw.has && qi >= tw.qi_lo && qi <= tw.qi_hi && qj >= tw.qj_lo && qj <= tw.qj_hi) {
                                uint32_t *claim_row = g.claims + (static_cast<int64_t>(j_lo) - g.row_begin) * g.dst_w;
                                for (int gj = j_lo; gj <= j_hi; ++gj, claim_row += g.dst_w) {
                                    const double rj = static_cast<double>(gj - j_lo);
                                    double a1 = fma(rj, fa.dy1, fa.e1), a2 = fma(rj, fa.dy2, fa.e2);
                                    double b1 = fma(rj, fb.dy1, fb.e1), b2 = fma(rj, fb.dy2, fb.e2);
                                    double a3 = fa.k3 - (a1 + a2), b3 = fb.k3 - (b1 + b2);  // third condition, stepped too
                                    for (int gi = i_lo; gi <= i_hi; ++gi) {
                                        // straight-line decisions (no nested branches): both triangles are
                                        // evaluated, the atomic is the only predicated operation
                                        const uint32_t ha1 = __double2hiint(a1), ha2 = __double2hiint(a2), ha3 = __double2hiint(a3);
                                        const uint32_t hb1 = __double2hiint(b1), hb2 = __double2hiint(b2), hb3 = __double2hiint(b3);
                                        const bool acc_a = static_cast<int>(ha1 | ha2 | ha3) >= 0;
                                        const bool rej_a = max(max(ha1, ha2), ha3) > rej_hi_a;
                                        const bool acc_b = !acc_a && rej_a && static_cast<int>(hb1 | hb2 | hb3) >= 0;
                                        const bool rej_b = max(max(hb1, hb2), hb3) > rej_hi_b;
                                        // within the margin of an edge: the whole quad is redone by the generic
                                        // kernel with the reference's arithmetic (atomicMin is idempotent)
                                        slow = slow || (!acc_a && !acc_b && !(rej_a && rej_b));
                                        if (acc_a || acc_b) atomicMin(claim_row + gi, (qkey << 1) | (acc_b ? 1u : 0u));
                                        a1 += fa.dx1; a2 += fa.dx2; a3 += dx3_a;
                                        b1 += fb.dx1; b2 += fb.dx2; b3 += dx3_b;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        // quads for the generic path are queued (warp-aggregated append) and handled by k1_scatter_slow,
        // one quad per thread, instead of stalling the 30 other lanes of this warp
        const unsigned slow_mask = __ballot_sync(0xffffffffu, slow);
        if (slow_mask) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(g.slow_count, __popc(slow_mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (slow) g.slow_list[base + __popc(slow_mask & ((1u << lane) - 1u))] = static_cast<uint32_t>(j * nqi + col);
        }
        fx0 = fx2; fy0 = fy2; fx1 = fx3; fy1 = fy3;
        pi0 = pi2; pj0 = pj2; pi1 = pi3; pj1 = pj3;
        ok0 = ok2; ok1 = ok3;
    }
}

// The queued quads, one per thread (grid-stride over the queue).
__global__ void __launch_bounds__(256) k1_scatter_slow(IjGeom g) {
    ScatterConst k;
    k.x_scale = g.x_res; k.y_scale = g.j_up ? g.y_res : -g.y_res;
    k.inv_xs = 1.0 / k.x_scale; k.inv_ys = 1.0 / k.y_scale;
    k.uv_lo = -g.uv_delta; k.uv_hi = dadd(1.0, dmul(2.0, g.uv_delta));
    const unsigned n = *g.slow_count;
    const int64_t nqi = g.src_w - 1;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint32_t qkey = g.slow_list[e];
        const int64_t j0 = qkey / nqi, i0 = qkey - j0 * nqi;
        const int64_t s0 = j0 * g.src_pitch + i0, s2 = s0 + g.src_pitch;
        scatter_quad_generic(g, k, __ldg(g.x + s0), __ldg(g.y + s0), __ldg(g.x + s0 + 1), __ldg(g.y + s0 + 1),
                             __ldg(g.x + s2), __ldg(g.y + s2), __ldg(g.x + s2 + 1), __ldg(g.y + s2 + 1),
                             static_cast<int>(i0), static_cast<int>(j0), qkey);
    }
}

// ---------------------------------------------------------------------------
// k1_resolve
// ---------------------------------------------------------------------------
// One block = K1R_PX * K1R_THREADS consecutive pixels of one target row; a thread resolves K1R_PX of
// them, K1R_THREADS apart (coalesced), sharing the row context; the claim words are loaded first so
// that their latency and that of the gathered vertex loads overlap.
constexpr int K1R_PX = 4;
#ifndef XRS_K1R_MINBLOCKS
#define XRS_K1R_MINBLOCKS 5  // 48 registers: 0.317 -> 0.292 ms on C2 (variants 1 / 5 / 6: tools/r2f_pass.sh, profiles/r02c_k1_resolve_variants.txt)
#endif
__global__ void __launch_bounds__(K1R_THREADS, XRS_K1R_MINBLOCKS) k1_resolve(const __grid_constant__ IjGeom g) {
    const int64_t c_first = static_cast<int64_t>(blockIdx.x) * (K1R_THREADS * K1R_PX) + threadIdx.x;
    const int64_t r = g.row_begin + blockIdx.y;
    const int64_t n_rows = g.row_end - g.row_begin;
    const int64_t row_o = static_cast<int64_t>(blockIdx.y) * g.dst_w;
    const ResolveRow row = resolve_row(g, r);
    uint32_t claim[K1R_PX];
#pragma unroll
    for (int k = 0; k < K1R_PX; ++k) {
        const int64_t c = c_first + k * K1R_THREADS;
        claim[k] = c < g.dst_w ? __ldcs(g.claims + row_o + c) : K1_NOCLAIM;
    }
#pragma unroll
    for (int k = 0; k < K1R_PX; ++k) {
        const int64_t c = c_first + k * K1R_THREADS;
        if (c >= g.dst_w) break;
        double oi, oj;
        resolve_pixel(g, row, c, claim[k], oi, oj);
        st_stream(g.ij + row_o + c, oi);
        st_stream(g.ij + n_rows * g.dst_w + row_o + c, oj);
    }
}

static int64_t claims_bytes_of(int64_t rows, int64_t dst_w) {
    return (rows * dst_w * static_cast<int64_t>(sizeof(uint32_t)) + 15) / 16 * 16;
}

int k1_make_geom(const char *who, const double *x, const double *y, int64_t src_h, int64_t src_w, int64_t src_pitch,
                 const int64_t *tile_boxes, int64_t dst_h, int64_t dst_w, int32_t tile_h, int32_t tile_w, double x_min,
                 double y_min, double y_max, double x_res, double y_res, int32_t is_j_axis_up, double uv_delta,
                 int64_t row_begin, int64_t row_end, void *workspace, IjGeom *out) {
    const std::string w(who);
    if (!x || !y || !tile_boxes || !workspace) return fail(w + ": null pointer");
    if (row_begin < 0 || row_end > dst_h || row_begin >= row_end) return fail(w + ": bad row range");
    if (src_h < 2 || src_w < 2 || src_pitch < src_w) return fail(w + ": source must be at least 2x2");
    if (dst_h < 1 || dst_w < 1 || tile_h < 1 || tile_w < 1) return fail(w + ": bad target shape");
    if (dst_h > (1 << 30) || dst_w > (1 << 30) || src_w > (1 << 30) || src_h > (1 << 30)) return fail(w + ": image too large");
    if (!(x_res > 0.0) || !(y_res > 0.0)) return fail(w + ": resolution must be positive");
    if ((src_h - 1) * (src_w - 1) >= 0x7fffffffLL) return fail(w + ": source has too many quads (2^31 or more)");
    if (reinterpret_cast<uintptr_t>(workspace) & 15) return fail(w + ": workspace must be 16-byte aligned");
    if (ceil_div(src_h - 1, K1S_ROWS) > 65535) return fail(w + ": source too tall");
    IjGeom g;
    g.x = x; g.y = y; g.src_h = src_h; g.src_w = src_w; g.src_pitch = src_pitch;
    g.tile_boxes = tile_boxes; g.ij = nullptr; g.claims = static_cast<uint32_t *>(workspace);
    g.dst_h = dst_h; g.dst_w = dst_w;
    g.tile_h = static_cast<int>(std::min<int64_t>(tile_h, dst_h));
    g.tile_w = static_cast<int>(std::min<int64_t>(tile_w, dst_w));
    g.ntx = static_cast<int>(ceil_div(dst_w, g.tile_w));
    g.nty = static_cast<int>(ceil_div(dst_h, g.tile_h));
    g.x_min = x_min; g.y_min = y_min; g.y_max = y_max; g.x_res = x_res; g.y_res = y_res;
    g.j_up = is_j_axis_up ? 1 : 0; g.uv_delta = uv_delta;
    g.row_begin = row_begin; g.row_end = row_end;
    g.fp_cols = nullptr;
    g.magic_nqi = div_magic_of(static_cast<uint64_t>(src_w - 1));
    g.magic_tw = div_magic_of(static_cast<uint64_t>(g.tile_w));
    g.magic_th = div_magic_of(static_cast<uint64_t>(g.tile_h));
    g.slow_list = reinterpret_cast<uint32_t *>(static_cast<char *>(workspace) + claims_bytes_of(row_end - row_begin, dst_w));
    g.slow_count = reinterpret_cast<unsigned int *>(g.slow_list + (src_h - 1) * (src_w - 1));
    *out = g;
    return 0;
}

int k1_enqueue_claims(const IjGeom &g, cudaStream_t st) {
    const int64_t n_rows = g.row_end - g.row_begin;
    const int64_t n_vec = ceil_div(n_rows * g.dst_w, 4);
    XRS_TIMED("k1_init_claims", st, k1_init_claims<<<static_cast<unsigned>(ceil_div(n_vec, 256)), 256, 0, st>>>(reinterpret_cast<uint4 *>(g.claims), n_vec, g.slow_count, g));
    XRS_LAUNCH_CHECK("k1_init_claims");
    const dim3 sgrid(static_cast<unsigned>(ceil_div(ceil_div(g.src_w - 1, 31), K1S_WARPS)),
                     static_cast<unsigned>(ceil_div(g.src_h - 1, K1S_ROWS)));
    XRS_TIMED("k1_scatter", st, k1_scatter<<<sgrid, K1S_WARPS * 32, 0, st>>>(g));
    XRS_LAUNCH_CHECK("k1_scatter");
    int dev = 0, sms = 0;
    XRS_CUDA(cudaGetDevice(&dev));
    XRS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    XRS_TIMED("k1_scatter_slow", st, k1_scatter_slow<<<static_cast<unsigned>(sms * 4), 256, 0, st>>>(g));
    XRS_LAUNCH_CHECK("k1_scatter_slow");
    return 0;
}

}  // namespace xrs

using namespace xrs;

extern "C" {

static int64_t claims_bytes(int64_t rows, int64_t dst_w) {
    return (rows * dst_w * static_cast<int64_t>(sizeof(uint32_t)) + 15) / 16 * 16;
}

// layout: [claims: 4 B per target pixel of the row band][slow-quad queue: 4 B per source quad]
// [counter, quad-row range: 16 B][quad-column range per block of K1S_ROWS quad rows: 8 B each]
int64_t xrs_rectify_ij_workspace_bytes(int64_t src_h, int64_t src_w, int64_t dst_rows, int64_t dst_w) {
    if (src_h < 2 || src_w < 2 || dst_rows < 1 || dst_w < 1) return 0;
    return claims_bytes(dst_rows, dst_w) + (src_h - 1) * (src_w - 1) * static_cast<int64_t>(sizeof(uint32_t)) + 16 +
           8 * ceil_div(src_h - 1, K1S_ROWS);
}

int xrs_rectify_ij(const double *x, const double *y, int64_t src_h, int64_t src_w, int64_t src_pitch,
                   const int64_t *tile_boxes, double *ij, int64_t dst_h, int64_t dst_w, int32_t tile_h,
                   int32_t tile_w, double x_min, double y_min, double y_max, double x_res, double y_res,
                   int32_t is_j_axis_up, double uv_delta, int64_t row_begin, int64_t row_end,
                   const int32_t *src_col_ranges, void *workspace, void *stream) {
    if (!ij) return fail("xrs_rectify_ij: null pointer");
    IjGeom g;
    if (int rc = k1_make_geom("xrs_rectify_ij", x, y, src_h, src_w, src_pitch, tile_boxes, dst_h, dst_w, tile_h, tile_w,
                              x_min, y_min, y_max, x_res, y_res, is_j_axis_up, uv_delta, row_begin, row_end, workspace,
                              &g))
        return rc;
    g.ij = ij;
    g.fp_cols = src_col_ranges;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int rc = k1_enqueue_claims(g, st)) return rc;
    const int64_t n_rows = row_end - row_begin;
    if (n_rows > 65535) return fail("xrs_rectify_ij: more than 65535 target rows per call");
    const dim3 rgrid(static_cast<unsigned>(ceil_div(dst_w, K1R_THREADS * K1R_PX)), static_cast<unsigned>(n_rows));
    XRS_TIMED("k1_resolve", st, k1_resolve<<<rgrid, K1R_THREADS, 0, st>>>(g));
    XRS_LAUNCH_CHECK("k1_resolve");
    return 0;
}

}  // extern "C"

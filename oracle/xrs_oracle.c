/*
 * oracle/xrs_oracle.c -- CPU restatement of the xcube-resampling hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the timed CPU baseline.  The product path (xcube_resampling_b200) never
 * imports it.
 *
 * Every function restates, in plain scalar C, the arithmetic of a numba /
 * numpy kernel of the reference (paths relative to /root/reference):
 *
 *   xrso_ij_bboxes        gridmapping/bboxes.py:28-106   compute_ij_bboxes
 *   xrso_rectify_ij_block rectify.py:424-576             _compute_target_source_ij_sequential/_line
 *   xrso_rectify_ij       rectify.py:373-419             _compute_target_source_ij_block, one call per tile
 *   xrso_gather_ij        rectify.py:605-734             _compute_var_image_block/_sequential/_for_dest_line
 *   xrso_reproject_block  reproject.py:268-335           _reproject_block
 *   xrso_mode             coarsen.py:114-155             mode / _mode_from_normalized
 *
 * Parity pinning: tests/test_oracle_golden.py checks these functions against
 * tests/golden/*.npz, which were produced by tests/golden/make_golden.py
 * running the reference's own numba kernels (imported from /root/reference
 * with stubbed third-party modules) and against the expected arrays of the
 * reference's unit tests.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (no FMA
 * contraction: numba/LLVM does not contract either).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define XRSO_EXPORT __attribute__((visibility("default")))

enum { XRSO_NEAREST = 0, XRSO_BILINEAR = 1, XRSO_TRIANGULAR = 2 };

/* dtype codes shared with include/xrs.h */
enum {
    XRSO_F32 = 0, XRSO_F64 = 1, XRSO_U8 = 2, XRSO_I8 = 3, XRSO_U16 = 4,
    XRSO_I16 = 5, XRSO_I32 = 6, XRSO_U32 = 7, XRSO_I64 = 8
};

XRSO_EXPORT int xrso_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

XRSO_EXPORT void xrso_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* np.floor(v).astype(np.int64) on x86: NaN, +inf, -inf and out-of-range
 * all give INT64_MIN (cvttsd2si "integer indefinite"), SURVEY.md 7.4. */
static inline int64_t floor_to_i64(double v) {
    double f = floor(v);
    if (!(f >= -9223372036854775808.0 && f < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)f;
}

/* ------------------------------------------------------------------ */
/* gridmapping/bboxes.py:28-106                                        */
/* ------------------------------------------------------------------ */
XRSO_EXPORT void xrso_ij_bboxes(const double *x_image, const double *y_image, int64_t h, int64_t w,
                                const double *xy_boxes, int64_t n, double xy_border, int64_t ij_border,
                                int64_t *ij_boxes /* (n,4), pre-set to -1 */) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t k = 0; k < n; ++k) {
        int64_t *box = ij_boxes + 4 * k;
        const double x_lo = xy_boxes[4 * k + 0] - xy_border;
        const double y_lo = xy_boxes[4 * k + 1] - xy_border;
        const double x_hi = xy_boxes[4 * k + 2] + xy_border;
        const double y_hi = xy_boxes[4 * k + 3] + xy_border;
        for (int64_t j = 0; j < h; ++j) {
            const double *xr = x_image + j * w;
            const double *yr = y_image + j * w;
            for (int64_t i = 0; i < w; ++i) {
                const double x = xr[i];
                if (!(x_lo <= x && x <= x_hi)) continue;
                const double y = yr[i];
                if (!(y_lo <= y && y <= y_hi)) continue;
                if (box[0] < 0) {
                    box[0] = i; box[1] = j; box[2] = i + 1; box[3] = j + 1;
                } else {
                    if (i < box[0]) box[0] = i;
                    if (j < box[1]) box[1] = j;
                    if (i + 1 > box[2]) box[2] = i + 1;
                    if (j + 1 > box[3]) box[3] = j + 1;
                }
            }
        }
        if (ij_border != 0 && box[0] != -1) {
            int64_t i0 = box[0] - ij_border, j0 = box[1] - ij_border;
            int64_t i1 = box[2] + ij_border, j1 = box[3] + ij_border;
            if (i0 < 0) i0 = 0;
            if (j0 < 0) j0 = 0;
            if (i1 > w) i1 = w;
            if (j1 > h) j1 = h;
            box[0] = i0; box[1] = j0; box[2] = i1; box[3] = j1;
        }
    }
}

/* ------------------------------------------------------------------ */
/* rectify.py:737-768 helpers                                          */
/* ------------------------------------------------------------------ */
static inline double tri_det(double ax, double ay, double bx, double by, double cx, double cy) {
    return (ax - bx) * (ay - cy) - (ax - cx) * (ay - by);
}
static inline double tri_u(double px, double py, double ax, double ay, double cx, double cy) {
    return (ax - px) * (ay - cy) - (ay - py) * (ax - cx);
}
static inline double tri_v(double px, double py, double ax, double ay, double bx, double by) {
    return (ay - py) * (ax - bx) - (ax - px) * (ay - by);
}
static inline double clamp01(double t) { return t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t); }

/* rectify.py:424-576 -- one reference tile, strictly sequential scatter.
 * x/y: source window (win_h, win_w) with row stride src_stride (elements).
 * ij: two planes (i then j) of the tile, row stride dst_stride, plane
 * stride dst_plane; the tile is (dst_h, dst_w) and is NaN-filled first. */
XRSO_EXPORT void xrso_rectify_ij_block(const double *x, const double *y, int64_t win_h, int64_t win_w,
                                       int64_t src_stride, int64_t src_i_min, int64_t src_j_min,
                                       double *ij, int64_t dst_h, int64_t dst_w, int64_t dst_stride,
                                       int64_t dst_plane, double x_off, double y_off, double x_scale,
                                       double y_scale, double uv_delta) {
    double *out_i = ij, *out_j = ij + dst_plane;
    for (int64_t r = 0; r < dst_h; ++r)
        for (int64_t c = 0; c < dst_w; ++c) out_i[r * dst_stride + c] = out_j[r * dst_stride + c] = NAN;

    const double lo = -uv_delta, hi = 1.0 + 2 * uv_delta;
    for (int64_t j0 = 0; j0 + 1 < win_h; ++j0) {
        for (int64_t i0 = 0; i0 + 1 < win_w; ++i0) {
            const double qx[4] = {x[j0 * src_stride + i0], x[j0 * src_stride + i0 + 1],
                                  x[(j0 + 1) * src_stride + i0], x[(j0 + 1) * src_stride + i0 + 1]};
            const double qy[4] = {y[j0 * src_stride + i0], y[j0 * src_stride + i0 + 1],
                                  y[(j0 + 1) * src_stride + i0], y[(j0 + 1) * src_stride + i0 + 1]};
            int64_t ci_lo = INT64_MAX, ci_hi = INT64_MIN, cj_lo = INT64_MAX, cj_hi = INT64_MIN;
            for (int k = 0; k < 4; ++k) {
                int64_t pi = floor_to_i64((qx[k] - x_off) / x_scale);
                int64_t pj = floor_to_i64((qy[k] - y_off) / y_scale);
                if (pi < ci_lo) ci_lo = pi;
                if (pi > ci_hi) ci_hi = pi;
                if (pj < cj_lo) cj_lo = pj;
                if (pj > cj_hi) cj_hi = pj;
            }
            if (ci_hi < 0 || cj_hi < 0 || ci_lo >= dst_w || cj_lo >= dst_h) continue;
            if (ci_lo < 0) ci_lo = 0;
            if (ci_hi >= dst_w) ci_hi = dst_w - 1;
            if (cj_lo < 0) cj_lo = 0;
            if (cj_hi >= dst_h) cj_hi = dst_h - 1;

            double det_a = tri_det(qx[0], qy[0], qx[1], qy[1], qx[2], qy[2]);
            if (isnan(det_a)) det_a = 0.0;
            double det_b = tri_det(qx[3], qy[3], qx[2], qy[2], qx[1], qy[1]);
            if (isnan(det_b)) det_b = 0.0;
            if (det_a == 0.0 && det_b == 0.0) continue;

            for (int64_t dj = cj_lo; dj <= cj_hi; ++dj) {
                const double py = y_off + ((double)dj + 0.5) * y_scale;
                for (int64_t di = ci_lo; di <= ci_hi; ++di) {
                    if (!isnan(out_i[dj * dst_stride + di])) continue; /* first writer wins */
                    const double px = x_off + ((double)di + 0.5) * x_scale;
                    double si = -1.0, sj = -1.0;
                    if (det_a != 0.0) {
                        const double u = tri_u(px, py, qx[0], qy[0], qx[2], qy[2]) / det_a;
                        const double v = tri_v(px, py, qx[0], qy[0], qx[1], qy[1]) / det_a;
                        if (u >= lo && v >= lo && u + v <= hi) {
                            si = (double)i0 + clamp01(u);
                            sj = (double)j0 + clamp01(v);
                        }
                    }
                    if (si == -1.0 && det_b != 0.0) {
                        const double u = tri_u(px, py, qx[3], qy[3], qx[1], qy[1]) / det_b;
                        const double v = tri_v(px, py, qx[3], qy[3], qx[2], qy[2]) / det_b;
                        if (u >= lo && v >= lo && u + v <= hi) {
                            si = (double)(i0 + 1) - clamp01(u);
                            sj = (double)(j0 + 1) - clamp01(v);
                        }
                    }
                    if (si != -1.0) {
                        out_i[dj * dst_stride + di] = (double)src_i_min + si;
                        out_j[dj * dst_stride + di] = (double)src_j_min + sj;
                    }
                }
            }
        }
    }
}

/* rectify.py:312-419 -- whole (2,H,W) ij image, one sequential block per
 * reference tile; tiles in row-major block_id order (dask.py:97-120), run on
 * an OpenMP team the way dask's threaded scheduler runs them. */
XRSO_EXPORT void xrso_rectify_ij(const double *x, const double *y, int64_t src_h, int64_t src_w,
                                 const int64_t *src_ij_bboxes /* (n_tiles,4) */, double *ij /* (2,H,W) */,
                                 int64_t dst_h, int64_t dst_w, int64_t tile_h, int64_t tile_w, double x_min,
                                 double y_min, double y_max, double x_res, double y_res, int is_j_axis_up,
                                 double uv_delta) {
    const int64_t nty = (dst_h + tile_h - 1) / tile_h, ntx = (dst_w + tile_w - 1) / tile_w;
    const int64_t plane = dst_h * dst_w;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < nty * ntx; ++t) {
        const int64_t ty = t / ntx, tx = t % ntx;
        const int64_t r0 = ty * tile_h, c0 = tx * tile_w;
        const int64_t th = (r0 + tile_h <= dst_h) ? tile_h : dst_h - r0;
        const int64_t tw = (c0 + tile_w <= dst_w) ? tile_w : dst_w - c0;
        double *blk = ij + r0 * dst_w + c0;
        const int64_t *bb = src_ij_bboxes + 4 * t;
        if (bb[0] == -1) {
            for (int64_t r = 0; r < th; ++r)
                for (int64_t c = 0; c < tw; ++c) blk[r * dst_w + c] = blk[plane + r * dst_w + c] = NAN;
            continue;
        }
        /* slice [j_min : j_max+1, i_min : i_max+1] clipped to the image (rectify.py:397-399) */
        int64_t j_end = bb[3] + 1 < src_h ? bb[3] + 1 : src_h;
        int64_t i_end = bb[2] + 1 < src_w ? bb[2] + 1 : src_w;
        const double x_off = x_min + (double)c0 * x_res;
        const double y_off = is_j_axis_up ? y_min + (double)r0 * y_res : y_max - (double)r0 * y_res;
        xrso_rectify_ij_block(x + bb[1] * src_w + bb[0], y + bb[1] * src_w + bb[0], j_end - bb[1], i_end - bb[0],
                              src_w, bb[0], bb[1], blk, th, tw, dst_w, plane, x_off, y_off, x_res,
                              is_j_axis_up ? y_res : -y_res, uv_delta);
    }
}

/* ------------------------------------------------------------------ */
/* rectify.py:605-734 gather                                           */
/* ------------------------------------------------------------------ */
static inline double load_as_f64(const void *p, int dtype, int64_t idx) {
    switch (dtype) {
    case XRSO_F32: return (double)((const float *)p)[idx];
    case XRSO_F64: return ((const double *)p)[idx];
    case XRSO_U8: return (double)((const uint8_t *)p)[idx];
    case XRSO_I8: return (double)((const int8_t *)p)[idx];
    case XRSO_U16: return (double)((const uint16_t *)p)[idx];
    case XRSO_I16: return (double)((const int16_t *)p)[idx];
    case XRSO_I32: return (double)((const int32_t *)p)[idx];
    case XRSO_U32: return (double)((const uint32_t *)p)[idx];
    default: return (double)((const int64_t *)p)[idx];
    }
}
/* numba's float64 -> T store (rectify.py:734) is a C cast */
static inline void store_from_f64(void *p, int dtype, int64_t idx, double v) {
    switch (dtype) {
    case XRSO_F32: ((float *)p)[idx] = (float)v; break;
    case XRSO_F64: ((double *)p)[idx] = v; break;
    case XRSO_U8: ((uint8_t *)p)[idx] = (uint8_t)(int64_t)v; break;
    case XRSO_I8: ((int8_t *)p)[idx] = (int8_t)(int64_t)v; break;
    case XRSO_U16: ((uint16_t *)p)[idx] = (uint16_t)(int64_t)v; break;
    case XRSO_I16: ((int16_t *)p)[idx] = (int16_t)(int64_t)v; break;
    case XRSO_I32: ((int32_t *)p)[idx] = (int32_t)(int64_t)v; break;
    case XRSO_U32: ((uint32_t *)p)[idx] = (uint32_t)(int64_t)v; break;
    default: ((int64_t *)p)[idx] = (int64_t)v; break;
    }
}
static inline int64_t iclamp(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Whole-image form of rectify.py:605-734.  The per-tile source bbox the
 * reference derives from nanmin/nanmax(ij) only shifts indices exactly and
 * clamps at the true source edge (DESIGN.md "K2 tile independence"), so the
 * result equals a per-pixel gather against the full source.
 * src: (bands, src_h, src_w), dst: (bands, dst_h, dst_w) pre-filled by the
 * caller with the fill value; pixels whose ij is NaN are left untouched. */
XRSO_EXPORT int xrso_gather_ij(const void *src, int dtype, int64_t bands, int64_t src_h, int64_t src_w,
                               const double *ij, void *dst, int64_t dst_h, int64_t dst_w, int method) {
    if (method != XRSO_NEAREST && method != XRSO_BILINEAR && method != XRSO_TRIANGULAR) return 1;
    const int64_t splane = src_h * src_w, dplane = dst_h * dst_w;
    const int64_t esz = (dtype == XRSO_F64 || dtype == XRSO_I64) ? 8
                        : (dtype == XRSO_F32 || dtype == XRSO_I32 || dtype == XRSO_U32) ? 4
                        : (dtype == XRSO_U16 || dtype == XRSO_I16) ? 2 : 1;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < dst_h; ++r) {
        for (int64_t c = 0; c < dst_w; ++c) {
            const double fi = ij[r * dst_w + c], fj = ij[dplane + r * dst_w + c];
            if (isnan(fi) || isnan(fj)) continue;
            int64_t i0 = (int64_t)fi, j0 = (int64_t)fj;
            const double u = fi - (double)i0, v = fj - (double)j0;
            for (int64_t b = 0; b < bands; ++b) {
                const char *sp = (const char *)src + b * splane * esz;
                char *dp = (char *)dst + b * dplane * esz;
                double val;
                if (method == XRSO_NEAREST) {
                    int64_t ii = i0, jj = j0;
                    if (u > 0.5) ii = iclamp(i0 + 1, 0, src_w - 1);
                    if (v > 0.5) jj = iclamp(j0 + 1, 0, src_h - 1);
                    val = load_as_f64(sp, dtype, jj * src_w + ii);
                } else {
                    const int64_t i1 = iclamp(i0 + 1, 0, src_w - 1), j1 = iclamp(j0 + 1, 0, src_h - 1);
                    const double v01 = load_as_f64(sp, dtype, j0 * src_w + i1);
                    const double v10 = load_as_f64(sp, dtype, j1 * src_w + i0);
                    if (method == XRSO_BILINEAR) {
                        const double v00 = load_as_f64(sp, dtype, j0 * src_w + i0);
                        const double v11 = load_as_f64(sp, dtype, j1 * src_w + i1);
                        const double a = v00 + u * (v01 - v00);
                        const double bb = v10 + u * (v11 - v10);
                        val = a + v * (bb - a);
                    } else if (u + v < 1.0) {
                        const double v00 = load_as_f64(sp, dtype, j0 * src_w + i0);
                        val = v00 + u * (v01 - v00) + v * (v10 - v00);
                    } else {
                        const double v11 = load_as_f64(sp, dtype, j1 * src_w + i1);
                        val = v11 + (1.0 - u) * (v10 - v11) + (1.0 - v) * (v01 - v11);
                    }
                }
                store_from_f64(dp, dtype, r * dst_w + c, val);
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* coarsen.py:114-155 mode                                             */
/* ------------------------------------------------------------------ */
/* windows: (n_win, win_len) int64 values already normalised by the block
 * minimum (coarsen.py:132); out[i] = offset + argmax(count), lowest value
 * winning ties (coarsen.py:150-153). */
XRSO_EXPORT void xrso_mode(const int64_t *windows, int64_t n_win, int64_t win_len, int64_t offset,
                           int64_t mode_range, int64_t *out) {
#pragma omp parallel
    {
        int64_t *counts = (int64_t *)malloc((size_t)mode_range * sizeof(int64_t));
#pragma omp for schedule(static)
        for (int64_t w = 0; w < n_win; ++w) {
            memset(counts, 0, (size_t)mode_range * sizeof(int64_t));
            for (int64_t k = 0; k < win_len; ++k) counts[windows[w * win_len + k]] += 1;
            int64_t best = 0, best_n = counts[0];
            for (int64_t m = 1; m < mode_range; ++m)
                if (counts[m] > best_n) { best_n = counts[m]; best = m; }
            out[w] = best + offset;
        }
        free(counts);
    }
}

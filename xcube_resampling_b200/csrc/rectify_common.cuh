// rectify_common.cuh -- pieces of the rectification kernels shared between rectify_ij.cu (K1) and
// gather.cu (K2 and the fused K1-resolve + K2 path): geometry, the reference's determinant /
// barycentric expressions and the per-pixel resolve step.
#pragma once

#include "common.cuh"

namespace xrs {

constexpr uint32_t K1_NOCLAIM = 0xffffffffu;
constexpr int K1_SENTINEL = INT32_MIN;  // stands for np.int64 min (non-finite vertex)
constexpr int K1S_ROWS = 32;            // quad rows marched by one warp
constexpr int K1S_WARPS = 8;
constexpr int K1R_THREADS = 256;

struct IjGeom {
    const double *x, *y;
    int64_t src_h, src_w, src_pitch;
    const int64_t *tile_boxes;
    double *ij;
    uint32_t *claims;  // (row_end - row_begin, dst_w): smallest accepting quad index per pixel
    int64_t dst_h, dst_w;
    int tile_h, tile_w, ntx, nty;
    double x_min, y_min, y_max, x_res, y_res;
    int j_up;
    double uv_delta;
    int64_t row_begin, row_end;  // target rows computed by this call
    uint32_t *slow_list;         // quads that need the generic (multi-tile) treatment
    unsigned int *slow_count;
    const int32_t *fp_cols;      // optional: per group of K1S_ROWS quad rows (c_min, -c_max) of the quads the
                                 // caller made resident (xrs_band_quad_footprints); nullptr = whole swath
    uint64_t magic_nqi, magic_tw, magic_th;  // ceil(2^64 / d) for d = src_w - 1, tile_w, tile_h (0 when d == 1)
};

// n / d for 32-bit n through the precomputed M = ceil(2^64 / d): exact (the excess n * (M - 2^64/d) / 2^64
// stays below 2^-32 < 1/d), four multiply-adds instead of the ~20 instructions of a generic division.
__host__ __device__ static inline uint64_t div_magic_of(uint64_t d) {
    return d <= 1 ? 0ull : (~0ull / d) + 1ull;  // floor((2^64 - 1) / d) + 1 == ceil(2^64 / d) for every d >= 2
}
__device__ __forceinline__ uint32_t div_magic(uint32_t n, uint64_t magic) {
    return magic ? static_cast<uint32_t>(__umul64hi(static_cast<uint64_t>(n), magic)) : n;
}

__device__ __forceinline__ double tri_det(double ax, double ay, double bx, double by, double cx, double cy) {
    return dsub(dmul(dsub(ax, bx), dsub(ay, cy)), dmul(dsub(ax, cx), dsub(ay, by)));
}
__device__ __forceinline__ double tri_u(double px, double py, double ax, double ay, double cx, double cy) {
    return dsub(dmul(dsub(ax, px), dsub(ay, cy)), dmul(dsub(ay, py), dsub(ax, cx)));
}
__device__ __forceinline__ double tri_v(double px, double py, double ax, double ay, double bx, double by) {
    return dsub(dmul(dsub(ay, py), dsub(ax, bx)), dmul(dsub(ax, px), dsub(ay, by)));
}
__device__ __forceinline__ double clamp01(double t) { return t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t); }


// What resolve_pixel needs from the target row alone (tile row, its origin, the pixel-centre
// y coordinate): computed once per row and reused for every pixel of it.
struct ResolveRow {
    int ty, r0;
    double py;
};
__device__ __forceinline__ ResolveRow resolve_row(const IjGeom &g, int64_t r) {
    ResolveRow w;
    w.ty = static_cast<int>(div_magic(static_cast<uint32_t>(r), g.magic_th));
    w.r0 = w.ty * g.tile_h;
    const double y_off = g.j_up ? dadd(g.y_min, dmul(static_cast<double>(w.r0), g.y_res))
                                : dsub(g.y_max, dmul(static_cast<double>(w.r0), g.y_res));
    const double y_scale = g.j_up ? g.y_res : -g.y_res;
    w.py = dadd(y_off, dmul(dadd(static_cast<double>(static_cast<int>(r) - w.r0), 0.5), y_scale));
    return w;
}

// rectify.py:545-576 for target pixel (r, c) (global row r) whose claim word is `claim`: the winning
// triangle recomputes u, v with the reference's exact expressions (tile-local offsets, _rn
// arithmetic, divided form) and yields the fractional source index (oi, oj); NaN if unclaimed.
__device__ __forceinline__ void resolve_pixel(const IjGeom &g, const ResolveRow &row, int64_t c, uint32_t claim,
                                              double &oi, double &oj) {
    oi = oj = NAN;
    if (claim == K1_NOCLAIM) return;
    const uint32_t qkey = claim >> 1, nqi = static_cast<uint32_t>(g.src_w - 1);
    const bool tri_b = claim & 1u;  // which triangle the scatter accepted (A is tried first, rectify.py:556-573)
    const uint32_t j0 = div_magic(qkey, g.magic_nqi), i0 = qkey - j0 * nqi;
    const int tx = static_cast<int>(div_magic(static_cast<uint32_t>(c), g.magic_tw));
    const int c0 = tx * g.tile_w;
    const int64_t *bb = g.tile_boxes + 4 * (static_cast<int64_t>(row.ty) * g.ntx + tx);
    const int bb0 = static_cast<int>(__ldg(bb)), bb1 = static_cast<int>(__ldg(bb + 1));
    const double x_off = dadd(g.x_min, dmul(static_cast<double>(c0), g.x_res));
    const double px = dadd(x_off, dmul(dadd(static_cast<double>(static_cast<int>(c) - c0), 0.5), g.x_res));
    const double py = row.py;
    const int64_t s0 = static_cast<int64_t>(j0) * g.src_pitch + i0, s2 = s0 + g.src_pitch;
    // origin vertex o, u-direction vertex pu, v-direction vertex pv of the accepted triangle:
    // A = (p0; p1, p2), B = (p3; p2, p1)
    const int64_t so = tri_b ? s2 + 1 : s0, su = tri_b ? s2 : s0 + 1, sv = tri_b ? s0 + 1 : s2;
    const double ox = __ldg(g.x + so), oy = __ldg(g.y + so);
    const double ux = __ldg(g.x + su), uy = __ldg(g.y + su);
    const double vx = __ldg(g.x + sv), vy = __ldg(g.y + sv);
    const double det = tri_det(ox, oy, ux, uy, vx, vy);
    const double u = ddiv(tri_u(px, py, ox, oy, vx, vy), det);
    const double v = ddiv(tri_v(px, py, ox, oy, ux, uy), det);
    const double fi = clamp01(u), fj = clamp01(v);
    // rectify.py:564-576: window-local index + fraction, then + window origin
    const int wi = static_cast<int>(i0) - bb0, wj = static_cast<int>(j0) - bb1;
    const double li = tri_b ? dsub(static_cast<double>(wi + 1), fi) : dadd(static_cast<double>(wi), fi);
    const double lj = tri_b ? dsub(static_cast<double>(wj + 1), fj) : dadd(static_cast<double>(wj), fj);
    oi = dadd(static_cast<double>(bb0), li);
    oj = dadd(static_cast<double>(bb1), lj);
}
__device__ __forceinline__ void resolve_pixel(const IjGeom &g, int64_t r, int64_t c, uint32_t claim, double &oi,
                                              double &oj) {
    resolve_pixel(g, resolve_row(g, r), c, claim, oi, oj);
}

// rectify_ij.cu: argument checks + geometry of one xrs_rectify_ij-style call, and the claim stage
// (k1_init_claims, k1_scatter, k1_scatter_slow) enqueued on `st`.
int k1_make_geom(const char *who, const double *x, const double *y, int64_t src_h, int64_t src_w, int64_t src_pitch,
                 const int64_t *tile_boxes, int64_t dst_h, int64_t dst_w, int32_t tile_h, int32_t tile_w, double x_min,
                 double y_min, double y_max, double x_res, double y_res, int32_t is_j_axis_up, double uv_delta,
                 int64_t row_begin, int64_t row_end, void *workspace, IjGeom *g);
int k1_enqueue_claims(const IjGeom &g, cudaStream_t st);

}  // namespace xrs

// rectify.cu -- sm_100a kernels of the rectification path.
//
//   K0  xrs_tile_src_bboxes   gridmapping/bboxes.py:28-106   (one O(S) pass)
//   K1  xrs_rectify_ij        rectify.py:312-576, 737-768
//   K2  xrs_gather_ij         rectify.py:579-734
//
// Compiled with -fmad=false; the parity-critical expressions additionally use
// explicit round-to-nearest intrinsics so they round like numba's LLVM code.
#include "common.cuh"

namespace xrs {

// ===========================================================================
// K0: per-tile source windows
// ===========================================================================
constexpr int K0_THREADS = 256;
constexpr int K0_COLS_PER_THREAD = 4;
constexpr int K0_SEG = K0_THREADS * K0_COLS_PER_THREAD;
constexpr int K0_SMEM_TILES = 2048;  // tile table kept in shared memory up to this many tiles

// number of entries of ascending arr[0..n) that are <= v
__device__ __forceinline__ int count_le(const double *arr, int n, double v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (arr[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// number of entries of ascending arr[0..n) that are < v
__device__ __forceinline__ int count_lt(const double *arr, int n, double v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (arr[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void k0_init_table(int4 *table, int n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) table[t] = make_int4(INT32_MAX, INT32_MAX, -1, -1);
}

// x axis arrays ascend with tx; the y arrays may descend with ty (j-axis-down grids).
template <bool SMEM_TABLE>
__global__ void __launch_bounds__(K0_THREADS)
k0_tile_windows(const double *__restrict__ x, const double *__restrict__ y, int64_t h, int64_t w, int64_t pitch,
                const double *__restrict__ g_x_lo, const double *__restrict__ g_x_hi, int ntx,
                const double *__restrict__ g_y_lo, const double *__restrict__ g_y_hi, int nty,
                int4 *__restrict__ table) {
    extern __shared__ __align__(16) unsigned char k0_smem[];
    double *s_x_lo = reinterpret_cast<double *>(k0_smem);
    double *s_x_hi = s_x_lo + ntx;
    double *s_y_lo = s_x_hi + ntx;
    double *s_y_hi = s_y_lo + nty;
    int4 *s_tab = reinterpret_cast<int4 *>(s_y_hi + nty);
    const int n_tiles = ntx * nty;
    // the y axis arrives in tile-row order; j-axis-down grids have it descending
    const bool y_desc = nty > 1 && g_y_lo[0] > g_y_lo[nty - 1];

    for (int k = threadIdx.x; k < ntx; k += blockDim.x) { s_x_lo[k] = g_x_lo[k]; s_x_hi[k] = g_x_hi[k]; }
    for (int k = threadIdx.x; k < nty; k += blockDim.x) {
        const int src = y_desc ? nty - 1 - k : k;
        s_y_lo[k] = g_y_lo[src];
        s_y_hi[k] = g_y_hi[src];
    }
    if (SMEM_TABLE)
        for (int k = threadIdx.x; k < n_tiles; k += blockDim.x) s_tab[k] = make_int4(INT32_MAX, INT32_MAX, -1, -1);
    __syncthreads();
    int4 *tab = SMEM_TABLE ? s_tab : table;

    const int64_t n_chunks = ceil_div(w, K0_SEG);
    const int64_t n_seg = h * n_chunks;
    for (int64_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
        const int64_t j = seg / n_chunks;
        const int64_t col0 = (seg - j * n_chunks) * K0_SEG;
        double vx[K0_COLS_PER_THREAD], vy[K0_COLS_PER_THREAD];
#pragma unroll
        for (int k = 0; k < K0_COLS_PER_THREAD; ++k) {
            const int64_t i = col0 + k * K0_THREADS + threadIdx.x;
            const bool in = i < w;
            vx[k] = in ? ld_stream(x + j * pitch + i) : NAN;
            vy[k] = in ? ld_stream(y + j * pitch + i) : NAN;
        }
#pragma unroll
        for (int k = 0; k < K0_COLS_PER_THREAD; ++k) {
            const int i = static_cast<int>(col0 + k * K0_THREADS + threadIdx.x);
            // {t : lo[t] <= v <= hi[t]} is the contiguous range [a, b]
            const int xb = count_le(s_x_lo, ntx, vx[k]) - 1, xa = count_lt(s_x_hi, ntx, vx[k]);
            const int yb = count_le(s_y_lo, nty, vy[k]) - 1, ya = count_lt(s_y_hi, nty, vy[k]);
            const bool valid = (xa <= xb) && (ya <= yb);  // NaN -> counts 0 -> b = -1 -> invalid
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            if (!valid) continue;
            const unsigned long long key = (static_cast<unsigned long long>(xa) << 48) |
                                           (static_cast<unsigned long long>(xb) << 32) |
                                           (static_cast<unsigned long long>(ya) << 16) |
                                           static_cast<unsigned long long>(yb);
            int uniform = 0;
            __match_all_sync(m, key, &uniform);
            int i_lo = i, i_hi = i;
            bool writer = true;
            if (uniform) {
                i_lo = __reduce_min_sync(m, i);
                i_hi = __reduce_max_sync(m, i);
                writer = (threadIdx.x & 31) == (__ffs(m) - 1);
            }
            if (writer) {
                const int jj = static_cast<int>(j);
                for (int ky = ya; ky <= yb; ++ky) {
                    const int ty = y_desc ? nty - 1 - ky : ky;
                    for (int tx = xa; tx <= xb; ++tx) {
                        int *e = reinterpret_cast<int *>(tab + ty * ntx + tx);
                        atomicMin(e + 0, i_lo);
                        atomicMin(e + 1, jj);
                        atomicMax(e + 2, i_hi + 1);
                        atomicMax(e + 3, jj + 1);
                    }
                }
            }
        }
    }
    if (SMEM_TABLE) {
        __syncthreads();
        for (int k = threadIdx.x; k < n_tiles; k += blockDim.x) {
            const int4 e = s_tab[k];
            if (e.z >= 0) {
                int *g = reinterpret_cast<int *>(table + k);
                atomicMin(g + 0, e.x);
                atomicMin(g + 1, e.y);
                atomicMax(g + 2, e.z);
                atomicMax(g + 3, e.w);
            }
        }
    }
}

__global__ void k0_finalize(const int4 *__restrict__ table, int n, int ij_border, int64_t w, int64_t h,
                            int64_t *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int4 e = table[t];
    int64_t b0 = -1, b1 = -1, b2 = -1, b3 = -1;
    if (e.z >= 0) {
        b0 = e.x; b1 = e.y; b2 = e.z; b3 = e.w;
        if (ij_border != 0) {
            b0 -= ij_border; b1 -= ij_border; b2 += ij_border; b3 += ij_border;
            if (b0 < 0) b0 = 0;
            if (b1 < 0) b1 = 0;
            if (b2 > w) b2 = w;
            if (b3 > h) b3 = h;
        }
    }
    out[4 * t + 0] = b0; out[4 * t + 1] = b1; out[4 * t + 2] = b2; out[4 * t + 3] = b3;
}

// ===========================================================================
// K1: source-index image
// ===========================================================================
constexpr int K1_CW = 64;        // CTA target tile width  [px]
constexpr int K1_CH = 32;        // CTA target tile height [px]
constexpr int K1_THREADS = 256;
constexpr int K1_VCAP = 3584;    // vertices of the source window staged per chunk
constexpr int K1_MAX_COLS = 254; // quad columns per chunk (row stride <= 256)
constexpr uint32_t K1_NOCLAIM = 0xffffffffu;
constexpr int K1_SENTINEL = INT32_MIN;  // stands for np.int64 min (non-finite vertex)

struct RectGeom {
    const double *x, *y;
    int64_t src_h, src_w, src_pitch;
    const int64_t *tile_boxes;
    double *ij;
    int64_t dst_h, dst_w;
    int tile_h, tile_w, ntx, nty;
    int ncx_per_tile, ncy_per_tile, ncx_total, ncy_total;
    double x_min, y_min, y_max, x_res, y_res;
    int j_up;
    double uv_delta;
    int4 *cta_win;  // per CTA tile: quad-index bounds (i_lo, j_lo, i_hi, j_hi), inclusive
    int64_t row_begin, row_end;  // target rows computed by this call; ij holds exactly these rows
    int cy_begin, cy_end;        // CTA tile rows intersecting [row_begin, row_end)
};

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

// np.floor(v).astype(np.int64) reduced to int32: non-finite / out-of-range -> sentinel
// (x86 gives INT64_MIN for those), everything else clamped to +-2^30, which preserves
// every comparison against pixel ranges.
__device__ __forceinline__ int floor_px(double v) {
    const double f = floor(v);
    if (!(f >= -9223372036854775808.0 && f < 9223372036854775808.0)) return K1_SENTINEL;
    return static_cast<int>(fmin(fmax(f, -1073741824.0), 1073741824.0));
}

__device__ __forceinline__ int cta_index_of_px(int g, int tile, int per_tile, int csize) {
    const int t = g / tile;
    return t * per_tile + (g - t * tile) / csize;
}

__global__ void k1_init_windows(int4 *win, int n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) win[t] = make_int4(INT32_MAX, INT32_MAX, -1, -1);
}

// K1a: every source quad registers itself with the CTA tiles its (conservative)
// target-pixel bounding box touches.
__global__ void __launch_bounds__(256) k1_bin_quads(RectGeom g) {
    const int64_t nqi = g.src_w - 1, nqj = g.src_h - 1;
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * 32 + (threadIdx.x & 31);
    const int64_t j0 = static_cast<int64_t>(blockIdx.y) * 8 + (threadIdx.x >> 5);
    bool valid = i0 < nqi && j0 < nqj;
    const double inv_xr = 1.0 / g.x_res, inv_yr = 1.0 / g.y_res;
    double fx[4], fy[4];
    bool fin[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t off = (j0 + (k >> 1)) * g.src_pitch + i0 + (k & 1);
        const double px = valid ? __ldg(g.x + off) : NAN, py = valid ? __ldg(g.y + off) : NAN;
        fx[k] = (px - g.x_min) * inv_xr;
        fy[k] = g.j_up ? (py - g.y_min) * inv_yr : (g.y_max - py) * inv_yr;
        fin[k] = isfinite(fx[k]) && isfinite(fy[k]);
    }
    double lo_x = INFINITY, hi_x = -INFINITY, lo_y = INFINITY, hi_y = -INFINITY;
    if (fin[0] && fin[1] && fin[2] && fin[3]) {
        // all-finite quad: the reference scans [min floor(p), max floor(p)] (rectify.py:500-526);
        // +-1 px absorbs the difference between global and tile-local pixel arithmetic
        lo_x = fmin(fmin(fx[0], fx[1]), fmin(fx[2], fx[3])) - 1.0;
        hi_x = fmax(fmax(fx[0], fx[1]), fmax(fx[2], fx[3])) + 1.0;
        lo_y = fmin(fmin(fy[0], fy[1]), fmin(fy[2], fy[3])) - 1.0;
        hi_y = fmax(fmax(fy[0], fy[1]), fmax(fy[2], fy[3])) + 1.0;
    } else {
        // A triangle with a non-finite vertex can never accept a pixel (its u or v is NaN),
        // so only triangles with three finite vertices contribute; their acceptance region is
        // the triangle grown by the uv tolerance, bounded here by a few % of its extent + 1 px.
        if (fin[0] && fin[1] && fin[2]) {
            lo_x = fmin(fx[0], fmin(fx[1], fx[2])); hi_x = fmax(fx[0], fmax(fx[1], fx[2]));
            lo_y = fmin(fy[0], fmin(fy[1], fy[2])); hi_y = fmax(fy[0], fmax(fy[1], fy[2]));
        }
        if (fin[3] && fin[2] && fin[1]) {
            lo_x = fmin(lo_x, fmin(fx[3], fmin(fx[1], fx[2]))); hi_x = fmax(hi_x, fmax(fx[3], fmax(fx[1], fx[2])));
            lo_y = fmin(lo_y, fmin(fy[3], fmin(fy[1], fy[2]))); hi_y = fmax(hi_y, fmax(fy[3], fmax(fy[1], fy[2])));
        }
        const double mx = (0.01 + 4.0 * g.uv_delta) * (hi_x - lo_x) + 1.0;
        const double my = (0.01 + 4.0 * g.uv_delta) * (hi_y - lo_y) + 1.0;
        lo_x -= mx; hi_x += mx; lo_y -= my; hi_y += my;
    }
    const double W = static_cast<double>(g.dst_w), H = static_cast<double>(g.dst_h);
    valid = valid && (lo_x <= hi_x) && !(hi_x < 0.0 || hi_y < 0.0 || lo_x >= W || lo_y >= H);
    int cxa = 0, cxb = 0, cya = 0, cyb = 0;
    if (valid) {
        const int gx0 = static_cast<int>(floor(fmax(lo_x, 0.0))), gx1 = static_cast<int>(floor(fmin(hi_x, W - 1.0)));
        const int gy0 = static_cast<int>(floor(fmax(lo_y, 0.0))), gy1 = static_cast<int>(floor(fmin(hi_y, H - 1.0)));
        cxa = cta_index_of_px(gx0, g.tile_w, g.ncx_per_tile, K1_CW);
        cxb = cta_index_of_px(gx1, g.tile_w, g.ncx_per_tile, K1_CW);
        cya = max(cta_index_of_px(gy0, g.tile_h, g.ncy_per_tile, K1_CH), g.cy_begin);
        cyb = min(cta_index_of_px(gy1, g.tile_h, g.ncy_per_tile, K1_CH), g.cy_end - 1);
        valid = cya <= cyb;
    }
    // warp aggregation: lanes with the same CTA range share one set of atomics
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const unsigned long long key = (static_cast<unsigned long long>(cxa) << 48) |
                                   (static_cast<unsigned long long>(cxb) << 32) |
                                   (static_cast<unsigned long long>(cya) << 16) |
                                   static_cast<unsigned long long>(cyb);
    const unsigned peers = __match_any_sync(m, key);
    const int qi = static_cast<int>(i0), qj = static_cast<int>(j0);
    const int qi_lo = __reduce_min_sync(peers, qi), qi_hi = __reduce_max_sync(peers, qi);
    if ((threadIdx.x & 31) != (__ffs(peers) - 1)) return;
    for (int cy = cya; cy <= cyb; ++cy)
        for (int cx = cxa; cx <= cxb; ++cx) {
            int *e = reinterpret_cast<int *>(g.cta_win + static_cast<int64_t>(cy) * g.ncx_total + cx);
            atomicMin(e + 0, qi_lo);
            atomicMin(e + 1, qj);
            atomicMax(e + 2, qi_hi);
            atomicMax(e + 3, qj);
        }
}

struct K1Smem {
    uint32_t claims[K1_CW * K1_CH];
    alignas(16) double vx[K1_VCAP];
    alignas(16) double vy[K1_VCAP];
    int vpi[K1_VCAP];
    int vpj[K1_VCAP];
    alignas(8) uint64_t bar;
};

// Triangle acceptance exactly as rectify.py:558-573; returns true and the
// fractional quad-local source position (relative to quad corner (0,0)).
__device__ __forceinline__ bool quad_accepts(double px, double py, const double qx[4], const double qy[4],
                                             double det_a, double det_b, double lo, double hi, double &fi,
                                             double &fj, bool &tri_b) {
    if (det_a != 0.0) {
        const double u = ddiv(tri_u(px, py, qx[0], qy[0], qx[2], qy[2]), det_a);
        const double v = ddiv(tri_v(px, py, qx[0], qy[0], qx[1], qy[1]), det_a);
        if (u >= lo && v >= lo && dadd(u, v) <= hi) {
            fi = clamp01(u); fj = clamp01(v); tri_b = false;
            return true;
        }
    }
    if (det_b != 0.0) {
        const double u = ddiv(tri_u(px, py, qx[3], qy[3], qx[1], qy[1]), det_b);
        const double v = ddiv(tri_v(px, py, qx[3], qy[3], qx[2], qy[2]), det_b);
        if (u >= lo && v >= lo && dadd(u, v) <= hi) {
            fi = clamp01(u); fj = clamp01(v); tri_b = true;
            return true;
        }
    }
    return false;
}

// K1b: one CTA per (sub-)tile of a reference tile.
__global__ void __launch_bounds__(K1_THREADS) k1_rectify_ij(RectGeom g) {
    extern __shared__ __align__(16) unsigned char k1_smem_raw[];
    K1Smem &s = *reinterpret_cast<K1Smem *>(k1_smem_raw);
    const int tid = threadIdx.x;

    const int cy = g.cy_begin + blockIdx.x / g.ncx_total, cx = blockIdx.x % g.ncx_total;
    const int cta = cy * g.ncx_total + cx;
    const int ty = cy / g.ncy_per_tile, sy = cy - ty * g.ncy_per_tile;
    const int tx = cx / g.ncx_per_tile, sx = cx - tx * g.ncx_per_tile;
    const int64_t r0 = static_cast<int64_t>(ty) * g.tile_h, c0 = static_cast<int64_t>(tx) * g.tile_w;
    const int th = static_cast<int>(min(static_cast<int64_t>(g.tile_h), g.dst_h - r0));
    const int tw = static_cast<int>(min(static_cast<int64_t>(g.tile_w), g.dst_w - c0));
    const int lx0 = sx * K1_CW;  // CTA origin, tile-local px
    const int lw = min(K1_CW, tw - lx0);
    // rows of this CTA tile, clipped to the tile and to the requested row range
    const int ly0 = static_cast<int>(max(static_cast<int64_t>(sy) * K1_CH, g.row_begin - r0));
    const int ly1 = static_cast<int>(min(static_cast<int64_t>(min((sy + 1) * K1_CH, th)), g.row_end - r0));
    const int lh = ly1 - ly0;
    if (lw <= 0 || lh <= 0) return;  // sub-tile beyond a clipped edge tile or outside the row range

    const int64_t plane = (g.row_end - g.row_begin) * g.dst_w;
    double *out_i = g.ij + (r0 + ly0 - g.row_begin) * g.dst_w + c0 + lx0;
    double *out_j = out_i + plane;

    // reference tile window (rectify.py:393-399) intersected with this CTA's quad window
    const int64_t *bb = g.tile_boxes + 4 * (static_cast<int64_t>(ty) * g.ntx + tx);
    const int64_t bb0 = bb[0], bb1 = bb[1], bb2 = bb[2], bb3 = bb[3];
    const int4 win = g.cta_win[cta];
    int64_t qi0 = 0, qi1 = -1, qj0 = 0, qj1 = -1;
    if (bb0 != -1 && win.z >= 0) {
        const int64_t i_end = min(bb2 + 1, g.src_w), j_end = min(bb3 + 1, g.src_h);  // vertex slice ends
        qi0 = max(static_cast<int64_t>(win.x), bb0); qi1 = min(static_cast<int64_t>(win.z), i_end - 2);
        qj0 = max(static_cast<int64_t>(win.y), bb1); qj1 = min(static_cast<int64_t>(win.w), j_end - 2);
    }
    if (qi1 < qi0 || qj1 < qj0) {  // CTA-uniform early out: nothing can land here
        for (int p = tid; p < lw * lh; p += K1_THREADS) {
            const int py = p / lw, px = p - py * lw;
            st_stream(out_i + static_cast<int64_t>(py) * g.dst_w + px, static_cast<double>(NAN));
            st_stream(out_j + static_cast<int64_t>(py) * g.dst_w + px, static_cast<double>(NAN));
        }
        return;
    }
    const int nqi = static_cast<int>(qi1 - qi0 + 1), nqj = static_cast<int>(qj1 - qj0 + 1);

    // tile-local geometry, same expressions as rectify.py:402-416
    const double x_off = dadd(g.x_min, dmul(static_cast<double>(c0), g.x_res));
    const double y_off = g.j_up ? dadd(g.y_min, dmul(static_cast<double>(r0), g.y_res))
                                : dsub(g.y_max, dmul(static_cast<double>(r0), g.y_res));
    const double x_scale = g.x_res, y_scale = g.j_up ? g.y_res : -g.y_res;
    const double uv_lo = -g.uv_delta, uv_hi = dadd(1.0, dmul(2.0, g.uv_delta));

    for (int p = tid; p < K1_CW * K1_CH; p += K1_THREADS) s.claims[p] = K1_NOCLAIM;
    if (tid == 0) {
        mbar_init(&s.bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const bool aligned = ((reinterpret_cast<uintptr_t>(g.x) | reinterpret_cast<uintptr_t>(g.y)) & 15) == 0;
    const int64_t n_valid = (g.src_h - 1) * g.src_pitch + g.src_w;  // elements known to be readable
    const int cq_i = min(nqi, K1_MAX_COLS);                        // quad columns per chunk
    const int rs = (cq_i + 1 + 2) & ~1;                            // even row stride >= vertex cols + 1
    const int cq_j = min(nqj, K1_VCAP / rs - 1);                   // quad rows per chunk
    uint32_t parity = 0;

    for (int64_t cj = qj0; cj <= qj1; cj += cq_j) {
        const int nj = static_cast<int>(min(static_cast<int64_t>(cq_j), qj1 - cj + 1));  // quad rows
        for (int64_t ci = qi0; ci <= qi1; ci += cq_i) {
            const int ni = static_cast<int>(min(static_cast<int64_t>(cq_i), qi1 - ci + 1));  // quad cols
            const int nvj = nj + 1, nvi = ni + 1;
            // ---- stage vertex rows [cj, cj+nvj) x [ci, ci+nvi) ------------------------
            // Row r lands at s.vx[r*rs + shift_r + c] with shift_r making the global source
            // address 16-byte aligned for the bulk copy.
            if (tid < 32) {
                uint32_t bytes = 0;
                if (aligned)
                    for (int r = tid; r < nvj; r += 32) {
                        const int64_t e = (cj + r) * g.src_pitch + ci;
                        const int sh = static_cast<int>(e & 1);
                        const int cnt = (nvi + sh + 1) & ~1;
                        if (e - sh + cnt <= n_valid) bytes += 2u * 8u * cnt;
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
                if (tid == 0) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(&s.bar, bytes);
                }
                __syncwarp();
                if (aligned)
                    for (int r = tid; r < nvj; r += 32) {
                        const int64_t e = (cj + r) * g.src_pitch + ci;
                        const int sh = static_cast<int>(e & 1);
                        const int cnt = (nvi + sh + 1) & ~1;
                        if (e - sh + cnt <= n_valid) {
                            bulk_g2s(&s.vx[r * rs], g.x + e - sh, 8u * cnt, &s.bar);
                            bulk_g2s(&s.vy[r * rs], g.y + e - sh, 8u * cnt, &s.bar);
                        }
                    }
            }
            // rows the bulk engine cannot take (unaligned base or tail of the allocation)
            for (int r = 0; r < nvj; ++r) {
                const int64_t e = (cj + r) * g.src_pitch + ci;
                const int sh = static_cast<int>(e & 1);
                const int cnt = (nvi + sh + 1) & ~1;
                if (aligned && e - sh + cnt <= n_valid) continue;
                for (int c = tid; c < nvi; c += K1_THREADS) {
                    s.vx[r * rs + sh + c] = g.x[e + c];
                    s.vy[r * rs + sh + c] = g.y[e + c];
                }
            }
            mbar_wait(&s.bar, parity);
            parity ^= 1;
            __syncthreads();
            // ---- vertex -> tile-local pixel index (rectify.py:500-501) --------------
            for (int v = tid; v < nvj * nvi; v += K1_THREADS) {
                const int r = v / nvi, c = v - r * nvi;
                const int sh = static_cast<int>(((cj + r) * g.src_pitch + ci) & 1);
                const int k = r * rs + sh + c;
                s.vpi[k] = floor_px(ddiv(dsub(s.vx[k], x_off), x_scale));
                s.vpj[k] = floor_px(ddiv(dsub(s.vy[k], y_off), y_scale));
            }
            __syncthreads();
            // ---- quads claim pixels ---------------------------------------------------
            for (int q = tid; q < nj * ni; q += K1_THREADS) {
                const int r = q / ni, c = q - r * ni;
                const int sh0 = static_cast<int>(((cj + r) * g.src_pitch + ci) & 1);
                const int sh1 = static_cast<int>(((cj + r + 1) * g.src_pitch + ci) & 1);
                const int k0 = r * rs + sh0 + c, k2 = (r + 1) * rs + sh1 + c;
                const int pi0 = s.vpi[k0], pi1 = s.vpi[k0 + 1], pi2 = s.vpi[k2], pi3 = s.vpi[k2 + 1];
                const int pj0 = s.vpj[k0], pj1 = s.vpj[k0 + 1], pj2 = s.vpj[k2], pj3 = s.vpj[k2 + 1];
                int i_lo = min(min(pi0, pi1), min(pi2, pi3)), i_hi = max(max(pi0, pi1), max(pi2, pi3));
                int j_lo = min(min(pj0, pj1), min(pj2, pj3)), j_hi = max(max(pj0, pj1), max(pj2, pj3));
                // rectify.py:508-526 (tile clip) then restriction to this CTA's pixels
                if (i_hi < 0 || j_hi < 0 || i_lo >= tw || j_lo >= th) continue;
                i_lo = max(i_lo, lx0); i_hi = min(i_hi, lx0 + lw - 1);
                j_lo = max(j_lo, ly0); j_hi = min(j_hi, ly0 + lh - 1);
                if (i_lo > i_hi || j_lo > j_hi) continue;
                const double qx[4] = {s.vx[k0], s.vx[k0 + 1], s.vx[k2], s.vx[k2 + 1]};
                const double qy[4] = {s.vy[k0], s.vy[k0 + 1], s.vy[k2], s.vy[k2 + 1]};
                double det_a = tri_det(qx[0], qy[0], qx[1], qy[1], qx[2], qy[2]);
                if (det_a != det_a) det_a = 0.0;
                double det_b = tri_det(qx[3], qy[3], qx[2], qy[2], qx[1], qy[1]);
                if (det_b != det_b) det_b = 0.0;
                if (det_a == 0.0 && det_b == 0.0) continue;
                const uint32_t qkey = static_cast<uint32_t>((cj + r - qj0) * nqi + (ci + c - qi0));
                for (int dj = j_lo; dj <= j_hi; ++dj) {
                    const double py = dadd(y_off, dmul(dadd(static_cast<double>(dj), 0.5), y_scale));
                    for (int di = i_lo; di <= i_hi; ++di) {
                        const double px = dadd(x_off, dmul(dadd(static_cast<double>(di), 0.5), x_scale));
                        double fi, fj;
                        bool tb;
                        if (quad_accepts(px, py, qx, qy, det_a, det_b, uv_lo, uv_hi, fi, fj, tb))
                            atomicMin(&s.claims[(dj - ly0) * K1_CW + (di - lx0)], qkey);
                    }
                }
            }
            __syncthreads();  // chunk buffers are reused by the next bulk copy
        }
    }

    // ---- resolve: the winning quad recomputes its fractional index ------------------
    for (int p = tid; p < lw * lh; p += K1_THREADS) {
        const int py_l = p / lw, px_l = p - py_l * lw;
        const uint32_t qkey = s.claims[py_l * K1_CW + px_l];
        double oi = NAN, oj = NAN;
        if (qkey != K1_NOCLAIM) {
            const int64_t j0 = qj0 + qkey / nqi, i0 = qi0 + qkey % nqi;
            double qx[4], qy[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t off = (j0 + (k >> 1)) * g.src_pitch + i0 + (k & 1);
                qx[k] = __ldg(g.x + off);
                qy[k] = __ldg(g.y + off);
            }
            double det_a = tri_det(qx[0], qy[0], qx[1], qy[1], qx[2], qy[2]);
            if (det_a != det_a) det_a = 0.0;
            double det_b = tri_det(qx[3], qy[3], qx[2], qy[2], qx[1], qy[1]);
            if (det_b != det_b) det_b = 0.0;
            const double py = dadd(y_off, dmul(dadd(static_cast<double>(ly0 + py_l), 0.5), y_scale));
            const double px = dadd(x_off, dmul(dadd(static_cast<double>(lx0 + px_l), 0.5), x_scale));
            double fi = 0.0, fj = 0.0;
            bool tb = false;
            quad_accepts(px, py, qx, qy, det_a, det_b, uv_lo, uv_hi, fi, fj, tb);
            // rectify.py:564-576: window-local index + fraction, then + window origin
            const double li = tb ? dsub(static_cast<double>(i0 + 1 - bb0), fi) : dadd(static_cast<double>(i0 - bb0), fi);
            const double lj = tb ? dsub(static_cast<double>(j0 + 1 - bb1), fj) : dadd(static_cast<double>(j0 - bb1), fj);
            oi = dadd(static_cast<double>(bb0), li);
            oj = dadd(static_cast<double>(bb1), lj);
        }
        st_stream(out_i + static_cast<int64_t>(py_l) * g.dst_w + px_l, oi);
        st_stream(out_j + static_cast<int64_t>(py_l) * g.dst_w + px_l, oj);
    }
}

// ===========================================================================
// K2: gather of all bands through the ij image
// ===========================================================================
constexpr int K2_MAX_BANDS = 32;
constexpr int K2_BX = 32, K2_BY = 8;

template <typename T>
struct PlaneTable {
    const T *src[K2_MAX_BANDS];
    T *dst[K2_MAX_BANDS];
};

template <typename T>
__device__ __forceinline__ double ld_f64(const T *p) { return static_cast<double>(__ldg(p)); }

template <typename T, int METHOD>
__global__ void __launch_bounds__(K2_BX *K2_BY)
k2_gather(PlaneTable<T> planes, int n_bands, int64_t src_h, int64_t src_w, int64_t src_pitch, int64_t win_i0,
          int64_t win_j0, const double *__restrict__ ij, int64_t dst_h, int64_t dst_w, T fill) {
    const int64_t c = static_cast<int64_t>(blockIdx.x) * K2_BX + threadIdx.x;
    const int64_t r = static_cast<int64_t>(blockIdx.y) * K2_BY + threadIdx.y;
    if (c >= dst_w || r >= dst_h) return;
    const int64_t o = r * dst_w + c;
    const double fi = ld_stream(ij + o), fj = ld_stream(ij + dst_h * dst_w + o);
    if (fi != fi || fj != fj) {
        for (int b = 0; b < n_bands; ++b) st_stream(planes.dst[b] + o, fill);
        return;
    }
    // rectify.py:689-692: int() truncation of non-negative values
    int64_t i0 = static_cast<int64_t>(fi), j0 = static_cast<int64_t>(fj);
    const double u = dsub(fi, static_cast<double>(i0)), v = dsub(fj, static_cast<double>(j0));
    if (METHOD == XRS_NEAREST) {
        if (u > 0.5) i0 = min(max(i0 + 1, int64_t(0)), src_w - 1);
        if (v > 0.5) j0 = min(max(j0 + 1, int64_t(0)), src_h - 1);
        const int64_t so = (j0 - win_j0) * src_pitch + (i0 - win_i0);
#pragma unroll 4
        for (int b = 0; b < n_bands; ++b) st_stream(planes.dst[b] + o, __ldg(planes.src[b] + so));
        return;
    }
    const int64_t i1 = min(max(i0 + 1, int64_t(0)), src_w - 1), j1 = min(max(j0 + 1, int64_t(0)), src_h - 1);
    const int64_t o00 = (j0 - win_j0) * src_pitch + (i0 - win_i0), o01 = (j0 - win_j0) * src_pitch + (i1 - win_i0);
    const int64_t o10 = (j1 - win_j0) * src_pitch + (i0 - win_i0), o11 = (j1 - win_j0) * src_pitch + (i1 - win_i0);
    if (METHOD == XRS_BILINEAR) {
#pragma unroll 4
        for (int b = 0; b < n_bands; ++b) {
            const T *sp = planes.src[b];
            const double v00 = ld_f64(sp + o00), v01 = ld_f64(sp + o01);
            const double v10 = ld_f64(sp + o10), v11 = ld_f64(sp + o11);
            const double a = dadd(v00, dmul(u, dsub(v01, v00)));
            const double bb = dadd(v10, dmul(u, dsub(v11, v10)));
            st_stream(planes.dst[b] + o, cast_from_f64<T>(dadd(a, dmul(v, dsub(bb, a)))));
        }
    } else {  // triangular, rectify.py:699-717
        const bool near = dadd(u, v) < 1.0;
#pragma unroll 4
        for (int b = 0; b < n_bands; ++b) {
            const T *sp = planes.src[b];
            const double v01 = ld_f64(sp + o01), v10 = ld_f64(sp + o10);
            double val;
            if (near) {
                const double v00 = ld_f64(sp + o00);
                val = dadd(dadd(v00, dmul(u, dsub(v01, v00))), dmul(v, dsub(v10, v00)));
            } else {
                const double v11 = ld_f64(sp + o11);
                val = dadd(dadd(v11, dmul(dsub(1.0, u), dsub(v10, v11))), dmul(dsub(1.0, v), dsub(v01, v11)));
            }
            st_stream(planes.dst[b] + o, cast_from_f64<T>(val));
        }
    }
}

template <typename T>
static int launch_gather(const void *const *src_planes, void *const *dst_planes, int n_bands, int64_t src_h,
                         int64_t src_w, int64_t src_pitch, int64_t win_i0, int64_t win_j0, const double *ij,
                         int64_t dst_h, int64_t dst_w, int method, double fill, cudaStream_t st) {
    const dim3 block(K2_BX, K2_BY);
    const dim3 grid(static_cast<unsigned>(ceil_div(dst_w, K2_BX)), static_cast<unsigned>(ceil_div(dst_h, K2_BY)));
    T fill_t;
    if (std::is_floating_point<T>::value) fill_t = static_cast<T>(fill);
    else fill_t = static_cast<T>(static_cast<long long>(fill));
    for (int b0 = 0; b0 < n_bands; b0 += K2_MAX_BANDS) {
        const int nb = std::min(K2_MAX_BANDS, n_bands - b0);
        PlaneTable<T> pt;
        for (int b = 0; b < K2_MAX_BANDS; ++b) {
            pt.src[b] = b < nb ? static_cast<const T *>(src_planes[b0 + b]) : nullptr;
            pt.dst[b] = b < nb ? static_cast<T *>(dst_planes[b0 + b]) : nullptr;
        }
        switch (method) {
        case XRS_NEAREST:
            k2_gather<T, XRS_NEAREST><<<grid, block, 0, st>>>(pt, nb, src_h, src_w, src_pitch, win_i0, win_j0, ij, dst_h, dst_w, fill_t);
            break;
        case XRS_BILINEAR:
            k2_gather<T, XRS_BILINEAR><<<grid, block, 0, st>>>(pt, nb, src_h, src_w, src_pitch, win_i0, win_j0, ij, dst_h, dst_w, fill_t);
            break;
        default:
            k2_gather<T, XRS_TRIANGULAR><<<grid, block, 0, st>>>(pt, nb, src_h, src_w, src_pitch, win_i0, win_j0, ij, dst_h, dst_w, fill_t);
            break;
        }
        XRS_LAUNCH_CHECK("k2_gather");
    }
    return 0;
}

}  // namespace xrs

using namespace xrs;

extern "C" {

int64_t xrs_tile_src_bboxes_workspace_bytes(int32_t ntx, int32_t nty) {
    return static_cast<int64_t>(ntx) * nty * static_cast<int64_t>(sizeof(int4));
}

int xrs_tile_src_bboxes(const double *x, const double *y, int64_t src_h, int64_t src_w, int64_t src_pitch,
                        const double *x_lo, const double *x_hi, int32_t ntx, const double *y_lo,
                        const double *y_hi, int32_t nty, int32_t ij_border, int64_t *out_boxes, void *workspace,
                        void *stream) {
    if (!x || !y || !x_lo || !x_hi || !y_lo || !y_hi || !out_boxes || !workspace) return fail("xrs_tile_src_bboxes: null pointer");
    if (src_h < 1 || src_w < 1 || src_pitch < src_w) return fail("xrs_tile_src_bboxes: bad source shape");
    if (ntx < 1 || nty < 1 || ntx > 65535 || nty > 65535) return fail("xrs_tile_src_bboxes: tile counts must be in [1, 65535]");
    if (src_w > INT32_MAX - 1 || src_h > INT32_MAX - 1) return fail("xrs_tile_src_bboxes: source too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_tiles = ntx * nty;
    int4 *table = static_cast<int4 *>(workspace);
    k0_init_table<<<static_cast<unsigned>(ceil_div(n_tiles, 256)), 256, 0, st>>>(table, n_tiles);
    XRS_LAUNCH_CHECK("k0_init_table");

    int dev = 0, sms = 0;
    XRS_CUDA(cudaGetDevice(&dev));
    XRS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t n_seg = src_h * ceil_div(src_w, K0_SEG);
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(n_seg, static_cast<int64_t>(sms) * 8));
    const size_t axis_bytes = static_cast<size_t>(2 * ntx + 2 * nty) * sizeof(double);
    if (n_tiles <= K0_SMEM_TILES) {
        const size_t smem = axis_bytes + static_cast<size_t>(n_tiles) * sizeof(int4);
        XRS_CUDA(cudaFuncSetAttribute(k0_tile_windows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        k0_tile_windows<true><<<grid, K0_THREADS, smem, st>>>(x, y, src_h, src_w, src_pitch, x_lo, x_hi, ntx, y_lo, y_hi, nty, table);
    } else {
        if (axis_bytes > 200 * 1024) return fail("xrs_tile_src_bboxes: too many tile rows/columns");
        XRS_CUDA(cudaFuncSetAttribute(k0_tile_windows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(axis_bytes)));
        k0_tile_windows<false><<<grid, K0_THREADS, axis_bytes, st>>>(x, y, src_h, src_w, src_pitch, x_lo, x_hi, ntx, y_lo, y_hi, nty, table);
    }
    XRS_LAUNCH_CHECK("k0_tile_windows");
    k0_finalize<<<static_cast<unsigned>(ceil_div(n_tiles, 256)), 256, 0, st>>>(table, n_tiles, ij_border, src_w, src_h, out_boxes);
    XRS_LAUNCH_CHECK("k0_finalize");
    return 0;
}

static void k1_cta_grid(int64_t dst_h, int64_t dst_w, int32_t tile_h, int32_t tile_w, int &ntx, int &nty, int &ncx, int &ncy) {
    ntx = static_cast<int>(ceil_div(dst_w, tile_w));
    nty = static_cast<int>(ceil_div(dst_h, tile_h));
    ncx = static_cast<int>(ceil_div(tile_w, K1_CW));
    ncy = static_cast<int>(ceil_div(tile_h, K1_CH));
}

int64_t xrs_rectify_ij_workspace_bytes(int64_t dst_h, int64_t dst_w, int32_t tile_h, int32_t tile_w) {
    if (dst_h < 1 || dst_w < 1 || tile_h < 1 || tile_w < 1) return 0;
    int ntx, nty, ncx, ncy;
    k1_cta_grid(dst_h, dst_w, static_cast<int32_t>(std::min<int64_t>(tile_h, dst_h)),
                static_cast<int32_t>(std::min<int64_t>(tile_w, dst_w)), ntx, nty, ncx, ncy);
    return static_cast<int64_t>(ntx) * ncx * nty * ncy * static_cast<int64_t>(sizeof(int4));
}

int xrs_rectify_ij(const double *x, const double *y, int64_t src_h, int64_t src_w, int64_t src_pitch,
                   const int64_t *tile_boxes, double *ij, int64_t dst_h, int64_t dst_w, int32_t tile_h,
                   int32_t tile_w, double x_min, double y_min, double y_max, double x_res, double y_res,
                   int32_t is_j_axis_up, double uv_delta, int64_t row_begin, int64_t row_end, void *workspace,
                   void *stream) {
    if (!x || !y || !tile_boxes || !ij || !workspace) return fail("xrs_rectify_ij: null pointer");
    if (row_begin < 0 || row_end > dst_h || row_begin >= row_end) return fail("xrs_rectify_ij: bad row range");
    if (src_h < 2 || src_w < 2 || src_pitch < src_w) return fail("xrs_rectify_ij: source must be at least 2x2");
    if (dst_h < 1 || dst_w < 1 || tile_h < 1 || tile_w < 1) return fail("xrs_rectify_ij: bad target shape");
    if (dst_h > (1 << 30) || dst_w > (1 << 30) || src_w > (1 << 30) || src_h > (1 << 30)) return fail("xrs_rectify_ij: image too large");
    if (!(x_res > 0.0) || !(y_res > 0.0)) return fail("xrs_rectify_ij: resolution must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RectGeom g;
    g.x = x; g.y = y; g.src_h = src_h; g.src_w = src_w; g.src_pitch = src_pitch;
    g.tile_boxes = tile_boxes; g.ij = ij; g.dst_h = dst_h; g.dst_w = dst_w;
    g.tile_h = static_cast<int>(std::min<int64_t>(tile_h, dst_h));
    g.tile_w = static_cast<int>(std::min<int64_t>(tile_w, dst_w));
    k1_cta_grid(dst_h, dst_w, g.tile_h, g.tile_w, g.ntx, g.nty, g.ncx_per_tile, g.ncy_per_tile);
    g.ncx_total = g.ntx * g.ncx_per_tile;
    g.ncy_total = g.nty * g.ncy_per_tile;
    if (g.ncx_total > 65535 || g.ncy_total > 65535) return fail("xrs_rectify_ij: too many CTA tiles per axis (>65535)");
    g.x_min = x_min; g.y_min = y_min; g.y_max = y_max; g.x_res = x_res; g.y_res = y_res;
    g.j_up = is_j_axis_up ? 1 : 0; g.uv_delta = uv_delta;
    g.cta_win = static_cast<int4 *>(workspace);
    const int64_t n_cta = static_cast<int64_t>(g.ncx_total) * g.ncy_total;
    g.row_begin = row_begin; g.row_end = row_end;
    g.cy_begin = static_cast<int>(row_begin / g.tile_h) * g.ncy_per_tile + static_cast<int>(row_begin % g.tile_h) / K1_CH;
    g.cy_end = static_cast<int>((row_end - 1) / g.tile_h) * g.ncy_per_tile + static_cast<int>((row_end - 1) % g.tile_h) / K1_CH + 1;
    const int64_t n_launch = static_cast<int64_t>(g.ncx_total) * (g.cy_end - g.cy_begin);
    if ((src_h - 1) * (src_w - 1) >= 0xffffffffLL) return fail("xrs_rectify_ij: source has too many quads");

    k1_init_windows<<<static_cast<unsigned>(ceil_div(n_cta, 256)), 256, 0, st>>>(g.cta_win, static_cast<int>(n_cta));
    XRS_LAUNCH_CHECK("k1_init_windows");
    const dim3 bgrid(static_cast<unsigned>(ceil_div(src_w - 1, 32)), static_cast<unsigned>(ceil_div(src_h - 1, 8)));
    k1_bin_quads<<<bgrid, 256, 0, st>>>(g);
    XRS_LAUNCH_CHECK("k1_bin_quads");
    XRS_CUDA(cudaFuncSetAttribute(k1_rectify_ij, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(K1Smem))));
    k1_rectify_ij<<<static_cast<unsigned>(n_launch), K1_THREADS, sizeof(K1Smem), st>>>(g);
    XRS_LAUNCH_CHECK("k1_rectify_ij");
    return 0;
}

int xrs_gather_ij(const void *const *src_planes_host, void *const *dst_planes_host, int32_t n_bands,
                  int32_t dtype, int64_t src_h, int64_t src_w, int64_t src_pitch, int64_t win_i0, int64_t win_j0,
                  const double *ij, int64_t dst_h, int64_t dst_w, int32_t method, double fill, void *stream) {
    if (!src_planes_host || !dst_planes_host || !ij) return fail("xrs_gather_ij: null pointer");
    if (n_bands < 1) return fail("xrs_gather_ij: n_bands must be >= 1");
    if (method != XRS_NEAREST && method != XRS_BILINEAR && method != XRS_TRIANGULAR)
        return fail("interp_methods must be one of 0, 1, 'nearest', 'bilinear', 'triangular'");
    if (src_h < 1 || src_w < 1 || src_pitch < 1 || dst_h < 1 || dst_w < 1) return fail("xrs_gather_ij: bad shape");
    if (win_i0 < 0 || win_j0 < 0 || win_i0 >= src_w || win_j0 >= src_h) return fail("xrs_gather_ij: bad source window origin");
    if (ceil_div(dst_h, K2_BY) > 65535) return fail("xrs_gather_ij: target too tall for one launch");
    for (int b = 0; b < n_bands; ++b)
        if (!src_planes_host[b] || !dst_planes_host[b]) return fail("xrs_gather_ij: null plane pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    XRS_DISPATCH_DTYPE(dtype, T, return launch_gather<T>(src_planes_host, dst_planes_host, n_bands, src_h, src_w, src_pitch, win_i0, win_j0, ij, dst_h, dst_w, method, fill, st));
    return 0;
}

}  // extern "C"

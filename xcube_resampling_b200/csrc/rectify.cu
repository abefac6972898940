// rectify.cu -- sm_100a kernels of the rectification path.
//
//   K0  xrs_tile_src_bboxes   gridmapping/bboxes.py:28-106   (one O(S) pass)
//   (K1, the source-index image, lives in rectify_ij.cu)
//   (K2, the gather through the ij image, lives in gather.cu)
//
// Compiled with -fmad=false; the parity-critical expressions additionally use
// explicit round-to-nearest intrinsics so they round like numba's LLVM code.
#include <cstdlib>
#include "common.cuh"

namespace xrs {

// ===========================================================================
// K0: per-tile source windows
// ===========================================================================
constexpr int K0_THREADS = 256;       // one source column per thread
constexpr int K0_ROWS = 32;           // consecutive source rows marched by one block
constexpr int K0_UNROLL = 4;          // rows in flight per thread
constexpr int K0_SMEM_TILES = 2048;   // tile table kept in shared memory up to this many tiles

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

// The tiles whose (border-grown) interval contains v form the contiguous index range [a, b]
// (empty if a > b).  Every v' in [lower, upper] has the same range, which lets a thread that
// marches down a source column reuse the search result until the column leaves the interval.
struct AxisRange {
    int a, b;
    double lower, upper;
};
__device__ __forceinline__ AxisRange locate(const double *lo, const double *hi, int n, double v) {
    AxisRange r;
    if (v != v) {  // NaN never matches (bboxes.py:60-69 comparisons are false)
        r.a = 1; r.b = 0; r.lower = INFINITY; r.upper = -INFINITY;
        return r;
    }
    r.b = count_le(lo, n, v) - 1;
    r.a = count_lt(hi, n, v);
    const double l1 = r.b >= 0 ? lo[r.b] : -INFINITY;
    const double l2 = r.a > 0 ? nextafter(hi[r.a - 1], INFINITY) : -INFINITY;
    const double u1 = r.b + 1 < n ? nextafter(lo[r.b + 1], -INFINITY) : INFINITY;
    const double u2 = r.a < n ? hi[r.a] : INFINITY;
    r.lower = fmax(l1, l2);
    r.upper = fmin(u1, u2);
    return r;
}

// x axis arrays ascend with tx; the y arrays may descend with ty (j-axis-down grids).
// Block = K0_THREADS consecutive source columns x K0_ROWS consecutive rows; each thread marches
// down its column, accumulating the row span per cached tile range in registers.
// MINFORM (shared-memory table only): the block's entries go straight into a min-form exchange table
// (maxima negated, see below) that the caller initialised to INT32_MAX -- no table of its own to
// initialise and fold afterwards.
template <bool SMEM_TABLE, bool MINFORM = false>
__global__ void __launch_bounds__(K0_THREADS)
k0_tile_windows(const double *__restrict__ x, const double *__restrict__ y, int64_t h, int64_t w, int64_t pitch,
                const double *__restrict__ g_x_lo, const double *__restrict__ g_x_hi, int ntx,
                const double *__restrict__ g_y_lo, const double *__restrict__ g_y_hi, int nty,
                int4 *__restrict__ table, int j_offset, int rows_per_block) {
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

    const int64_t n_col_chunks = ceil_div(w, K0_THREADS);
    const int64_t chunk = blockIdx.x % n_col_chunks, row_chunk = blockIdx.x / n_col_chunks;
    const int64_t col = chunk * K0_THREADS + threadIdx.x;
    const int64_t row0 = row_chunk * rows_per_block, row1 = min(row0 + rows_per_block, h);
    AxisRange rx, ry;
    rx.a = ry.a = 1; rx.b = ry.b = 0;
    rx.lower = ry.lower = INFINITY; rx.upper = ry.upper = -INFINITY;  // nothing cached yet
    int j_first = -1, j_last = -1;
    const int ci = static_cast<int>(col);
    if (col < w) {
        auto flush = [&]() {
            if (j_first < 0 || rx.a > rx.b || ry.a > ry.b) return;
            for (int ky = ry.a; ky <= ry.b; ++ky) {
                const int ty = y_desc ? nty - 1 - ky : ky;
                for (int tx = rx.a; tx <= rx.b; ++tx) {
                    int *e = reinterpret_cast<int *>(tab + ty * ntx + tx);
                    atomicMin(e + 0, ci);
                    atomicMin(e + 1, j_first + j_offset);
                    atomicMax(e + 2, ci + 1);
                    atomicMax(e + 3, j_last + 1 + j_offset);
                }
            }
        };
        for (int64_t j = row0; j < row1; j += K0_UNROLL) {
            double vx[K0_UNROLL], vy[K0_UNROLL];
#pragma unroll
            for (int k = 0; k < K0_UNROLL; ++k) {
                const bool in = j + k < row1;
                vx[k] = in ? ld_stream(x + (j + k) * pitch + col) : NAN;
                vy[k] = in ? ld_stream(y + (j + k) * pitch + col) : NAN;
            }
#pragma unroll
            for (int k = 0; k < K0_UNROLL; ++k) {
                if (j + k >= row1) break;
                const bool hit = vx[k] >= rx.lower && vx[k] <= rx.upper && vy[k] >= ry.lower && vy[k] <= ry.upper;
                if (!hit) {
                    flush();
                    rx = locate(s_x_lo, s_x_hi, ntx, vx[k]);
                    ry = locate(s_y_lo, s_y_hi, nty, vy[k]);
                    j_first = -1;
                }
                if (rx.a <= rx.b && ry.a <= ry.b) {
                    if (j_first < 0) j_first = static_cast<int>(j + k);
                    j_last = static_cast<int>(j + k);
                }
            }
        }
    }
    // The last span of every thread: neighbouring columns nearly always sit in the same tile range with
    // the same row span, so the lanes of a warp that agree on the range combine their spans and ONE of them
    // updates the table (the per-thread version serialises 4 same-address atomics per lane and tile).
    {
        const bool have = col < w && j_first >= 0 && rx.a <= rx.b && ry.a <= ry.b;
        const unsigned long long key = !have ? ~0ull
            : static_cast<unsigned long long>(rx.a) | static_cast<unsigned long long>(rx.b) << 16 |
              static_cast<unsigned long long>(ry.a) << 32 | static_cast<unsigned long long>(ry.b) << 48;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (have) {
            const int c_lo = __reduce_min_sync(peers, ci), c_hi = __reduce_max_sync(peers, ci + 1);
            const int j_lo = __reduce_min_sync(peers, j_first + j_offset);
            const int j_hi = __reduce_max_sync(peers, j_last + 1 + j_offset);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) {
                for (int ky = ry.a; ky <= ry.b; ++ky) {
                    const int ty = y_desc ? nty - 1 - ky : ky;
                    for (int tx = rx.a; tx <= rx.b; ++tx) {
                        int *e = reinterpret_cast<int *>(tab + ty * ntx + tx);
                        atomicMin(e + 0, c_lo);
                        atomicMin(e + 1, j_lo);
                        atomicMax(e + 2, c_hi);
                        atomicMax(e + 3, j_hi);
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
                if (MINFORM) {
                    atomicMin(g + 2, -e.z);
                    atomicMin(g + 3, -e.w);
                } else {
                    atomicMax(g + 2, e.z);
                    atomicMax(g + 3, e.w);
                }
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

// "Min-form" tables: every entry is merged across partial scans (row slabs of the swath handled by
// different GPUs) with MIN alone -- maxima are stored negated -- so that one all-reduce(MIN) or one
// element-wise minimum on the host combines them.  A tile entry is (i_min, j_min, -i_max1, -j_max1);
// INT32_MAX everywhere = no source point seen.
__global__ void k0_fold_minform(const int4 *__restrict__ table, int n, int32_t *__restrict__ minform) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int4 e = table[t];
    if (e.z < 0) return;
    atomicMin(minform + 4 * t + 0, e.x);
    atomicMin(minform + 4 * t + 1, e.y);
    atomicMin(minform + 4 * t + 2, -e.z);
    atomicMin(minform + 4 * t + 3, -e.w);
}

__global__ void k0_finalize_minform(const int32_t *__restrict__ minform, int n, int ij_border, int64_t w, int64_t h,
                                    int64_t *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t b0 = -1, b1 = -1, b2 = -1, b3 = -1;
    if (minform[4 * t] != INT32_MAX) {
        b0 = minform[4 * t]; b1 = minform[4 * t + 1]; b2 = -static_cast<int64_t>(minform[4 * t + 2]);
        b3 = -static_cast<int64_t>(minform[4 * t + 3]);
        if (ij_border != 0) {  // bboxes.py:90-106
            b0 -= ij_border; b1 -= ij_border; b2 += ij_border; b3 += ij_border;
            if (b0 < 0) b0 = 0;
            if (b1 < 0) b1 = 0;
            if (b2 > w) b2 = w;
            if (b3 > h) b3 = h;
        }
    }
    out[4 * t + 0] = b0; out[4 * t + 1] = b1; out[4 * t + 2] = b2; out[4 * t + 3] = b3;
}

}  // namespace xrs

using namespace xrs;

extern "C" {

int64_t xrs_tile_src_bboxes_workspace_bytes(int32_t ntx, int32_t nty) {
    return static_cast<int64_t>(ntx) * nty * static_cast<int64_t>(sizeof(int4));
}

static int k0_scan(const char *who, const double *x, const double *y, int64_t src_h, int64_t src_w, int64_t src_pitch,
                   int64_t j_offset, const double *x_lo, const double *x_hi, int32_t ntx, const double *y_lo,
                   const double *y_hi, int32_t nty, int4 *table, cudaStream_t st, int32_t *minform = nullptr) {
    const std::string w(who);
    if (!x || !y || !x_lo || !x_hi || !y_lo || !y_hi || !table) return fail(w + ": null pointer");
    if (src_h < 1 || src_w < 1 || src_pitch < src_w) return fail(w + ": bad source shape");
    if (ntx < 1 || nty < 1 || ntx > 65535 || nty > 65535) return fail(w + ": tile counts must be in [1, 65535]");
    if (src_w > INT32_MAX - 1 || src_h + j_offset > INT32_MAX - 1 || j_offset < 0) return fail(w + ": source too large");
    const int n_tiles = ntx * nty;
    if (minform && n_tiles <= K0_SMEM_TILES) {  // straight into the caller's (initialised) exchange table
        const int rows = src_h >= 2048 ? K0_ROWS : 8;
        const int64_t blocks = ceil_div(src_w, K0_THREADS) * ceil_div(src_h, rows);
        if (blocks > 0x7fffffffLL) return fail(w + ": source too large");
        const size_t smem = static_cast<size_t>(2 * ntx + 2 * nty) * sizeof(double) + static_cast<size_t>(n_tiles) * sizeof(int4);
        XRS_CUDA(cudaFuncSetAttribute(k0_tile_windows<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        XRS_TIMED("k0_tile_windows", st, (k0_tile_windows<true, true><<<static_cast<unsigned>(blocks), K0_THREADS, smem, st>>>(x, y, src_h, src_w, src_pitch, x_lo, x_hi, ntx, y_lo, y_hi, nty, reinterpret_cast<int4 *>(minform), static_cast<int>(j_offset), rows)));
        XRS_LAUNCH_CHECK("k0_tile_windows");
        return 1 << 30;  // done, nothing to fold (positive: not an error code)
    }
    XRS_TIMED("k0_init_table", st, k0_init_table<<<static_cast<unsigned>(ceil_div(n_tiles, 256)), 256, 0, st>>>(table, n_tiles));
    XRS_LAUNCH_CHECK("k0_init_table");

    // a short slab (one GPU's share of the swath) is cut into shorter row chunks: more blocks, shorter chains
    const int rows_per_block = src_h >= 2048 ? K0_ROWS : 8;
    const int64_t n_blocks = ceil_div(src_w, K0_THREADS) * ceil_div(src_h, rows_per_block);
    if (n_blocks > 0x7fffffffLL) return fail(w + ": source too large");
    const unsigned grid = static_cast<unsigned>(n_blocks);
    const size_t axis_bytes = static_cast<size_t>(2 * ntx + 2 * nty) * sizeof(double);
    const int joff = static_cast<int>(j_offset);
    if (n_tiles <= K0_SMEM_TILES) {
        const size_t smem = axis_bytes + static_cast<size_t>(n_tiles) * sizeof(int4);
        XRS_CUDA(cudaFuncSetAttribute(k0_tile_windows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        XRS_TIMED("k0_tile_windows", st, k0_tile_windows<true><<<grid, K0_THREADS, smem, st>>>(x, y, src_h, src_w, src_pitch, x_lo, x_hi, ntx, y_lo, y_hi, nty, table, joff, rows_per_block));
    } else {
        if (axis_bytes > 200 * 1024) return fail(w + ": too many tile rows/columns");
        XRS_CUDA(cudaFuncSetAttribute(k0_tile_windows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(axis_bytes)));
        XRS_TIMED("k0_tile_windows", st, k0_tile_windows<false><<<grid, K0_THREADS, axis_bytes, st>>>(x, y, src_h, src_w, src_pitch, x_lo, x_hi, ntx, y_lo, y_hi, nty, table, joff, rows_per_block));
    }
    XRS_LAUNCH_CHECK("k0_tile_windows");
    return 0;
}

int xrs_tile_src_bboxes(const double *x, const double *y, int64_t src_h, int64_t src_w, int64_t src_pitch,
                        const double *x_lo, const double *x_hi, int32_t ntx, const double *y_lo,
                        const double *y_hi, int32_t nty, int32_t ij_border, int64_t *out_boxes, void *workspace,
                        void *stream) {
    if (!out_boxes || !workspace) return fail("xrs_tile_src_bboxes: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int4 *table = static_cast<int4 *>(workspace);
    if (int rc = k0_scan("xrs_tile_src_bboxes", x, y, src_h, src_w, src_pitch, 0, x_lo, x_hi, ntx, y_lo, y_hi, nty, table, st))
        return rc;
    const int n_tiles = ntx * nty;
    XRS_TIMED("k0_finalize", st, k0_finalize<<<static_cast<unsigned>(ceil_div(n_tiles, 256)), 256, 0, st>>>(table, n_tiles, ij_border, src_w, src_h, out_boxes));
    XRS_LAUNCH_CHECK("k0_finalize");
    return 0;
}

int xrs_tile_src_bboxes_partial(const double *x, const double *y, int64_t slab_h, int64_t src_w, int64_t src_pitch,
                                int64_t j_offset, const double *x_lo, const double *x_hi, int32_t ntx,
                                const double *y_lo, const double *y_hi, int32_t nty, int32_t *minform_table,
                                void *workspace, void *stream) {
    if (!minform_table || !workspace) return fail("xrs_tile_src_bboxes_partial: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int4 *table = static_cast<int4 *>(workspace);
    if (int rc = k0_scan("xrs_tile_src_bboxes_partial", x, y, slab_h, src_w, src_pitch, j_offset, x_lo, x_hi, ntx, y_lo,
                         y_hi, nty, table, st, minform_table))
        return rc == 1 << 30 ? 0 : rc;
    const int n_tiles = ntx * nty;
    XRS_TIMED("k0_fold_minform", st, k0_fold_minform<<<static_cast<unsigned>(ceil_div(n_tiles, 256)), 256, 0, st>>>(table, n_tiles, minform_table));
    XRS_LAUNCH_CHECK("k0_fold_minform");
    return 0;
}

int xrs_tile_src_bboxes_finalize(const int32_t *minform_table, int32_t n_tiles, int32_t ij_border, int64_t src_w,
                                 int64_t src_h, int64_t *out_boxes, void *stream) {
    if (!minform_table || !out_boxes) return fail("xrs_tile_src_bboxes_finalize: null pointer");
    if (n_tiles < 1) return fail("xrs_tile_src_bboxes_finalize: n_tiles must be >= 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    XRS_TIMED("k0_finalize", st, k0_finalize_minform<<<static_cast<unsigned>(ceil_div(n_tiles, 256)), 256, 0, st>>>(minform_table, n_tiles, ij_border, src_w, src_h, out_boxes));
    XRS_LAUNCH_CHECK("k0_finalize_minform");
    return 0;
}

}  // extern "C"

// bands.cu -- target row bands across GPUs: what a band needs from the source, and how it gets there.
//
// The reference fans rectification out over target tiles (rectify.py:263-309, dask.py:41-135) and
// every tile task slices its source window out of the whole coordinate / data arrays
// (rectify.py:397-399).  Here one GPU owns one target ROW BAND and must be handed, over PCIe, only
// the part of the swath that can reach it:
//
//   xrs_band_quad_footprints   per band and per group of K1S_ROWS source quad rows, the range of quad
//                              columns whose pixel box touches the band -- a ragged (diagonal, for a
//                              rotated swath) footprint instead of the bounding box of tile windows.
//                              Scans a row slab of the swath; slabs scanned on different GPUs are
//                              merged with MIN (maxima stored negated, see rectify.cu "min-form").
//   xrs_minform_init           fill a min-form table with "nothing seen"
//   xrs_copy2d_slices          strided host<->device copies of (slices, rows, width) windows on the
//                              copy engines (ragged footprint uploads, row-band downloads into one
//                              host array)
#include "rectify_common.cuh"

namespace xrs {

constexpr int KB_THREADS = 256;
constexpr int KB_COLS = 64;  // quad columns per block
constexpr int KB_MAX_BANDS = 64;
static_assert(K1S_ROWS % (KB_THREADS / KB_COLS) == 0, "row groups must split evenly over the thread rows");

__global__ void kb_fill_i32(int32_t *p, int64_t n, int32_t v) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct BandGrid {
    double x_min, y_min, y_max, inv_xr, inv_yr;
    int j_up;
    int64_t dst_w, dst_h;
};

// One block = KB_COLS quad columns x K1S_ROWS quad rows (one row group); a thread marches down a
// quarter of its quad column.  A quad counts for every band its pixel box (grown by a margin that covers the
// barycentric tolerance and the +-1 px between tile-local and global pixel arithmetic) overlaps.
// Quads with fewer than three finite vertices cannot accept a pixel (both triangles hold a NaN).
__global__ void __launch_bounds__(KB_THREADS)
kb_quad_footprints(const double *__restrict__ x, const double *__restrict__ y, int64_t slab_h, int64_t src_w,
                   int64_t pitch, int64_t j_offset, BandGrid g, const int32_t *__restrict__ band_edges, int n_bands,
                   int n_groups, int32_t *__restrict__ fp) {
    __shared__ int s_edges[KB_MAX_BANDS + 1];
    __shared__ int s_tab[KB_MAX_BANDS][2];
    for (int k = threadIdx.x; k <= n_bands; k += blockDim.x) s_edges[k] = band_edges[k];
    for (int k = threadIdx.x; k < n_bands; k += blockDim.x) { s_tab[k][0] = INT32_MAX; s_tab[k][1] = INT32_MAX; }
    __syncthreads();
    const int64_t nqi = src_w - 1;
    const int64_t n_col_chunks = ceil_div(nqi, KB_COLS);
    const int64_t chunk = blockIdx.x % n_col_chunks, row_chunk = blockIdx.x / n_col_chunks;
    const int64_t qi = chunk * KB_COLS + (threadIdx.x % KB_COLS);
    // the K1S_ROWS quad rows of the group are split over KB_THREADS / KB_COLS thread rows (short
    // dependent-load chains, four times as many blocks as one thread per column would give)
    constexpr int ROWS_PER_THREAD = K1S_ROWS / (KB_THREADS / KB_COLS);
    const int64_t g0 = row_chunk * K1S_ROWS;
    const int64_t q0 = g0 + (threadIdx.x / KB_COLS) * ROWS_PER_THREAD;
    const int64_t q1 = min(min(q0 + ROWS_PER_THREAD, g0 + K1S_ROWS), slab_h - 1);  // local quad rows
    // a thread's quads all lie in ONE column: it only has to remember which bands they touch (bit b)
    unsigned long long touched = 0;
    if (qi < nqi && q0 < q1) {
        // all vertex rows of the thread's quads are requested before any is used (one memory round trip
        // instead of one per row: the kernel is latency-bound on slabs of a few hundred rows)
        double vx[ROWS_PER_THREAD + 1][2], vy[ROWS_PER_THREAD + 1][2];
#pragma unroll
        for (int k = 0; k <= ROWS_PER_THREAD; ++k) {
            const bool in = q0 + k <= q1;
            const int64_t o = (q0 + k) * pitch + qi;
            vx[k][0] = in ? __ldg(x + o) : NAN; vx[k][1] = in ? __ldg(x + o + 1) : NAN;
            vy[k][0] = in ? __ldg(y + o) : NAN; vy[k][1] = in ? __ldg(y + o + 1) : NAN;
        }
        // One vertex row = the pair (qi, qi + 1): its pixel-space box and its number of usable vertices.
        // A vertex that is not usable (NaN / inf) becomes NaN, which fmin / fmax skip, so a quad's box is
        // the box of its two vertex rows and every row box is computed once for the two quads sharing it.
        struct RowBox { double lo_x, hi_x, lo_y, hi_y; int n_ok; };
        auto row_box = [&](int k) {
            double fx0 = (vx[k][0] - g.x_min) * g.inv_xr, fx1 = (vx[k][1] - g.x_min) * g.inv_xr;
            double fy0 = g.j_up ? (vy[k][0] - g.y_min) * g.inv_yr : (g.y_max - vy[k][0]) * g.inv_yr;
            double fy1 = g.j_up ? (vy[k][1] - g.y_min) * g.inv_yr : (g.y_max - vy[k][1]) * g.inv_yr;
            const bool ok0 = fabs(fx0) < 1e15 && fabs(fy0) < 1e15, ok1 = fabs(fx1) < 1e15 && fabs(fy1) < 1e15;
            if (!ok0) fx0 = fy0 = NAN;
            if (!ok1) fx1 = fy1 = NAN;
            RowBox r;
            r.lo_x = fmin(fx0, fx1); r.hi_x = fmax(fx0, fx1);
            r.lo_y = fmin(fy0, fy1); r.hi_y = fmax(fy0, fy1);
            r.n_ok = static_cast<int>(ok0) + static_cast<int>(ok1);
            return r;
        };
        // the bands a row interval [r_lo, r_hi] overlaps are b_lo .. b_hi with
        //   b_lo = first band whose end edge lies above r_lo,  b_hi = last band whose start edge is <= r_hi
        // (edges ascend); both stay valid while r_lo / r_hi stay inside the cached edge intervals, which
        // they do from one quad of a column to the next almost always
        int b_lo = 0, b_hi = -1;
        int lo_from = 1, lo_to = 0, hi_from = 1, hi_to = 0;  // empty: the first quad searches
        RowBox up = row_box(0);
#pragma unroll
        for (int k = 0; k < ROWS_PER_THREAD; ++k) {
            if (q0 + k >= q1) break;
            const RowBox dn = row_box(k + 1);
            if (up.n_ok + dn.n_ok >= 3) {
                double lo_x = fmin(up.lo_x, dn.lo_x), hi_x = fmax(up.hi_x, dn.hi_x);
                double lo_y = fmin(up.lo_y, dn.lo_y), hi_y = fmax(up.hi_y, dn.hi_y);
                const double mrg = 2.0 + 0.01 * fmax(hi_x - lo_x, hi_y - lo_y);
                lo_x = floor(lo_x) - mrg; hi_x = floor(hi_x) + mrg;
                lo_y = floor(lo_y) - mrg; hi_y = floor(hi_y) + mrg;
                if (hi_x >= 0.0 && lo_x < static_cast<double>(g.dst_w) && hi_y >= 0.0 &&
                    lo_y < static_cast<double>(g.dst_h)) {
                    const int r_lo = static_cast<int>(fmax(lo_y, 0.0));
                    const int r_hi = static_cast<int>(fmin(hi_y, static_cast<double>(g.dst_h - 1)));
                    if (r_lo < lo_from || r_lo >= lo_to) {
                        b_lo = 0;
                        while (b_lo < n_bands && s_edges[b_lo + 1] <= r_lo) ++b_lo;
                        lo_from = b_lo > 0 ? s_edges[b_lo] : INT32_MIN;
                        lo_to = b_lo < n_bands ? s_edges[b_lo + 1] : INT32_MAX;
                    }
                    if (r_hi < hi_from || r_hi >= hi_to) {
                        b_hi = n_bands - 1;
                        while (b_hi >= 0 && s_edges[b_hi] > r_hi) --b_hi;
                        hi_from = b_hi >= 0 ? s_edges[b_hi] : INT32_MIN;
                        hi_to = b_hi + 1 < n_bands ? s_edges[b_hi + 1] : INT32_MAX;
                    }
                    if (b_lo <= b_hi) touched |= (~0ull >> (63 - b_hi)) & (~0ull << b_lo);
                }
            }
            up = dn;
        }
    }
    // per band: the warp's lowest and highest touching column (lanes = consecutive columns), one pair of
    // shared atomics per warp and band instead of one per quad
    {
        const int ci = static_cast<int>(qi);
        unsigned long long any = touched;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, o);
        while (any) {
            const int b = __ffsll(static_cast<long long>(any)) - 1;
            any &= any - 1;
            if (s_edges[b] >= s_edges[b + 1]) continue;  // an empty band between two others (warp-uniform)
            const unsigned lanes = __ballot_sync(0xffffffffu, (touched >> b) & 1ull);
            const int lane = threadIdx.x & 31;
            const int first = __ffs(lanes) - 1, last = 31 - __clz(lanes);
            if (lane == first) atomicMin(&s_tab[b][0], ci);
            if (lane == last) atomicMin(&s_tab[b][1], -ci);
        }
    }
    __syncthreads();
    const int64_t group = j_offset / K1S_ROWS + row_chunk;
    if (group < n_groups)
        for (int b = threadIdx.x; b < n_bands; b += blockDim.x)
            if (s_tab[b][0] != INT32_MAX) {
                int32_t *e = fp + (static_cast<int64_t>(b) * n_groups + group) * 2;
                atomicMin(e, s_tab[b][0]);
                atomicMin(e + 1, s_tab[b][1]);
            }
}

}  // namespace xrs

using namespace xrs;

extern "C" {

int32_t xrs_quad_row_group(void) { return K1S_ROWS; }

int xrs_minform_init(int32_t *table, int64_t n, void *stream) {
    if (!table || n < 1) return fail("xrs_minform_init: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    XRS_TIMED("kb_fill_i32", st, kb_fill_i32<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, st>>>(table, n, INT32_MAX));
    XRS_LAUNCH_CHECK("kb_fill_i32");
    return 0;
}

int xrs_band_quad_footprints(const double *x, const double *y, int64_t slab_h, int64_t src_w, int64_t src_pitch,
                             int64_t j_offset, int64_t src_h, int64_t dst_h, int64_t dst_w, double x_min, double y_min,
                             double y_max, double x_res, double y_res, int32_t is_j_axis_up,
                             const int32_t *band_edges, int32_t n_bands, int32_t *minform_fp, void *stream) {
    if (!x || !y || !band_edges || !minform_fp) return fail("xrs_band_quad_footprints: null pointer");
    if (slab_h < 2 || src_w < 2 || src_pitch < src_w || src_h < 2) return fail("xrs_band_quad_footprints: slab must hold at least 2x2 vertices");
    if (j_offset < 0 || j_offset % K1S_ROWS != 0 || j_offset + slab_h > src_h)
        return fail("xrs_band_quad_footprints: slab must start on a multiple of the quad row group and lie inside the source");
    if (n_bands < 1 || n_bands > KB_MAX_BANDS) return fail("xrs_band_quad_footprints: 1..64 bands");
    if (!(x_res > 0.0) || !(y_res > 0.0) || dst_h < 1 || dst_w < 1) return fail("xrs_band_quad_footprints: bad target grid");
    if (src_w > INT32_MAX - 1 || src_h > INT32_MAX - 1) return fail("xrs_band_quad_footprints: source too large");
    BandGrid g;
    g.x_min = x_min; g.y_min = y_min; g.y_max = y_max; g.inv_xr = 1.0 / x_res; g.inv_yr = 1.0 / y_res;
    g.j_up = is_j_axis_up ? 1 : 0; g.dst_w = dst_w; g.dst_h = dst_h;
    const int n_groups = static_cast<int>(ceil_div(src_h - 1, K1S_ROWS));
    const int64_t n_blocks = ceil_div(src_w - 1, KB_COLS) * ceil_div(slab_h - 1, K1S_ROWS);
    if (n_blocks > 0x7fffffffLL) return fail("xrs_band_quad_footprints: slab too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    XRS_TIMED("kb_quad_footprints", st, kb_quad_footprints<<<static_cast<unsigned>(n_blocks), KB_THREADS, 0, st>>>(x, y, slab_h, src_w, src_pitch, j_offset, g, band_edges, n_bands, n_groups, minform_fp));
    XRS_LAUNCH_CHECK("kb_quad_footprints");
    return 0;
}

int xrs_copy2d_slices(void *dst, int64_t dst_pitch_bytes, int64_t dst_slice_bytes, const void *src,
                      int64_t src_pitch_bytes, int64_t src_slice_bytes, int64_t width_bytes, int64_t rows,
                      int64_t slices, void *stream) {
    if (!dst || !src) return fail("xrs_copy2d_slices: null pointer");
    if (width_bytes < 0 || rows < 0 || slices < 0) return fail("xrs_copy2d_slices: negative extent");
    if (width_bytes == 0 || rows == 0 || slices == 0) return 0;
    if (dst_pitch_bytes < width_bytes || src_pitch_bytes < width_bytes) return fail("xrs_copy2d_slices: pitch smaller than the row width");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int64_t s = 0; s < slices; ++s) {
        char *d = static_cast<char *>(dst) + s * dst_slice_bytes;
        const char *p = static_cast<const char *>(src) + s * src_slice_bytes;
        if (dst_pitch_bytes == width_bytes && src_pitch_bytes == width_bytes) {
            XRS_CUDA(cudaMemcpyAsync(d, p, static_cast<size_t>(width_bytes) * rows, cudaMemcpyDefault, st));
        } else {
            XRS_CUDA(cudaMemcpy2DAsync(d, static_cast<size_t>(dst_pitch_bytes), p, static_cast<size_t>(src_pitch_bytes),
                                       static_cast<size_t>(width_bytes), static_cast<size_t>(rows), cudaMemcpyDefault, st));
        }
    }
    return 0;
}

}  // extern "C"

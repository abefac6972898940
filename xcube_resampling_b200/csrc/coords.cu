// coords.cu -- statistics of 2-D coordinate images on the device, so that a GridMapping can be
// derived from coordinates that never leave HBM (CRS-transformed or pre-downscaled swath
// coordinates inside rectify_dataset).
//
//   xrs_coords_stats   gridmapping/coords.py:102-103 (any(x > 180)) and :226-252 (cell areas of an
//                      irregular grid: the resolution estimate takes the square roots of the
//                      smallest and the largest positive area)
//   xrs_lon_360        gridmapping/helpers.py to_lon_360 (x < 0 -> x + 360), in place
//
// The few 1-D slices the derivation also needs (first / last rows and columns: regularity test,
// bounding box, axis direction) are fetched by the host; the images themselves are not.
#include "common.cuh"

namespace xrs {

constexpr int KC_THREADS = 256;
constexpr double KC_ER = 6371000.0;  // coords.py:46

// fabs(a) with NaN and |a| <= 1e-8 (np.isclose(a, 0)) mapped to 0 -- coords.py:340-342 _abs_no_nan
__device__ __forceinline__ double abs_no_nan(double a) {
    a = fabs(a);
    return (a != a || a <= 1e-8) ? 0.0 : a;
}

// out: [0] any(x > 180) as 0/1 (uint64), [1] min positive area, [2] max positive area, both as the
// bit patterns of positive doubles (ordered like unsigned integers); initialised by the host side.
__global__ void __launch_bounds__(KC_THREADS)
kc_coords_stats(const double *__restrict__ x, const double *__restrict__ y, int64_t h, int64_t w, int64_t pitch,
                int geographic, unsigned long long *__restrict__ out) {
    const int64_t n = h * w;
    unsigned long long a_min = ~0ull, a_max = 0ull;
    int gt180 = 0;
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n;
         p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t j = p / w, i = p - j * w;
        // coords.py:232-241: differences along x and y; the last column / row repeats its neighbour's
        const int64_t ic = i < w - 1 ? i : w - 2, jc = j < h - 1 ? j : h - 2;
        const double x00 = x[j * pitch + i];
        gt180 |= x00 > 180.0;
        const double x_x = abs_no_nan(dsub(x[j * pitch + ic + 1], x[j * pitch + ic]));
        const double y_x = abs_no_nan(dsub(y[j * pitch + ic + 1], y[j * pitch + ic]));
        const double x_y = abs_no_nan(dsub(x[(jc + 1) * pitch + i], x[jc * pitch + i]));
        const double y_y = abs_no_nan(dsub(y[(jc + 1) * pitch + i], y[jc * pitch + i]));
        // coords.py:243-250
        double x_abs = sqrt(dadd(dmul(x_x, x_x), dmul(x_y, x_y)));
        double y_abs = sqrt(dadd(dmul(y_x, y_x), dmul(y_y, y_y)));
        if (geographic) {
            const double xr = x_abs * (3.14159265358979323846 / 180.0), yr = y_abs * (3.14159265358979323846 / 180.0);
            x_abs = dmul(dmul(KC_ER, cos(xr)), yr);
            y_abs = dmul(KC_ER, yr);
        }
        const double area = dmul(x_abs, y_abs);
        if (area > 0.0 && area < INFINITY) {
            const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(area));
            a_min = min(a_min, bits);
            a_max = max(a_max, bits);
        }
    }
    __shared__ unsigned long long s_min, s_max;
    __shared__ int s_gt;
    if (threadIdx.x == 0) { s_min = ~0ull; s_max = 0ull; s_gt = 0; }
    __syncthreads();
    if (a_min != ~0ull) atomicMin(&s_min, a_min);
    if (a_max != 0ull) atomicMax(&s_max, a_max);
    if (gt180) atomicOr(&s_gt, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_gt) atomicOr(out + 0, 1ull);
        if (s_min != ~0ull) atomicMin(out + 1, s_min);
        if (s_max != 0ull) atomicMax(out + 2, s_max);
    }
}

__global__ void kc_stats_init(unsigned long long *out) {
    out[0] = 0ull; out[1] = ~0ull; out[2] = 0ull; out[3] = 0ull;
}

__global__ void kc_lon_360(double *x, int64_t h, int64_t w, int64_t pitch) {
    const int64_t n = h * w;
    for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n;
         p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t j = p / w, i = p - j * w;
        const double v = x[j * pitch + i];
        if (v < 0.0) x[j * pitch + i] = dadd(v, 360.0);
    }
}

}  // namespace xrs

using namespace xrs;

extern "C" {

int xrs_coords_stats(const double *x, const double *y, int64_t h, int64_t w, int64_t pitch, int32_t is_geographic,
                     uint64_t *out4, void *stream) {
    if (!x || !y || !out4) return fail("xrs_coords_stats: null pointer");
    if (h < 2 || w < 2 || pitch < w) return fail("xrs_coords_stats: coordinate images must be at least 2x2");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long *out = reinterpret_cast<unsigned long long *>(out4);
    XRS_TIMED("kc_stats_init", st, kc_stats_init<<<1, 1, 0, st>>>(out));
    XRS_LAUNCH_CHECK("kc_stats_init");
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(h * w, KC_THREADS), 148 * 8));
    XRS_TIMED("kc_coords_stats", st, kc_coords_stats<<<blocks, KC_THREADS, 0, st>>>(x, y, h, w, pitch, is_geographic ? 1 : 0, out));
    XRS_LAUNCH_CHECK("kc_coords_stats");
    return 0;
}

int xrs_lon_360(double *x, int64_t h, int64_t w, int64_t pitch, void *stream) {
    if (!x) return fail("xrs_lon_360: null pointer");
    if (h < 1 || w < 1 || pitch < w) return fail("xrs_lon_360: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(h * w, 256), 148 * 8));
    XRS_TIMED("kc_lon_360", st, kc_lon_360<<<blocks, 256, 0, st>>>(x, h, w, pitch));
    XRS_LAUNCH_CHECK("kc_lon_360");
    return 0;
}

}  // extern "C"

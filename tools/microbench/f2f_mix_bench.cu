// Issue rates of the instructions that bound the bilinear gathers (K2, K3) on sm_100a:
//   F2F.F64.F32 alone, DADD alone, the per-band-pixel mix of the kernels (4 x F2F.F64.F32 + 9 fp64
//   adds/multiplies + 1 x F2F.F32.F64), and the same mix with two of the four widenings done on the
//   integer pipe.  Answers: do the conversions share the fp64 pipe with DADD/DMUL (then the mix takes
//   the SUM of the parts) or run beside it (then the MAX)?
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o f2f_mix_bench f2f_mix_bench.cu && ./f2f_mix_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double widen_int(float f) {  // exact for normal floats
    const unsigned b = __float_as_uint(f);
    const unsigned hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
    return __hiloint2double(hi, b << 29);
}

template <int MODE>
__global__ void k(const float *in, float *out, int iters) {
    float f0 = in[threadIdx.x], f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
    const double u = 0.25 + 1e-3 * threadIdx.x, v = 0.5 + 1e-3 * threadIdx.x;
    double acc = 0.0;
    unsigned long long bits = 0;
    float facc = 0.f;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {         // 4 x F2F.F64.F32, results folded with integer xor (no fp64 arithmetic)
            bits ^= (unsigned long long)__double_as_longlong((double)f0) ^ (unsigned long long)__double_as_longlong((double)f1) ^
                    (unsigned long long)__double_as_longlong((double)f2) ^ (unsigned long long)__double_as_longlong((double)f3);
        } else if (MODE == 1) {  // 9 dependent-free fp64 operations
            double a = acc + u, b = acc * v, c = a + v, d = b * u, e = c + d, g = e * u, h = g + a, p = h * v;
            acc = p + c;
        } else if (MODE == 2 || MODE == 3) {  // the bilinear mix
            const double v00 = MODE == 3 ? widen_int(f0) : (double)f0, v01 = (double)f1;
            const double v10 = MODE == 3 ? widen_int(f2) : (double)f2, v11 = (double)f3;
            const double a = __dadd_rn(v00, __dmul_rn(u, __dsub_rn(v01, v00)));
            const double b = __dadd_rn(v10, __dmul_rn(u, __dsub_rn(v11, v10)));
            facc += (float)__dadd_rn(a, __dmul_rn(v, __dsub_rn(b, a)));
        }
        f0 = __int_as_float(__float_as_int(f0) ^ (i & 0xff)); f1 = __int_as_float(__float_as_int(f1) ^ (i & 0xff));
        f2 = __int_as_float(__float_as_int(f2) ^ (i & 0xff)); f3 = __int_as_float(__float_as_int(f3) ^ (i & 0xff));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc + (float)(bits & 0xffff) + facc;
}

template <int MODE>
float run(const float *in, float *out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(in, out, iters);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(in, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 4096);
    cudaMemset(in, 0x3f, 4096);
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    int clk_khz = 1965000;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const float t0 = run<0>(in, out, iters), t1 = run<1>(in, out, iters), t2 = run<2>(in, out, iters),
                t3 = run<3>(in, out, iters);
    const double warp_iters_per_sm = 8.0 * 8 * iters;  // 8 CTAs x 8 warps per SM
    auto clk = [&](float ms) { return ms * 1e-3 * clk_khz * 1e3 / warp_iters_per_sm; };
    printf("{\"clock_khz\": %d, \"clk_per_warp_iteration\": {\"4xF2F.F64.F32\": %.2f, \"9xfp64\": %.2f, "
           "\"bilinear_mix_4F2F_9fp64_1F2F\": %.2f, \"mix_with_2_int_widenings\": %.2f}}\n",
           clk_khz, clk(t0), clk(t1), clk(t2), clk(t3));
    return 0;
}

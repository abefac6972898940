#include <cstdio>
#include <cuda_runtime.h>
// throughput of F2F.F64.F32, F2F.F32.F64, DADD, int-emulated f32->f64 per SM
template <int MODE>
__global__ void k(float *in, double *out, int iters) {
    float f0 = in[threadIdx.x], f1 = f0 + 1.f, f2 = f0 + 2.f, f3 = f0 + 3.f;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // cvt f32->f64 + dadd
            a0 += (double)f0; a1 += (double)f1; a2 += (double)f2; a3 += (double)f3;
            f0 = __int_as_float(__float_as_int(f0) ^ i); f1 = __int_as_float(__float_as_int(f1) ^ i);
            f2 = __int_as_float(__float_as_int(f2) ^ i); f3 = __int_as_float(__float_as_int(f3) ^ i);
        } else if (MODE == 1) {  // dadd only (+ same int ops)
            a0 += 1.5; a1 += 2.5; a2 += 3.5; a3 += 4.5;
            f0 = __int_as_float(__float_as_int(f0) ^ i); f1 = __int_as_float(__float_as_int(f1) ^ i);
            f2 = __int_as_float(__float_as_int(f2) ^ i); f3 = __int_as_float(__float_as_int(f3) ^ i);
        } else {  // integer-emulated conversion of normal floats + dadd
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                float f = q == 0 ? f0 : q == 1 ? f1 : q == 2 ? f2 : f3;
                unsigned b = __float_as_uint(f);
                unsigned hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
                unsigned lo = b << 29;
                double d = __hiloint2double(hi, lo);
                if (q == 0) a0 += d; else if (q == 1) a1 += d; else if (q == 2) a2 += d; else a3 += d;
            }
            f0 = __int_as_float(__float_as_int(f0) ^ i); f1 = __int_as_float(__float_as_int(f1) ^ i);
            f2 = __int_as_float(__float_as_int(f2) ^ i); f3 = __int_as_float(__float_as_int(f3) ^ i);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
template <int MODE> float run(float *in, double *out, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(in, out, iters);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(in, out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float *in; double *out; cudaMalloc(&in, 4096); cudaMemset(in, 0x3f, 4096); cudaMalloc(&out, 148 * 8 * 256 * 8);
    int iters = 20000;
    float t0 = run<0>(in, out, iters), t1 = run<1>(in, out, iters), t2 = run<2>(in, out, iters);
    double warps = 148.0 * 8 * 8, n = warps * iters * 4;  // warp-level (cvt+dadd) pairs
    printf("cvt+dadd %.3f ms  dadd %.3f ms  intcvt+dadd %.3f ms\n", t0, t1, t2);
    printf("per SM per clk (1.9GHz): cvt+dadd %.3f warp-pairs/clk, dadd %.3f, intcvt %.3f\n",
           n / 148 / (t0 * 1e-3 * 1.9e9), n / 148 / (t1 * 1e-3 * 1.9e9), n / 148 / (t2 * 1e-3 * 1.9e9));
    return 0;
}

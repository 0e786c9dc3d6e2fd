// FP32 FMA-pipe peak probe: the measured denominator of the SIMT roofline.
#include "common.cuh"
#include "kernels.h"

namespace sk {

__global__ void __launch_bounds__(256) fp32_peak_kernel(float *sink, int iters) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(1.0f + threadIdx.x * 1e-6f + i, 0.5f + i);
    const float2 b = make_float2(1.000001f, 0.999999f);
    const float2 c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], b, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) sink[0] = s;
}

cudaError_t launch_fp32_peak(float *sink, int iters, int grid, cudaStream_t st) {
    fp32_peak_kernel<<<grid, 256, 0, st>>>(sink, iters);
    return cudaGetLastError();
}

}  // namespace sk

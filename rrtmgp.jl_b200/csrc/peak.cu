// FP32 CUDA-core peak of the device this library runs on, measured: the denominator of bench.py's `roofline_fp32`
// (BASELINE.md §2 asks for an FMA microbenchmark instead of the nominal 148 x 128 x 2 x clock).  Two figures: a
// stream of independent scalar FFMA (what the fused column kernels are made of) and of packed FFMA2 (fma.rn.f32x2).
#include <cuda_runtime.h>

#include "../../include/rrtmgp_b200.h"

namespace {

template <bool PACKED>
__global__ void __launch_bounds__(512) fma_stream_kernel(float* out, int iters, float s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    const float2 m = make_float2(s, s * 0.5f), c = make_float2(0.25f * s, 0.125f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (PACKED) a[i] = __ffma2_rn(a[i], m, c);
                else { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <bool PACKED> double measure(int sms, float* out) {
    const int threads = 512, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_stream_kernel<PACKED><<<sms, threads>>>(out, 64, 1.0001f);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        fma_stream_kernel<PACKED><<<sms, threads>>>(out, iters, 1.0001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * (double)sms * threads * iters * 64 * 2;   // 128 scalar FMA per thread and iteration
        if (ms > 0.f) best = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

}  // namespace

extern "C" int rrtmgp_b200_measure_fp32_peak(int32_t device, double* ffma_tflops, double* ffma2_tflops) {
    if (!ffma_tflops || !ffma2_tflops) return RRTMGP_B200_ERR_INVALID_ARG;
    int prev = 0, sms = 0;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return RRTMGP_B200_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float* out = nullptr;
    if (sms <= 0 || cudaMalloc(&out, (size_t)sms * 512 * sizeof(float)) != cudaSuccess) { cudaSetDevice(prev); return RRTMGP_B200_ERR_CUDA; }
    *ffma_tflops = measure<false>(sms, out);
    *ffma2_tflops = measure<true>(sms, out);
    cudaFree(out);
    const cudaError_t e = cudaGetLastError();
    cudaSetDevice(prev);
    return e == cudaSuccess ? RRTMGP_B200_OK : RRTMGP_B200_ERR_CUDA;
}

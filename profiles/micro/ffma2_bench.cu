// Microbenchmark: issue throughput of FFMA vs packed FFMA2/FADD2/FMUL2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(512) k(float* out, int iters, float s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    unsigned b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = threadIdx.x * 7 + i;
    const unsigned mi = __float_as_uint(s);
    const float2 m = make_float2(s, s * 0.5f), c = make_float2(0.25f * s, 0.125f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (KIND == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }   // 2 FFMA
                if (KIND == 1) a[i] = __ffma2_rn(a[i], m, c);                                              // 1 FFMA2
                if (KIND == 2) { a[i] = __fmul2_rn(a[i], m); a[i] = __fadd2_rn(a[i], c); }                 // FMUL2 + FADD2
                if (KIND == 4) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); b[i] = (b[i] ^ mi) + (b[i] >> 3); }
                if (KIND == 5) { a[i] = __ffma2_rn(a[i], m, c); b[i] = (b[i] ^ mi) + (b[i] >> 3); }
                if (KIND == 3) { a[i].x = a[i].x * m.x; a[i].y = a[i].y * m.y; a[i].x += c.x; a[i].y += c.y; }   // 2 FMUL + 2 FADD
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y + (KIND >= 4 ? (float)b[i] : 0.f);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int KIND> void run(const char* name, int warps_per_sm, double flop_per_inner) {
    int sms = 148, iters = 4096;
    float* out; cudaMalloc(&out, sms * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<sms, warps_per_sm * 32>>>(out, 16, 1.0001f);
    cudaEventRecord(e0);
    k<KIND><<<sms, warps_per_sm * 32>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double elem_ops = (double)sms * warps_per_sm * 32 * iters * 64 * 2;   // scalar mul-add pairs
    printf("%-22s warps/SM %2d: %.3f ms  %.2f T scalar-FMA-equiv/s  (%.1f per SM per ns)\n", name, warps_per_sm, ms,
           elem_ops / ms / 1e9, elem_ops / ms / 1e6 / sms);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 12, 16}) {
        run<0>("FFMA x2 (scalar)", w, 0);
        run<1>("FFMA2 (packed)", w, 0);
        run<2>("FMUL2+FADD2 (packed)", w, 0);
        run<3>("FMUL,FADD x2 (scalar)", w, 0);
        run<4>("2 FFMA + int ops", w, 0);
        run<5>("FFMA2 + int ops", w, 0);
    }
    return 0;
}

// Microbenchmark: does packed FFMA2 free issue slots in a MIXED instruction stream (FP + ALU + LDS) on sm_100a?
// Per unit of work: 4 scalar FMAs (or 2 FFMA2) + 1 integer op + 1 shared-memory load.
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(512) k(float* out, int iters, float s) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    float2 a[8], c[4];
    unsigned b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = threadIdx.x * 7 + i; c[i] = make_float2(0.1f * i, 0.2f * i); }
    const float2 m = make_float2(s, s * 0.5f);
    const unsigned mi = __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                b[i] = (b[i] ^ mi) + r;                                   // 1 ALU op (LOP3 / IADD3 fused?)
                float l = sm[(b[i] & 1023)];                              // LOP3 + LDS
                c[i].x += l;                                              // FADD
                if (KIND == 0) {
                    a[2 * i].x = fmaf(a[2 * i].x, m.x, c[i].x); a[2 * i].y = fmaf(a[2 * i].y, m.y, c[i].y);
                    a[2 * i + 1].x = fmaf(a[2 * i + 1].x, m.x, c[i].x); a[2 * i + 1].y = fmaf(a[2 * i + 1].y, m.y, c[i].y);
                } else {
                    a[2 * i] = __ffma2_rn(a[2 * i], m, c[i]);
                    a[2 * i + 1] = __ffma2_rn(a[2 * i + 1], m, c[i]);
                }
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + b[0] + b[1] + b[2] + b[3];
}

template <int KIND> void run(const char* name, int warps_per_sm) {
    int sms = 148, iters = 4096;
    float* out; cudaMalloc(&out, sms * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<sms, warps_per_sm * 32>>>(out, 16, 1.0001f);
    cudaEventRecord(e0);
    k<KIND><<<sms, warps_per_sm * 32>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double units = (double)warps_per_sm * iters * 32;   // work units per SM (warp level)
    printf("%-28s warps/SM %2d: %.3f ms  %.2f warp-units per SM per us\n", name, warps_per_sm, ms, units / ms / 1e3);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 12, 16}) {
        run<0>("4 FFMA + FADD + 2 ALU + LDS", w);
        run<1>("2 FFMA2 + FADD + 2 ALU + LDS", w);
    }
    return 0;
}

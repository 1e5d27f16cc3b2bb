// Microbenchmark: how fast can one SM pull small, data-dependent table segments (the k-distribution corner gathers of
// the fused column kernels) out of L2 -- per-lane LDG (round-1 design) vs cp.async.bulk (UBLKCP) vs tensor-map TMA
// (cp.async.bulk.tensor, UTMALDG) into a per-warp shared-memory ring that the lanes then read with LDS.
//
// Shape of the real problem (LW, one minor-slot group): per (layer, 32 g-points) a warp needs 2 bands x 2 KB =
// 4 KB: per band 8 corners x 128 B ({kmajor, pfrac} pairs of 16 g-points) + 4 corners x 256 B (four minor slots).
// Persistent CTA of 12 warps per SM; every warp runs its own chain of iterations, `STAGES` deep.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_gather_bench tma_gather_bench.cu
// (no -lcuda: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tensor4_g2s(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

constexpr int WARPS = 12;
constexpr int ITER_BYTES = 4096;      // per warp per iteration
// table geometry (floats): [p 60][T 14][eta 9][512] = LW {kmajor, pfrac} table, 15.5 MB
constexpr int NP = 60, NT = 14, NE = 9, ROW = 512;

// MODE 0: per-lane LDG (8 x LDG.64 + 4 x LDG.128 per lane, the round-1 gather)
// MODE 1: cp.async.bulk, NCOPY copies of CBYTES each per iteration, issued by lane 0
// MODE 2: tensor-map TMA, 4 boxes of (32 floats x 2 eta x 2 p x 1 T) = 512 B per band-major block + 4 boxes for the
//         "minor" half (same map here), issued by lane 0
template <int MODE, int CBYTES, int STAGES, bool CONSUME>
__global__ void __launch_bounds__(WARPS * 32, 1) k(const float* __restrict__ table, const __grid_constant__ CUtensorMap map,
                                                    int iters, float* out, unsigned long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[WARPS * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* ring = smem + (size_t)warp * STAGES * ITER_BYTES;
    uint64_t* bar = bars + warp * 4;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(bar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t rng = 12345u + blockIdx.x * 977u + warp * 131u;
    float acc = 0.f;
    const long long t0 = clock64();
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
            // two bands (half-warps); per band: two (p, T, eta) rows for the majors, two (T, eta) rows for the minors
            uint32_t r = rng;
            for (int i = 0; i < 4; ++i) lcg(rng);
            uint32_t h = r + (lane >> 4) * 7919u;
            const int jp = lcg(h) % (NP - 1), jt = lcg(h) % (NT - 1), je1 = lcg(h) % (NE - 1), je2 = lcg(h) % (NE - 1);
            const int g = (lane & 15) + 16 * ((lane >> 4) + 2 * (it & 7));
            const float2* pa = reinterpret_cast<const float2*>(table) + ((jp * NT + jt) * NE + je1) * (ROW / 2) + g;
            const float2* pb = reinterpret_cast<const float2*>(table) + ((jp * NT + jt + 1) * NE + je2) * (ROW / 2) + g;
            constexpr int KE = ROW / 2, KP = NT * NE * (ROW / 2);
            float2 c[8];
            c[0] = __ldg(pa); c[1] = __ldg(pa + KE); c[2] = __ldg(pa + KP); c[3] = __ldg(pa + KP + KE);
            c[4] = __ldg(pb); c[5] = __ldg(pb + KE); c[6] = __ldg(pb + KP); c[7] = __ldg(pb + KP + KE);
            const float4* ma = reinterpret_cast<const float4*>(table) + ((jt)*NE + je1) * (ROW / 2) + g;   // (reuses the table as the minor table)
            const float4* mb = reinterpret_cast<const float4*>(table) + ((jt + 1) * NE + je2) * (ROW / 2) + g;
            float4 m0 = __ldg(ma), m1 = __ldg(ma + ROW / 2), m2 = __ldg(mb), m3 = __ldg(mb + ROW / 2);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += c[i].x * 1.0001f + c[i].y;
            acc += m0.x + m0.w + m1.y + m1.z + m2.x + m2.w + m3.y + m3.z;
        }
    } else {
        constexpr int NCOPY = ITER_BYTES / CBYTES;
        auto issue = [&](int it) {
            const int s = it % STAGES;
            if (lane == 0) {
                mbar_expect_tx(bar + s, ITER_BYTES);
                unsigned char* dst = ring + s * ITER_BYTES;
                if (MODE == 1) {
#pragma unroll 4
                    for (int c = 0; c < NCOPY; ++c) {
                        const uint32_t off = (lcg(rng) % (uint32_t)((size_t)NP * NT * NE * ROW * 4 / CBYTES)) * CBYTES;
                        bulk_g2s(dst + c * CBYTES, reinterpret_cast<const unsigned char*>(table) + off, CBYTES, bar + s);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {   // 8 boxes of 512 B
                        const int jp = lcg(rng) % (NP - 1), jt = lcg(rng) % NT, je = lcg(rng) % (NE - 1), gb = (lcg(rng) & 15) * 32;
                        tensor4_g2s(dst + c * 512, &map, gb, je, jp, jt, bar + s);
                    }
                }
            }
        };
        for (int it = 0; it < STAGES - 1 && it < iters; ++it) issue(it);
        for (int it = 0; it < iters; ++it) {
            if (it + STAGES - 1 < iters) issue(it + STAGES - 1);
            const int s = it % STAGES;
            mbar_wait(bar + s, (it / STAGES) & 1);
            if (CONSUME) {   // the lanes read the 4 KB once: 8 x LDS.64 + 4 x LDS.128 per lane
                const unsigned char* src = ring + s * ITER_BYTES;
                const float2* c2 = reinterpret_cast<const float2*>(src) + lane;
                const float4* c4 = reinterpret_cast<const float4*>(src + 2048) + lane;
#pragma unroll
                for (int i = 0; i < 8; ++i) { float2 v = c2[i * 32]; acc += v.x * 1.0001f + v.y; }
#pragma unroll
                for (int i = 0; i < 4; ++i) { float4 v = c4[i * 32]; acc += v.x + v.w; }
            }
            __syncwarp();   // every lane has consumed stage s before lane 0 re-arms it (next iteration's issue)
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int CBYTES, int STAGES, bool CONSUME>
void run(const char* name, const float* table, const CUtensorMap& map, int iters) {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float* out; unsigned long long* cyc;
    CK(cudaMalloc(&out, sms * WARPS * 32 * sizeof(float))); CK(cudaMalloc(&cyc, sms * sizeof(unsigned long long)));
    const size_t smem = (size_t)WARPS * STAGES * ITER_BYTES;
    auto kern = k<MODE, CBYTES, STAGES, CONSUME>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<sms, WARPS * 32, smem>>>(table, map, 64, out, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kern<<<sms, WARPS * 32, smem>>>(table, map, iters, out, cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<unsigned long long> h(sms);
    CK(cudaMemcpy(h.data(), cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double cmax = 0; for (auto c : h) cmax = c > cmax ? (double)c : cmax;
    const double warp_iters = (double)sms * WARPS * iters;
    const int ncopy = MODE == 0 ? 0 : (MODE == 1 ? ITER_BYTES / CBYTES : 8);
    printf("%-44s %8.3f ms  %6.1f clk per warp-iteration per SM (%5.1f B/clk/SM, %6.2f TB/s chip)  %5.2f clk per copy per SM\n", name, ms,
           cmax / (iters * (double)WARPS), ITER_BYTES * (double)WARPS * iters / cmax, warp_iters * ITER_BYTES / ms / 1e9,
           ncopy ? cmax / (iters * (double)WARPS * ncopy) : 0.0);
    CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main() {
    const size_t n = (size_t)NP * NT * NE * ROW;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 977) * 1e-3f;
    float* table; CK(cudaMalloc(&table, n * sizeof(float)));
    CK(cudaMemcpy(table, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    CUtensorMap map;
    {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
        const cuuint64_t dims[4] = {ROW, NE, NP, NT};                      // innermost first: g-point (x2), eta, p, T  [table is [p][T][eta][row]]
        const cuuint64_t strides[3] = {ROW * 4ull, (cuuint64_t)NT * NE * ROW * 4ull, (cuuint64_t)NE * ROW * 4ull};   // bytes of dims 1..3
        const cuuint32_t box[4] = {32, 2, 2, 1}, estr[4] = {1, 1, 1, 1};
        CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled -> %d\n", (int)r); return 1; }
    }
    const int iters = 4096;
    printf("12 warps/SM, 4 KB per warp-iteration, table %.1f MB (L2 resident)\n", n * 4 / 1e6);
    run<0, 512, 1, true>("LDG per lane (8 x .64 + 4 x .128)", table, map, iters);
    run<1, 128, 2, true>("cp.async.bulk 32 x 128 B, 2 stages + LDS", table, map, iters);
    run<1, 256, 2, true>("cp.async.bulk 16 x 256 B, 2 stages + LDS", table, map, iters);
    run<1, 512, 2, true>("cp.async.bulk  8 x 512 B, 2 stages + LDS", table, map, iters);
    run<1, 1024, 2, true>("cp.async.bulk  4 x 1 KB, 2 stages + LDS", table, map, iters);
    run<1, 2048, 2, true>("cp.async.bulk  2 x 2 KB, 2 stages + LDS", table, map, iters);
    run<1, 256, 3, true>("cp.async.bulk 16 x 256 B, 3 stages + LDS", table, map, iters);
    run<1, 512, 3, true>("cp.async.bulk  8 x 512 B, 3 stages + LDS", table, map, iters);
    run<1, 256, 2, false>("cp.async.bulk 16 x 256 B, 2 stages, no LDS", table, map, iters);
    run<1, 512, 2, false>("cp.async.bulk  8 x 512 B, 2 stages, no LDS", table, map, iters);
    run<2, 512, 2, true>("tensor TMA 8 boxes 32x2x2 (512 B), 2 st + LDS", table, map, iters);
    run<2, 512, 3, true>("tensor TMA 8 boxes 32x2x2 (512 B), 3 st + LDS", table, map, iters);
    run<2, 512, 2, false>("tensor TMA 8 boxes 32x2x2, 2 st, no LDS", table, map, iters);
    return 0;
}

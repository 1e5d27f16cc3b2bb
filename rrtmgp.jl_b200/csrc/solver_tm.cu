// Host-side launch of the Float64 tensor-memory kernels (solver_tm.cuh).
#include "solver_tm.cuh"

#include "solver_launch.cuh"

namespace rb {

static int plan_smem_tm(SolveParams<double>& P, TmSmem& F) {
    const int nlay = P.nlay, nlev = nlay + 1, maxb = P.lut.maxb;
    const int nrec = nlay < 32 ? nlay : 32;                     // generic band records, 32 layers at a time
    P.rec_words = 4 + P.lut.nminor_max + 6;
    P.rec_row = maxb * P.rec_words | 1;
    int off = 0;
    P.off_colj = off; off = align_up(off + nlay * (int)sizeof(int), 16);
    P.off_colp = off; off = align_up(off + nlay * 4 * (int)sizeof(double), 16);
    P.off_recj = off; off = align_up(off + nrec * maxb * (int)sizeof(int), 16);
    P.off_rec = off;  off = align_up(off + nrec * P.rec_row * (int)sizeof(double), 16);
    P.off_plk = off;  off = align_up(off + maxb * 2 * nlev * (int)sizeof(double), 16);
    P.off_store = off;
    F.off_alpha = off; off = align_up(off + nlay * 32 * (int)sizeof(double), 128);
    F.off_stage = off; off = align_up(off + 16 * kTmStageStride * (int)sizeof(double), 16);
    F.off_acc = off;   off = align_up(off + 3 * kTmAccStride * (int)sizeof(double), 128);
    P.warp_bytes = off;
    return kTmWarps * off;
}

template <int MODE> static int launch_tm_t(SolveParams<double>& P, int max_smem_optin, cudaStream_t stream) {
    TmSmem F;
    const size_t smem = (size_t)plan_smem_tm(P, F);
    if ((int)smem > max_smem_optin - 64) return -1;              // does not fit: generic kernel
    auto kern = solve_kernel_tm<MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (P.work_counter == nullptr) return -1;
    e = cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return (int)e;
    const int sms = sm_count_of_current_device();
    const int grid = P.ncol < sms ? P.ncol : sms;               // persistent: one CTA per SM, fewer for fewer columns
    F.active_warps = (P.ncol + grid - 1) / grid < kTmWarps ? (P.ncol + grid - 1) / grid : kTmWarps;
    kern<<<grid, kTmWarps * 32, smem, stream>>>(P, F);
    return (int)cudaGetLastError();
}

// Float64 two-stream, nlay <= 64, broadband fluxes only; returns -1 when not applicable
int launch_tm(int mode, SolveParams<double>& P, int max_smem_optin, cudaStream_t s) {
    if (P.nlay > kTmMaxLay || P.nlay < 2 || P.io.band_up != nullptr) return -1;
    // a handful of columns is a latency problem, not a throughput one: with at most two warps per SM the generic kernel's
    // three short loops finish a column sooner (128 columns: 0.79 + 0.71 ms against 0.90 + 1.23 ms here)
    if (P.ncol < 2 * sm_count_of_current_device()) return -1;
    if (mode == MODE_LW_2STREAM) return launch_tm_t<MODE_LW_2STREAM>(P, max_smem_optin, s);
    if (mode == MODE_SW_2STREAM) return launch_tm_t<MODE_SW_2STREAM>(P, max_smem_optin, s);
    return -1;
}

}  // namespace rb

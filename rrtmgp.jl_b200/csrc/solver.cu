// Host-side launch of the fused column kernels (solver.cuh).
#include "solver.cuh"

namespace rb {

static inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

template <typename FT> int plan_smem(int mode, SolveParams<FT>& P) {
    const int nlay = P.nlay, nlev = nlay + 1, maxb = P.lut.maxb;
    const int nv = mode == MODE_LW_2STREAM ? 4 : 5;
    P.rec_words = 4 + P.lut.nminor_max + 6;
    int off = 0;
    P.off_colj = off; off = align_up(off + nlay * (int)sizeof(int), 16);
    P.off_colp = off; off = align_up(off + nlay * 4 * (int)sizeof(FT), 16);
    P.off_recj = off; off = align_up(off + nlay * maxb * (int)sizeof(int), 16);
    P.off_rec = off;  off = align_up(off + nlay * maxb * P.rec_words * (int)sizeof(FT), 16);
    P.off_plk = off;  off = align_up(off + maxb * 2 * nlev * (int)sizeof(FT), 16);
    P.off_store = off; off = align_up(off + nlev * nv * 32 * (int)sizeof(FT), 128);
    P.warp_bytes = off;
    return off;
}

template <typename FT, int MODE>
static int launch_mode(SolveParams<FT>& P, int max_smem_optin, cudaStream_t stream) {
    const int wb = plan_smem<FT>(MODE, P);
    int wpc = 8;
    while (wpc > 1 && wpc * wb > max_smem_optin) wpc >>= 1;
    if (wpc * wb > max_smem_optin) return (int)cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)wpc * wb;
    auto kern = solve_kernel<FT, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = (P.ncol + wpc - 1) / wpc;
    kern<<<grid, wpc * 32, smem, stream>>>(P);
    return (int)cudaGetLastError();
}

template <typename FT> int launch_solve(int mode, SolveParams<FT>& P, int max_smem_optin, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    switch (mode) {
        case MODE_LW_2STREAM: return launch_mode<FT, MODE_LW_2STREAM>(P, max_smem_optin, s);
        case MODE_LW_NOSCAT: return launch_mode<FT, MODE_LW_NOSCAT>(P, max_smem_optin, s);
        case MODE_SW_2STREAM: return launch_mode<FT, MODE_SW_2STREAM>(P, max_smem_optin, s);
    }
    return (int)cudaErrorInvalidValue;
}

template int launch_solve<float>(int, SolveParams<float>&, int, void*);
template int launch_solve<double>(int, SolveParams<double>&, int, void*);
template int plan_smem<float>(int, SolveParams<float>&);
template int plan_smem<double>(int, SolveParams<double>&);

}  // namespace rb

// Host-side launch of the fused column kernels (solver.cuh).
#include "solver_tmem.cuh"

#include <cstdlib>
#include <cstring>
#include <type_traits>

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

// Shared-memory plan of the TMEM kernels: the level store shrinks to the albedo array (or nothing).
static int plan_smem_tmem(SolveParams<float>& P, bool alpha_tmem) {
    const int nlay = P.nlay, nlev = nlay + 1, maxb = P.lut.maxb;
    P.rec_words = 4 + P.lut.nminor_max + 6;
    int off = 0;
    P.off_colj = off; off = align_up(off + nlay * (int)sizeof(int), 16);
    P.off_colp = off; off = align_up(off + nlay * 4 * (int)sizeof(float), 16);
    P.off_recj = off; off = align_up(off + nlay * maxb * (int)sizeof(int), 16);
    P.off_rec = off;  off = align_up(off + nlay * maxb * P.rec_words * (int)sizeof(float), 16);
    P.off_plk = off;  off = align_up(off + maxb * 2 * nlev * (int)sizeof(float), 16);
    P.off_store = off; off = align_up(off + (alpha_tmem ? 0 : nlay * 32 * (int)sizeof(float)), 128);
    P.warp_bytes = off;
    return off;
}

template <int MODE, bool ALPHA_TMEM>
static int launch_tmem(SolveParams<float>& P, int max_smem_optin, cudaStream_t stream) {
    const int wb = plan_smem_tmem(P, ALPHA_TMEM);
    size_t smem = (size_t)kTmemWarpsPerCta * wb;
    // residency must match the TMEM budget (512 columns / SM): 2 CTAs x 256 or <= 4 CTAs x 128
    if (ALPHA_TMEM && smem < 80 * 1024) smem = 80 * 1024;
    if ((int)smem > max_smem_optin) return (int)cudaErrorInvalidConfiguration;
    auto kern = solve_kernel_tmem<MODE, ALPHA_TMEM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = (P.ncol + kTmemWarpsPerCta - 1) / kTmemWarpsPerCta;
    kern<<<grid, kTmemWarpsPerCta * 32, smem, stream>>>(P);
    return (int)cudaGetLastError();
}

// RRTMGP_B200_KERNEL = smem | tmem2 | tmem3 overrides the level-store placement (experiments);
// default: TMEM (A, B) + shared-memory albedo whenever the configuration allows it.
static int kernel_choice() {
    const char* e = std::getenv("RRTMGP_B200_KERNEL");
    if (!e) return 2;
    if (!std::strcmp(e, "smem")) return 0;
    if (!std::strcmp(e, "tmem3")) return 3;
    return 2;
}

template <typename FT> static int try_tmem(int, SolveParams<FT>&, int, cudaStream_t) { return -1; }
template <> int try_tmem<float>(int mode, SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    const int choice = kernel_choice();
    if (choice == 0 || P.nlay > 64 || mode == MODE_LW_NOSCAT || P.io.band_up != nullptr) return -1;
    if (mode == MODE_LW_2STREAM)
        return choice == 3 ? launch_tmem<MODE_LW_2STREAM, true>(P, max_smem_optin, s) : launch_tmem<MODE_LW_2STREAM, false>(P, max_smem_optin, s);
    return choice == 3 ? launch_tmem<MODE_SW_2STREAM, true>(P, max_smem_optin, s) : launch_tmem<MODE_SW_2STREAM, false>(P, max_smem_optin, s);
}

template <typename FT> int launch_solve(int mode, SolveParams<FT>& P, int max_smem_optin, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int t = try_tmem<FT>(mode, P, max_smem_optin, s);
    if (t >= 0) return t;
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

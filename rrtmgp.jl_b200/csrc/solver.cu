// Host-side launch of the fused column kernels (solver.cuh).
#include "solver_launch.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace rb {

template <typename FT> int plan_smem(int mode, SolveParams<FT>& P) {
    const int nlay = P.nlay, nlev = nlay + 1, maxb = P.lut.maxb;
    const int nv = mode == MODE_LW_2STREAM ? 4 : 5;
    P.rec_words = 4 + P.lut.nminor_max + 6;
    P.rec_row = maxb * P.rec_words | 1;   // odd row stride: phase 1 writes records with lane = layer (scalar accesses only here)
    int off = 0;
    P.off_colj = off; off = align_up(off + nlay * (int)sizeof(int), 16);
    P.off_colp = off; off = align_up(off + nlay * 4 * (int)sizeof(FT), 16);
    P.off_recj = off; off = align_up(off + nlay * maxb * (int)sizeof(int), 16);
    P.off_rec = off;  off = align_up(off + nlay * P.rec_row * (int)sizeof(FT), 16);
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

// groups of four minor-absorber slots per band: 1 (synthetic pack) or 2 (up to 8 / 7 + Rayleigh, real tables)

// RRTMGP_B200_KERNEL=generic forces the shared-memory kernels of solver.cuh; =ws selects the warp-specialised
// pipeline of solver_ws.cuh instead of the single-role fast kernels of solver_fast.cuh (measured slower on B200,
// profiles/r2c_*: kept as a parity-tested experiment).
static bool fast_enabled() {
    const char* e = std::getenv("RRTMGP_B200_KERNEL");
    return !(e && !std::strcmp(e, "generic"));
}
static bool ws_enabled() {
    const char* e = std::getenv("RRTMGP_B200_KERNEL");
    return e && !std::strcmp(e, "ws");
}

int launch_tm(int mode, SolveParams<double>& P, int max_smem_optin, cudaStream_t s);   // solver_tm.cu

// Fast paths.  Float64: the tensor-memory kernels of solver_tm.cuh (two-stream, nlay <= 64, any table shape).
// Float32: nlay <= 95, real-table shape.  Return -1 when not applicable.
template <typename FT> static int try_fast(int, SolveParams<FT>&, int, cudaStream_t) { return -1; }
template <> int try_fast<double>(int mode, SolveParams<double>& P, int max_smem_optin, cudaStream_t s) {
    if (!fast_enabled()) return -1;
    return launch_tm(mode, P, max_smem_optin, s);
}
template <> int try_fast<float>(int mode, SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    const GasLut<float>& L = P.lut;
    if (!fast_enabled() || P.nlay >= FastGeom<8>::max_lay || P.nlay < 2) return -1;
    if (P.io.band_up != nullptr && !L.bands_of_16) return -1;   // per-band sums = half-row sums only for aligned 16-g-point bands
    if (L.n_eta != 9 || L.n_t != 14 || L.maxb != 2 || L.n_minor_groups > 2 || (L.n_gpt % 32) != 0) return -1;
    const bool ng1 = L.n_minor_groups == 1;
    if (ws_enabled() && P.nlay <= kWsMaxLay) {   // warp-specialised pipeline (gas warps -> RT warps)
        int t = -1;
        if (mode == MODE_LW_2STREAM && L.n_gpt == 256 && L.kmaj_pf != nullptr)
            t = ng1 ? launch_ws_lw_ng1(P, max_smem_optin, s) : launch_ws_lw_ng2(P, max_smem_optin, s);
        else if (mode == MODE_SW_2STREAM && L.n_gpt == 224)
            t = ng1 ? launch_ws_sw_ng1(P, max_smem_optin, s) : launch_ws_sw_ng2(P, max_smem_optin, s);
        if (t >= 0) return t;
    }
    if (mode == MODE_LW_2STREAM && L.n_gpt == 256 && L.kmaj_pf != nullptr)
        return ng1 ? launch_fast_lw_ng1(P, max_smem_optin, s) : launch_fast_lw_ng2(P, max_smem_optin, s);
    if (mode == MODE_LW_NOSCAT && L.n_gpt == 256 && L.kmaj_pf != nullptr) {
        if (P.n_mu == 1) return ng1 ? launch_fast_noscat1_ng1(P, max_smem_optin, s) : launch_fast_noscat1_ng2(P, max_smem_optin, s);
        return ng1 ? launch_fast_noscat4_ng1(P, max_smem_optin, s) : launch_fast_noscat4_ng2(P, max_smem_optin, s);
    }
    if (mode == MODE_SW_2STREAM && L.n_gpt == 224)
        return ng1 ? launch_fast_sw_ng1(P, max_smem_optin, s) : launch_fast_sw_ng2(P, max_smem_optin, s);
    return -1;
}

template <typename FT> int launch_solve(int mode, SolveParams<FT>& P, int max_smem_optin, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int t = try_fast<FT>(mode, P, max_smem_optin, s);
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

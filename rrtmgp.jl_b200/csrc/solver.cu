// Host-side launch of the fused column kernels (solver.cuh).
#include "solver_fast.cuh"

#include <algorithm>
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

// Shared-memory plan of the fast kernels: band records + high-level albedos + staging tile + accumulators.
template <int WARPS>
static int plan_smem_fast(SolveParams<float>& P, FastSmem& F, int max_smem_optin) {
    using Geom = FastGeom<WARPS>;
    constexpr int kAlphaTmemLevels = Geom::alpha_tmem_levels, kAccStride = Geom::acc_stride, kFastWarps = WARPS;
    const int nlay = P.nlay, nlev = nlay + 1, maxb = 2;
    const int nrec = nlay < 32 ? nlay : 32;                    // band records cover half a column at a time
    // record = 8 corner weights, {s1, s2, two major-table offsets}, 4 slot scalings per group,
    // {aerosol-only products, minor-table offset}, {cloud+aerosol products, minor-table offset}
    P.rec_words = 20 + 4 * P.lut.n_minor_groups;
    int off = 0;
    P.off_colj = off; off = align_up(off + nlay * (int)sizeof(int), 16);
    P.off_colp = off; off = align_up(off + nlay * 4 * (int)sizeof(float), 16);
    P.off_recj = off;   // unused by the fast kernels (eta offsets travel in the record)
    P.off_rec = off;  off = align_up(off + nrec * maxb * P.rec_words * (int)sizeof(float), 16);
    P.off_plk = off;  off = align_up(off + maxb * (nlev + 1) * (int)sizeof(float), 16);   // B(t_lev), B(t_sfc) per band
    P.off_store = off;
    const int n_hi = (nlay > kAlphaTmemLevels ? nlay - kAlphaTmemLevels : 0) + 1;   // + dummy slot
    F.off_alpha = off; off = align_up(off + n_hi * 32 * (int)sizeof(float), 128);
    F.off_stage = off; off = align_up(off + 16 * kStageStride * (int)sizeof(float), 16);
    F.off_acc = off;   off = align_up(off + 3 * kAccStride * (int)sizeof(float), 128);
    F.off_bacc = -1;
    if (P.io.band_up != nullptr) { F.off_bacc = off; off = align_up(off + 4 * kAccStride * (int)sizeof(float), 128); }
    P.warp_bytes = off;
    // CTA-shared tail: the staged small-table block and the global-mean vmr array
    int tail = kFastWarps * off;
    F.off_vmr = tail;  tail = align_up(tail + (P.ngas > 0 ? P.ngas : 1) * (int)sizeof(float), 128);
    F.off_blob = tail;
    int room = max_smem_optin - 64 - tail;   // 64: static shared memory of the kernel
    // RRTMGP_B200_STAGE_BYTES caps the staged prefix (A/B experiments: shared memory is taken from the L1 cache)
    if (const char* e = std::getenv("RRTMGP_B200_STAGE_BYTES")) room = std::min(room, std::atoi(e));
    F.staged_bytes = 0;
    for (int i = 0; i < P.lut.n_blob_cut; ++i)
        if (P.lut.blob_cut[i] <= room) F.staged_bytes = P.lut.blob_cut[i];
    return tail + F.staged_bytes;
}

static int sm_count_of_current_device() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    return cached[dev] > 0 ? cached[dev] : 148;
}

template <int MODE, int NGPT, int NG, bool HAS_CLD, bool HAS_AER, bool SPECTRAL, int WARPS>
static int launch_fast_t(SolveParams<float>& P, int max_smem_optin, cudaStream_t stream) {
    FastSmem F;
    constexpr int kFastWarps = WARPS;
    const size_t smem = (size_t)plan_smem_fast<WARPS>(P, F, max_smem_optin);
    if ((int)smem > max_smem_optin) return -1;   // does not fit: generic kernel
    auto kern = solve_kernel_fast<MODE, NGPT, NG, HAS_CLD, HAS_AER, SPECTRAL, WARPS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (P.work_counter == nullptr) return -1;
    e = cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return (int)e;
    const int need = (P.ncol + kFastWarps - 1) / kFastWarps;
    const int sms = sm_count_of_current_device();
    const int grid = need < sms ? need : sms;   // persistent: one CTA per SM
    kern<<<grid, kFastWarps * 32, smem, stream>>>(P, F);
    return (int)cudaGetLastError();
}

template <int MODE, int NGPT, int NG, bool SPECTRAL, int WARPS>
static int launch_fast_sp(SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    const bool c = P.use_cloud != 0, a = P.use_aero != 0;
    if (c && a) return launch_fast_t<MODE, NGPT, NG, true, true, SPECTRAL, WARPS>(P, max_smem_optin, s);
    if (c) return launch_fast_t<MODE, NGPT, NG, true, false, SPECTRAL, WARPS>(P, max_smem_optin, s);
    if (a) return launch_fast_t<MODE, NGPT, NG, false, true, SPECTRAL, WARPS>(P, max_smem_optin, s);
    return launch_fast_t<MODE, NGPT, NG, false, false, SPECTRAL, WARPS>(P, max_smem_optin, s);
}

// nlay <= 64: 12 warps per SM; taller columns (<= 95 layers): the 8-warp geometry, broadband fluxes only
template <int MODE, int NGPT, int NG>
static int launch_fast_ng(SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    if (P.nlay > FastGeom<12>::max_lay)
        return P.io.band_up != nullptr ? -1 : launch_fast_sp<MODE, NGPT, NG, false, 8>(P, max_smem_optin, s);
    return P.io.band_up != nullptr ? launch_fast_sp<MODE, NGPT, NG, true, 12>(P, max_smem_optin, s)
                                   : launch_fast_sp<MODE, NGPT, NG, false, 12>(P, max_smem_optin, s);
}

// groups of four minor-absorber slots per band: 1 (synthetic pack) or 2 (up to 8 / 7 + Rayleigh, real tables)
template <int MODE, int NGPT>
static int launch_fast_flags(SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    return P.lut.n_minor_groups == 1 ? launch_fast_ng<MODE, NGPT, 1>(P, max_smem_optin, s)
                                     : launch_fast_ng<MODE, NGPT, 2>(P, max_smem_optin, s);
}

// RRTMGP_B200_KERNEL=generic forces the shared-memory kernels of solver.cuh (experiments / A-B tests).
static bool fast_enabled() {
    const char* e = std::getenv("RRTMGP_B200_KERNEL");
    return !(e && !std::strcmp(e, "generic"));
}

// Fast path: Float32, two-stream, nlay <= 64, real-table shape. Returns -1 when not applicable.
template <typename FT> static int try_fast(int, SolveParams<FT>&, int, cudaStream_t) { return -1; }
template <> int try_fast<float>(int mode, SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    const GasLut<float>& L = P.lut;
    if (!fast_enabled() || P.nlay >= FastGeom<8>::max_lay || P.nlay < 2 || mode == MODE_LW_NOSCAT) return -1;
    if (P.io.band_up != nullptr && !L.bands_of_16) return -1;   // per-band sums = half-row sums only for aligned 16-g-point bands
    if (L.n_eta != 9 || L.n_t != 14 || L.maxb != 2 || L.n_minor_groups > 2 || (L.n_gpt % 32) != 0) return -1;
    if (mode == MODE_LW_2STREAM && L.n_gpt == 256 && L.kmaj_pf != nullptr) return launch_fast_flags<MODE_LW_2STREAM, 256>(P, max_smem_optin, s);
    if (mode == MODE_SW_2STREAM && L.n_gpt == 224) return launch_fast_flags<MODE_SW_2STREAM, 224>(P, max_smem_optin, s);
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

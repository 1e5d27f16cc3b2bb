// Launch plumbing of the fast kernels (solver_fast.cuh).  The kernels are instantiated in several translation
// units (solver_fast_inst.cu compiled once per (mode, minor-slot groups), see the Makefile) so the build
// parallelises; solver.cu dispatches to their entry points.
#pragma once
#include "solver_fast.cuh"
#include "solver_ws.cuh"

#include <algorithm>
#include <cstdlib>

namespace rb {

static inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

// entry points of the per-(mode, NG) translation units; return a cudaError_t as int, or -1 = not applicable
int launch_fast_lw_ng1(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);
int launch_fast_lw_ng2(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);
int launch_fast_sw_ng1(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);
int launch_fast_sw_ng2(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);
int launch_fast_noscat1_ng1(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);   // one Gauss angle
int launch_fast_noscat1_ng2(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);
int launch_fast_noscat4_ng1(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);   // two to four angles
int launch_fast_noscat4_ng2(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);
// warp-specialised two-stream kernels (solver_ws.cuh), nlay <= 64
int launch_ws_lw_ng1(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);
int launch_ws_lw_ng2(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);
int launch_ws_sw_ng1(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);
int launch_ws_sw_ng2(SolveParams<float>& P, int max_smem_optin, cudaStream_t s);

// Shared-memory plan of the fast kernels: band records + high-level albedos + staging tile + accumulators.
template <int WARPS>
static int plan_smem_fast(SolveParams<float>& P, FastSmem& F, int max_smem_optin, bool noscat) {
    using Geom = FastGeom<WARPS>;
    constexpr int kAlphaTmemLevels = Geom::alpha_tmem_levels, kAccStride = Geom::acc_stride, kFastWarps = WARPS;
    const int nlay = P.nlay, nlev = nlay + 1, maxb = 2;
    const int nrec = nlay < 32 ? nlay : 32;                    // band records cover half a column at a time
    // record = 8 corner weights, {s1, s2, two major-table offsets}, 4 slot scalings per group,
    // {aerosol-only products, minor-table offset}, {cloud+aerosol products, minor-table offset}
    P.rec_words = 20 + 4 * P.lut.n_minor_groups;
    // Phase 1 writes a record with lane = layer, so the lane stride is the row: 2 * rec_words = 48 / 56 words puts
    // every lane on bank 0 or 16 (16-way conflicts on each of the ~50 stores per lane; ncu: 8.4e6 excessive shared
    // wavefronts per SM, 28 % of the L1 data-pipe time).  A row of 4 * odd words keeps the 16-byte alignment of the
    // 128-bit record loads and spreads the lanes over 8 bank groups.
    P.rec_row = maxb * P.rec_words;
    if (((P.rec_row >> 2) & 1) == 0) P.rec_row += 4;
    int off = 0;
    P.off_colj = off; off = align_up(off + nlay * (int)sizeof(int), 16);
    P.off_colp = off; off = align_up(off + nlay * 4 * (int)sizeof(float), 16);
    P.off_recj = off;   // unused by the fast kernels (eta offsets travel in the record)
    P.off_rec = off;  off = align_up(off + nrec * P.rec_row * (int)sizeof(float), 16);
    // per band: B(t_lev), B(t_sfc) (+ B(t_lay), no-scattering)
    P.off_plk = off;  off = align_up(off + maxb * (noscat ? 2 * nlev : nlev + 1) * (int)sizeof(float), 16);
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
    F.off_vmr = tail;  tail = align_up(tail + ((P.ngas > 0 ? P.ngas : 1) + 1) * (int)sizeof(float), 128);
    F.off_blob = tail;
    int room = max_smem_optin - 256 - tail;   // 256: static shared memory of the kernel (barrier, pointer table)
    // Shared memory is carved out of the L1 cache, which serves the k-distribution gathers: staging the most-read
    // ~20 KB of the block (key species, reference vmr, minor-absorber lists, the Planck table; not the aerosol and cloud tables) and
    // leaving the rest of the space to L1 measured best on B200 at ncol = 1e5 (LW 17.9 ms against 18.4 with
    // everything staged and 18.7 with nothing; SW 16.8 / 17.0 / 17.1).  RRTMGP_B200_STAGE_BYTES overrides the cap.
    room = std::min(room, 20480);
    if (const char* e = std::getenv("RRTMGP_B200_STAGE_BYTES")) room = std::min(max_smem_optin - 256 - tail, std::atoi(e));
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

template <int MODE, int NGPT, int NG, bool HAS_CLD, bool HAS_AER, bool SPECTRAL, int WARPS, int NMU>
static int launch_fast_t(SolveParams<float>& P, int max_smem_optin, cudaStream_t stream) {
    FastSmem F;
    constexpr int kFastWarps = WARPS;
    const size_t smem = (size_t)plan_smem_fast<WARPS>(P, F, max_smem_optin, MODE == MODE_LW_NOSCAT);
    if ((int)smem > max_smem_optin - 256) return -1;   // does not fit (256: static shared memory): generic kernel
    auto kern = solve_kernel_fast<MODE, NGPT, NG, HAS_CLD, HAS_AER, SPECTRAL, WARPS, NMU>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (P.work_counter == nullptr) return -1;
    e = cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return (int)e;
    const int sms = sm_count_of_current_device();
    // small shards: P.split (1, 2 or 4 work items per column) comes from the handle (api.cu: split_of)
    if (P.split_scratch == nullptr || P.split_flags == nullptr || (P.split != 2 && P.split != 4)) P.split = 1;
    const long long items = (long long)P.ncol * P.split;
    const int grid = items < sms ? (int)items : sms;   // persistent: one CTA per SM; fewer CTAs only for fewer items than SMs
    const long long per_cta = (items + grid - 1) / grid;
    F.active_warps = per_cta < kFastWarps ? (int)per_cta : kFastWarps;
    kern<<<grid, kFastWarps * 32, smem, stream>>>(P, F);
    return (int)cudaGetLastError();
}

template <int MODE, int NGPT, int NG, bool SPECTRAL, int WARPS, int NMU = 1>
static int launch_fast_sp(SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    const bool c = P.use_cloud != 0, a = P.use_aero != 0;
    if (c && a) return launch_fast_t<MODE, NGPT, NG, true, true, SPECTRAL, WARPS, NMU>(P, max_smem_optin, s);
    if (c) return launch_fast_t<MODE, NGPT, NG, true, false, SPECTRAL, WARPS, NMU>(P, max_smem_optin, s);
    if (a) return launch_fast_t<MODE, NGPT, NG, false, true, SPECTRAL, WARPS, NMU>(P, max_smem_optin, s);
    return launch_fast_t<MODE, NGPT, NG, false, false, SPECTRAL, WARPS, NMU>(P, max_smem_optin, s);
}

// nlay <= 64: 12 warps per SM; taller columns (<= 95 layers): the 8-warp geometry
// NMU: Gauss angles of the no-scattering LW kernel (1, or 4 = up to four with P.n_mu active); 1 for the two-stream modes
template <int MODE, int NGPT, int NG, int NMU = 1>
static int launch_fast_ng(SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    if constexpr (MODE == MODE_LW_NOSCAT) {   // NoScatLWRTE carries no band fluxes (src/rte/RTE.jl:53-72)
        if (P.io.band_up != nullptr) return -1;
        return P.nlay > FastGeom<12>::max_lay ? launch_fast_sp<MODE, NGPT, NG, false, 8, NMU>(P, max_smem_optin, s)
                                              : launch_fast_sp<MODE, NGPT, NG, false, 12, NMU>(P, max_smem_optin, s);
    } else {
        if (P.nlay > FastGeom<12>::max_lay)
            return P.io.band_up != nullptr ? launch_fast_sp<MODE, NGPT, NG, true, 8, NMU>(P, max_smem_optin, s)
                                           : launch_fast_sp<MODE, NGPT, NG, false, 8, NMU>(P, max_smem_optin, s);
        return P.io.band_up != nullptr ? launch_fast_sp<MODE, NGPT, NG, true, 12, NMU>(P, max_smem_optin, s)
                                       : launch_fast_sp<MODE, NGPT, NG, false, 12, NMU>(P, max_smem_optin, s);
    }
}

// ---- warp-specialised kernels: per (gas warp, RT warp) pair the gas warp's Warp<> layout, the hand-off ring, the
// RT warp's staging tile and accumulators; CTA-shared tail as in plan_smem_fast ----
static int plan_smem_ws(SolveParams<float>& P, WsSmem& F, int max_smem_optin) {
    const int nlay = P.nlay, nlev = nlay + 1, maxb = 2;
    const int nrec = nlay < 32 ? nlay : 32;
    P.rec_words = 20 + 4 * P.lut.n_minor_groups;
    P.rec_row = maxb * P.rec_words;
    if (((P.rec_row >> 2) & 1) == 0) P.rec_row += 4;          // row of 4 * odd words (see plan_smem_fast)
    int off = 0;
    P.off_colj = off; off = align_up(off + nlay * (int)sizeof(int), 16);
    P.off_colp = off; off = align_up(off + nlay * 4 * (int)sizeof(float), 16);
    P.off_recj = off;
    P.off_rec = off;  off = align_up(off + nrec * P.rec_row * (int)sizeof(float), 16);
    P.off_plk = off;  off = align_up(off + maxb * (nlev + 1) * (int)sizeof(float), 16);
    P.off_store = off;
    F.off_ring = off;  off = align_up(off + kWsStages * kWsStageF4 * 16, 128);
    F.off_stage = off; off = align_up(off + 16 * kStageStride * (int)sizeof(float), 16);
    F.off_acc = off;   off = align_up(off + 3 * kWsAccStride * (int)sizeof(float), 128);
    F.off_bacc = -1;
    if (P.io.band_up != nullptr) { F.off_bacc = off; off = align_up(off + 4 * kWsAccStride * (int)sizeof(float), 128); }
    P.warp_bytes = off;                                        // bytes per pair
    int tail = kWsPairs * off;
    F.off_vmr = tail;  tail = align_up(tail + ((P.ngas > 0 ? P.ngas : 1) + 1) * (int)sizeof(float), 128);
    F.off_blob = tail;
    int room = max_smem_optin - 1024 - tail;                   // 1024: static shared memory (mbarriers, mailboxes)
    if (const char* e = std::getenv("RRTMGP_B200_STAGE_BYTES")) room = std::min(room, std::atoi(e));
    F.staged_bytes = 0;
    for (int i = 0; i < P.lut.n_blob_cut; ++i)
        if (P.lut.blob_cut[i] <= room) F.staged_bytes = P.lut.blob_cut[i];
    return tail + F.staged_bytes;
}

template <int MODE, int NGPT, int NG, bool HAS_CLD, bool HAS_AER, bool SPECTRAL>
static int launch_ws_t(SolveParams<float>& P, int max_smem_optin, cudaStream_t stream) {
    WsSmem F;
    const size_t smem = (size_t)plan_smem_ws(P, F, max_smem_optin);
    if ((int)smem > max_smem_optin - 1024) return -1;
    auto kern = solve_kernel_ws<MODE, NGPT, NG, HAS_CLD, HAS_AER, SPECTRAL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (P.work_counter == nullptr) return -1;
    e = cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return (int)e;
    const int need = (P.ncol + kWsPairs - 1) / kWsPairs;
    const int sms = sm_count_of_current_device();
    const int grid = need < sms ? need : sms;   // persistent: one CTA per SM
    kern<<<grid, kWsPairs * 64, smem, stream>>>(P, F);
    return (int)cudaGetLastError();
}

template <int MODE, int NGPT, int NG>
static int launch_ws_ng(SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    if (P.nlay > kWsMaxLay) return -1;
    const bool c = P.use_cloud != 0, a = P.use_aero != 0, sp = P.io.band_up != nullptr;
#define RB_WS(C, A, S) return launch_ws_t<MODE, NGPT, NG, C, A, S>(P, max_smem_optin, s)
    if (sp) {
        if (c && a) RB_WS(true, true, true);
        if (c) RB_WS(true, false, true);
        if (a) RB_WS(false, true, true);
        RB_WS(false, false, true);
    }
    if (c && a) RB_WS(true, true, false);
    if (c) RB_WS(true, false, false);
    if (a) RB_WS(false, true, false);
    RB_WS(false, false, false);
#undef RB_WS
}

}  // namespace rb

// C ABI (include/rrtmgp_b200.h): handle lifetime, argument validation, and the Layer-2
// orchestration of src/api/update_fluxes.jl on a CUDA stream.
#include <cuda_runtime.h>

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/rrtmgp_b200.h"
#include "lut.cuh"
#include "solver.cuh"

using namespace rb;

// ---- multi-GPU state (rrtmgp_b200_comm_*): NCCL through dlopen (types restated from nccl.h 2.27: the library has no
// link-time dependency on it), the gathered arrays and every peer's copy of them (CUDA IPC) ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclChar = 0, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kMaxRanks = 16, kGatheredViews = 8;
constexpr int kSplitMaxCols = 4 * 12 * 160;    // from four columns per warp (12 warps on up to 160 SMs) nothing is split
constexpr int kCopyStreams = 4;
struct CommState {
    cudaStream_t push_stream(int j) const { return j == 0 ? copy_stream : copy_extra[j - 1]; }
    NcclApi nccl;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 0;
    cudaStream_t copy_stream = nullptr;      // bootstrap collectives and pushes to every kCopyStreams-th peer
    cudaStream_t copy_extra[kCopyStreams - 1] = {};   // the other peers' pushes: several copy engines / NVLink ports at once
    cudaEvent_t ev_kernel[6] = {}, ev_copied[kCopyStreams] = {};
    unsigned char* gathered = nullptr;       // this rank's 8 arrays, [view][nranks * ncol][nlev]
    size_t view_bytes = 0;                   // bytes of one gathered view
    unsigned char* peer[kMaxRanks] = {};     // every rank's `gathered` as mapped here (peer[rank] == gathered)
    float* token = nullptr;                  // 2 floats for the framing all-reduces
    bool ready = false;                      // comm_init completed on this rank (and, by its collective steps, on every rank)
};

struct rrtmgp_b200_handle {
    CommState* comm = nullptr;
    rrtmgp_b200_config_t cfg;
    rrtmgp_b200_buffers_t buf;
    bool bound = false;
    LutStore luts;
    int max_smem_optin = 0;
    int sm_count = 0;
    unsigned long long call_counter = 0;
    int last_launches = 0;
    char cuda_err[256] = {0};
    // work-queue counters of the persistent kernels: one per launch, taken round-robin (launches of one handle
    // may overlap on several streams, rrtmgp_b200_update_fluxes_range)
    unsigned int* work_counters = nullptr;
    mutable unsigned work_seq = 0;
    // small shards (solver_fast.cuh): partial sums and arrival counters of the split columns, indexed by the handle's
    // column; allocated at create for handles of fewer than kSplitMaxCols Float32 columns
    float* split_scratch = nullptr;
    unsigned int* split_flags = nullptr;
    // interpolate_levels! configuration (rrtmgp_b200_set_level_interpolation)
    int interpolation = RRTMGP_B200_NO_INTERPOLATION, bottom_extrapolation = RRTMGP_B200_SAME_AS_INTERPOLATION;
    const void* center_z = nullptr;
    const void* face_z = nullptr;
    double cp_d = 0, r_d = 0;
};

namespace {

// ------------------------------------------------------------------------------------
// small elementwise kernels of prepare_atmosphere! and update_net_fluxes!
// ------------------------------------------------------------------------------------
// add_isothermal_boundary_layer! (grid_adaptation.jl:135-150,176-214); one thread per column
template <typename FT>
__global__ void boundary_layer_kernel(rrtmgp_b200_buffers_t B, long long col0, int ncol, int nlay, int ngas, int vmr_kind, FT p_min) {
    long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    col += col0;
    const int nlev = nlay + 1, top = nlay - 1;
    FT* ld = (FT*)B.layerdata + (size_t)col * nlay * 4;
    FT* p_lev = (FT*)B.p_lev + (size_t)col * nlev;
    FT* t_lev = (FT*)B.t_lev + (size_t)col * nlev;
    ld[4 * top + 1] = (p_lev[nlev - 2] + p_min) / FT(2);
    p_lev[nlev - 1] = p_min;
    ld[4 * top + 2] = t_lev[nlev - 2];
    t_lev[nlev - 1] = t_lev[nlev - 2];
    ld[4 * top + 3] = ld[4 * (top - 1) + 3];
    size_t k = (size_t)col * nlay + top;
    if (vmr_kind == RRTMGP_B200_VMR_GM) {
        ((FT*)B.vmr_h2o)[k] = ((FT*)B.vmr_h2o)[k - 1];
        ((FT*)B.vmr_o3)[k] = ((FT*)B.vmr_o3)[k - 1];
    } else {
        FT* v = (FT*)B.vmr;
        for (int g = 0; g < ngas; ++g) v[k * ngas + g] = v[(k - 1) * ngas + g];
    }
    if (B.cld_frac) {
        FT* arrs[5] = {(FT*)B.cld_r_eff_liq, (FT*)B.cld_r_eff_ice, (FT*)B.cld_path_liq, (FT*)B.cld_path_ice, (FT*)B.cld_frac};
        for (int a = 0; a < 5; ++a) arrs[a][k] = arrs[a][k - 1];
    }
    if (B.aero_mass) {
        FT* am = (FT*)B.aero_mass; FT* as = (FT*)B.aero_size;
        for (int i = 0; i < 15; ++i) { am[k * 15 + i] = am[(k - 1) * 15 + i]; as[k * 15 + i] = as[(k - 1) * 15 + i]; }
    }
}

// interpolate_levels! (grid_adaptation.jl:87-113) with interp! / extrap! (interpolation.jl:176-252); thread per
// (column, domain face).  `mode` of a face: 1..5 interpolation scheme, 6 UseSurfaceTempAtBottom, 7 HydrostaticBottom.
template <typename FT> __device__ __forceinline__ FT lev_pow(FT a, FT b);
template <> __device__ __forceinline__ float lev_pow<float>(float a, float b) { return powf(a, b); }
template <> __device__ __forceinline__ double lev_pow<double>(double a, double b) { return pow(a, b); }
template <typename FT> __device__ __forceinline__ FT lev_sqrt(FT a);
template <> __device__ __forceinline__ float lev_sqrt<float>(float a) { return sqrtf(a); }
template <> __device__ __forceinline__ double lev_sqrt<double>(double a) { return sqrt(a); }

template <typename FT>
__device__ __forceinline__ FT power_law_p(FT T, FT p1, FT T1, FT p2, FT T2) {   // interpolation.jl:152-153,162-164
    return p1 * lev_pow(p2 / p1, rlog(T / T1) / rlog(T2 / T1));
}

template <typename FT>
__global__ void interpolate_levels_kernel(rrtmgp_b200_buffers_t B, long long col0, int ncol, int nlay_total, int nlay,
                                          int interpolation, int bottom, const FT* center_z, const FT* face_z, FT grav,
                                          FT cp_d, FT r_d) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)ncol * (nlay + 1)) return;
    const long long lcol = idx / (nlay + 1);
    const int f = (int)(idx - lcol * (nlay + 1));            // face 0 = bottom, nlay = top of the domain
    const long long col = col0 + lcol;
    const FT* ld = (const FT*)B.layerdata + (size_t)col * nlay_total * 4;
    const FT* zc = center_z ? center_z + (size_t)col * nlay_total : nullptr;
    const FT* zf = face_z ? face_z + (size_t)col * (nlay_total + 1) : nullptr;
    // the two layers the face is computed from: interior faces (below, above); boundary faces (nearest, next)
    const int l1 = f == 0 ? 0 : (f == nlay ? nlay - 1 : f - 1);
    const int l2 = f == 0 ? 1 : (f == nlay ? nlay - 2 : f);
    const FT p1 = ld[4 * l1 + 1], T1 = ld[4 * l1 + 2], p2 = ld[4 * l2 + 1], T2 = ld[4 * l2 + 2];
    const FT z = zf ? zf[f] : FT(0), z1 = zc ? zc[l1] : FT(0), z2 = zc ? zc[l2] : FT(0);
    const bool boundary = f == 0 || f == nlay;
    const int mode = (f == 0 && bottom != RRTMGP_B200_SAME_AS_INTERPOLATION) ? 5 + bottom : interpolation;
    FT p, T;
    if (mode == RRTMGP_B200_ARITHMETIC_MEAN) {
        T = boundary ? (FT(3) * T1 - T2) / FT(2) : (T1 + T2) / FT(2);
        p = boundary ? (FT(3) * p1 - p2) / FT(2) : (p1 + p2) / FT(2);
    } else if (mode == RRTMGP_B200_GEOMETRIC_MEAN) {
        T = boundary ? lev_sqrt(T1 * T1 * T1 / T2) : lev_sqrt(T1 * T2);
        p = boundary ? lev_sqrt(p1 * p1 * p1 / p2) : lev_sqrt(p1 * p2);
    } else if (mode == RRTMGP_B200_UNIFORM_Z) {
        T = boundary ? (FT(3) * T1 - T2) / FT(2) : (T1 + T2) / FT(2);
        p = T1 == T2 ? lev_sqrt(p1 * p2) : power_law_p(T, p1, T1, p2, T2);
    } else if (mode == RRTMGP_B200_UNIFORM_P) {
        p = boundary ? (FT(3) * p1 - p2) / FT(2) : (p1 + p2) / FT(2);
        T = T1 * lev_pow(T2 / T1, rlog(p / p1) / rlog(p2 / p1));   // assumes p1 != p2 (interpolation.jl:188,222)
    } else if (mode == RRTMGP_B200_BEST_FIT) {
        T = T1 + (T2 - T1) * (z - z1) / (z2 - z1);
        p = T1 == T2 ? p1 * lev_pow(p2 / p1, (z - z1) / (z2 - z1)) : power_law_p(T, p1, T1, p2, T2);
    } else {   // bottom face only: dry isentrope through the first layer (interpolation.jl:227-252)
        T = mode == 5 + RRTMGP_B200_USE_SURFACE_TEMP_AT_BOTTOM ? ((const FT*)B.t_sfc)[col] : T1 + grav / cp_d * (z1 - z);
        p = p1 * lev_pow(T / T1, cp_d / r_d);
    }
    ((FT*)B.p_lev)[(size_t)col * (nlay_total + 1) + f] = p;
    ((FT*)B.t_lev)[(size_t)col * (nlay_total + 1) + f] = T;
}

// heating_rate (standalone.jl:106-124, GrayAtmosphere.jl:152-167): (g / cp) dF_net / dp per domain layer
template <typename FT>
__global__ void heating_rate_kernel(const FT* flux_net, const FT* p_lev, FT* hr, int ncol, int nlay_total, int nlay, FT grav, FT cp_d) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)ncol * nlay) return;
    const long long col = idx / nlay;
    const int l = (int)(idx - col * nlay);
    const size_t k = (size_t)col * (nlay_total + 1) + l;
    hr[(size_t)col * nlay + l] = grav * (flux_net[k + 1] - flux_net[k]) / (p_lev[k + 1] - p_lev[k]) / cp_d;
}

// clip! (grid_adaptation.jl:232-258) + compute_col_gas_kernel! (gas_optics.jl:16-41); thread per (col, level)
template <typename FT>
__global__ void prepare_kernel(rrtmgp_b200_buffers_t B, long long col0, int ncol, int nlay, int ngas, int vmr_kind, int idx_h2o, FT p_min,
                               FT t_min, FT t_max, FT grav, FT m_dry, FT m_h2o, FT avogad, int do_clip, int do_col) {
    const int nlev = nlay + 1;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)ncol * nlev) return;
    const long long lcol = idx / nlev;
    const int lev = (int)(idx - lcol * nlev);
    const long long col = col0 + lcol;
    FT* p_lev = (FT*)B.p_lev + (size_t)col * nlev;
    FT* t_lev = (FT*)B.t_lev + (size_t)col * nlev;
    // (a thread reads its own level and the one above; with do_clip both are clipped here exactly as the thread that
    // owns the upper level will store it, so the column amount sees clipped pressures without a second pass)
    const FT p_here = do_clip ? rmax(p_lev[lev], p_min) : p_lev[lev];
    FT p_above = FT(0);
    if (lev < nlay) p_above = do_clip ? rmax(p_lev[lev + 1], p_min) : p_lev[lev + 1];
    if (do_clip) {
        p_lev[lev] = p_here;
        t_lev[lev] = rmin(rmax(t_lev[lev], t_min), t_max);
    }
    if (lev < nlay) {
        size_t k = (size_t)col * nlay + lev;
        FT* ld = (FT*)B.layerdata + k * 4;
        FT* h2o = vmr_kind == RRTMGP_B200_VMR_GM ? (FT*)B.vmr_h2o + k : (FT*)B.vmr + k * ngas + (idx_h2o - 1);
        FT v = *h2o;
        if (do_clip) {
            v = rmax(v, FT(0));
            *h2o = v;
            ld[1] = rmax(ld[1], p_min);
            ld[2] = rmin(rmax(ld[2], t_min), t_max);
        }
        if (!do_col) return;
        const FT helmert2 = FT(0.02586), m2_to_cm2 = FT(100 * 100);
        FT g0 = grav;
        if (B.lat) g0 = grav - helmert2 * rcos(FT(2) * Num<FT>::pi() * ((const FT*)B.lat)[col] / FT(180));
        FT dp = p_here - p_above;
        FT m_air = m_dry + m_h2o * v;
        ld[0] = dp * avogad / (m2_to_cm2 * m_air * g0);
    }
}

// compute_relative_humidity_kernel! (gas_optics.jl:58-80)
template <typename FT>
__global__ void rel_hum_kernel(rrtmgp_b200_buffers_t B, int ncol, int nlay, int ngas, int vmr_kind, int idx_h2o, FT mwd) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)ncol * nlay) return;
    FT* ld = (FT*)B.layerdata + (size_t)k * 4;
    FT vmr = vmr_kind == RRTMGP_B200_VMR_GM ? ((const FT*)B.vmr_h2o)[k] : ((const FT*)B.vmr)[(size_t)k * ngas + (idx_h2o - 1)];
    FT mmr = vmr * mwd;
    FT q = mmr / (FT(1) + mmr);
    FT q_tmp = rmax(FT(1e-7), q);
    FT t = ld[2];
    FT es = rexp((FT(17.67) * (t - FT(273.16))) / (t - FT(29.65)));
    ld[3] = rmax(FT(0.01) * (FT(0.263) * ld[1] * q_tmp) / es, FT(0));
}

// validate_inputs (validation.jl:28-37, 56-74): mode 0 = positive and finite, 1 = non-negative and finite,
// 2 = in [0, 1], 3 = in [-1, 1]; element i of the array is x[(i / inner) * stride + offset + i % inner]
template <typename FT>
__global__ void validate_kernel(const FT* x, long long n, int inner, long long stride, long long offset, int mode, unsigned bit, unsigned* mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < n) {
        const FT v = x[(i / inner) * stride + offset + i % inner];
        const bool finite = v == v && v - v == FT(0);
        bool ok;
        switch (mode) {
            case 0: ok = finite && v > FT(0); break;
            case 1: ok = finite && v >= FT(0); break;
            case 2: ok = finite && v >= FT(0) && v <= FT(1); break;
            default: ok = finite && v >= FT(-1) && v <= FT(1); break;
        }
        bad = !ok;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(mask, bit);
}

// transpose_sum_into! without the transpose (Fluxes.jl:423-435)
template <typename FT> __global__ void add_kernel(const FT* a, const FT* b, FT* out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

int fail_cuda(rrtmgp_b200_handle* h, cudaError_t e) {
    std::snprintf(h->cuda_err, sizeof(h->cuda_err), "%s", cudaGetErrorString(e));
    return RRTMGP_B200_ERR_CUDA;
}

template <typename FT> const Luts<FT>& luts_of(const rrtmgp_b200_handle* h);
template <> const Luts<float>& luts_of<float>(const rrtmgp_b200_handle* h) { return h->luts.f32; }
template <> const Luts<double>& luts_of<double>(const rrtmgp_b200_handle* h) { return h->luts.f64; }

// Typed view of the bound buffers, advanced to column `c0` (all arrays are [ncol][...] except
// inc_flux_lw [n_gpt][ncol] and the band fluxes [n_bnd][ncol][nlev], which keep the full column stride).
template <typename FT>
void fill_io(const rrtmgp_b200_handle* h, ColumnIO<FT>& io, long long c0) {
    const rrtmgp_b200_buffers_t& B = h->buf;
    const rrtmgp_b200_config_t& c = h->cfg;
    const long long nlay = c.nlay, nlev = c.nlay + 1;
    auto adv = [&](void* p, long long per_col) -> FT* { return p ? (FT*)p + c0 * per_col : nullptr; };
    std::memset(&io, 0, sizeof(io));
    io.layerdata = adv(B.layerdata, nlay * 4); io.p_lev = adv(B.p_lev, nlev); io.t_lev = adv(B.t_lev, nlev);
    io.t_sfc = adv(B.t_sfc, 1); io.vmr_h2o = adv(B.vmr_h2o, nlay); io.vmr_o3 = adv(B.vmr_o3, nlay);
    io.vmr = c.vmr_kind == RRTMGP_B200_VMR_GM ? (const FT*)B.vmr : adv(B.vmr, nlay * c.ngas);
    io.cld_r_eff_liq = adv(B.cld_r_eff_liq, nlay); io.cld_r_eff_ice = adv(B.cld_r_eff_ice, nlay);
    io.cld_path_liq = adv(B.cld_path_liq, nlay); io.cld_path_ice = adv(B.cld_path_ice, nlay);
    io.cld_frac = adv(B.cld_frac, nlay);
    io.aero_mass = adv(B.aero_mass, nlay * 15); io.aero_size = adv(B.aero_size, nlay * 15);
    io.sfc_emis = adv(B.sfc_emis, h->luts.n_bnd_lw); io.inc_flux_lw = adv(B.inc_flux_lw, 1);
    io.cos_zenith = adv(B.cos_zenith, 1); io.toa_flux = adv(B.toa_flux, 1);
    io.sfc_alb_direct = adv(B.sfc_alb_direct, h->luts.n_bnd_sw); io.sfc_alb_diffuse = adv(B.sfc_alb_diffuse, h->luts.n_bnd_sw);
    io.metric_scaling = adv(B.metric_scaling, nlev);
}

// src/optics/AngularDiscretizations.jl:34-63
template <typename FT> void gauss_angles(int n, FT* Ds, FT* wts) {
    static const double mu[4][4] = {{0.6096748751, 0, 0, 0},
                                    {0.2509907356, 0.7908473988, 0, 0},
                                    {0.1024922169, 0.4417960320, 0.8633751621, 0},
                                    {0.0454586727, 0.2322334416, 0.5740198775, 0.9030775973}};
    static const double w[4][4] = {{1, 0, 0, 0},
                                   {0.2300253764, 0.7699746236, 0, 0},
                                   {0.0437820218, 0.3875796738, 0.5686383044, 0},
                                   {0.0092068785, 0.1285704278, 0.4323381850, 0.4298845087}};
    for (int i = 0; i < 4; ++i) { Ds[i] = FT(0); wts[i] = FT(0); }
    for (int i = 0; i < n; ++i) { Ds[i] = (FT)(1.0 / mu[n - 1][i]); wts[i] = (FT)w[n - 1][i]; }
}

template <typename FT>
void base_params(const rrtmgp_b200_handle* h, SolveParams<FT>& P, bool sw, bool clouds, unsigned long long seed, long long c0, int count) {
    const rrtmgp_b200_config_t& c = h->cfg;
    const Luts<FT>& L = luts_of<FT>(h);
    std::memset(&P, 0, sizeof(P));
    fill_io(h, P.io, c0);
    P.lut = sw ? L.sw : L.lw;
    P.cld = sw ? L.cld_sw : L.cld_lw;
    P.aero = sw ? L.aero_sw : L.aero_lw;
    P.ncol = count; P.ncol_total = c.ncol; P.nlay = c.nlay; P.ngas = c.ngas; P.vmr_kind = c.vmr_kind; P.ice_rgh = c.ice_rgh;
    P.use_cloud = clouds && h->buf.cld_frac != nullptr;
    P.use_aero = c.aerosol_radiation && h->buf.aero_mass != nullptr;
    P.n_mu = c.op_lw == RRTMGP_B200_ONE_SCALAR ? c.n_gauss_angles : 1;
    P.col_offset = c.col_offset + c0;
    P.seed = seed;
    P.work_counter = h->work_counters ? h->work_counters + (h->work_seq++ % 255u) : nullptr;   // slot 255: validate_inputs
    // small shards: the number of work items per column follows from the HANDLE's column count, so every launch of a
    // handle (full range or column ranges, which touch disjoint columns of the scratch) sums in the same order
    {
        const int warps = (c.nlay > 64 ? 8 : 12) * (h->sm_count > 0 ? h->sm_count : 148);
        const long long per_warp = (long long)c.ncol / warps;
        // measured on B200 (tools/configs_sweep.py with RRTMGP_B200_SPLIT = 1 / 2 / 4): every halving of the work items
        // costs ~9 % (phase 0, set-up and the combination run per item; L1 hit rate of the gathers drops), so it only
        // pays below four columns per warp: 4096 columns 2.34e6 -> 2.52e6 col/s with two items per column, but 1e4
        // columns 2.68e6 -> 2.47e6 and four items per column never
        P.split = h->split_scratch == nullptr ? 1 : (per_warp >= 4 ? 1 : 2);
        if (const char* e = std::getenv("RRTMGP_B200_SPLIT"))                  // A/B experiments: 1, 2 or 4
            if (h->split_scratch != nullptr) P.split = std::atoi(e) == 4 ? 4 : (std::atoi(e) == 2 ? 2 : 1);
        const size_t scr = 3 * (size_t)((c.nlay > 64 ? 96 : 64) + 4) + 4;      // solver_fast.cuh: kScr of the geometry in use
        P.split_scratch = h->split_scratch ? h->split_scratch + (size_t)c0 * P.split * scr : nullptr;
        P.split_flags = h->split_flags ? h->split_flags + c0 : nullptr;
    }
    gauss_angles<FT>(P.n_mu, P.Ds, P.wts);
}

unsigned long long effective_seed(rrtmgp_b200_handle* h, uint64_t seed, int have_seed) {
    // `seedval = nothing` (update_fluxes.jl:150-156): the RNG stream simply continues; here a
    // per-call counter keys the draws so successive unseeded calls sample differently.
    return have_seed ? (unsigned long long)seed : 0xA5A5F00DULL + (++h->call_counter) * 0x632BE59BD9B4E019ULL;
}

// output arrays advanced to column c0: [ncol][nlev] (and band arrays, whose column stride stays ncol) / [ncol]
#define OFFL(T, p) ((p) ? (T*)(p) + c0 * (long long)(h->cfg.nlay + 1) : (T*)nullptr)
#define OFFC(T, p) ((p) ? (T*)(p) + c0 : (T*)nullptr)

template <typename FT>
int solve_lw_t(rrtmgp_b200_handle* h, unsigned long long seed, long long c0, int count, cudaStream_t s) {
    const rrtmgp_b200_config_t& c = h->cfg;
    const rrtmgp_b200_buffers_t& B = h->buf;
    const int mode = c.op_lw == RRTMGP_B200_ONE_SCALAR ? MODE_LW_NOSCAT : MODE_LW_2STREAM;
    const size_t nb = (size_t)h->luts.n_bnd_lw * c.ncol * (c.nlay + 1) * sizeof(FT);
    SolveParams<FT> P;
    if (c.method == RRTMGP_B200_ALL_SKY_WITH_CLEAR) {   // update_fluxes.jl:39-65
        base_params<FT>(h, P, false, false, seed, c0, count);
        P.io.out_up = OFFL(FT, B.clear_lw_flux_up); P.io.out_dn = OFFL(FT, B.clear_lw_flux_dn); P.io.out_net = OFFL(FT, B.clear_lw_flux_net);
        int e = launch_solve<FT>(mode, P, h->max_smem_optin, s);
        if (e) return fail_cuda(h, (cudaError_t)e);
        ++h->last_launches;
    }
    base_params<FT>(h, P, false, c.method >= RRTMGP_B200_ALL_SKY, seed, c0, count);
    P.io.out_up = OFFL(FT, B.lw_flux_up); P.io.out_dn = OFFL(FT, B.lw_flux_dn); P.io.out_net = OFFL(FT, B.lw_flux_net);
    P.io.cld_cover = OFFC(FT, B.cld_cover_lw);
    if (c.spectral_fluxes && mode == MODE_LW_2STREAM) {
        P.io.band_up = OFFL(FT, B.lw_band_flux_up); P.io.band_dn = OFFL(FT, B.lw_band_flux_dn); P.io.band_net = OFFL(FT, B.lw_band_flux_net);
        cudaError_t e = cudaMemsetAsync(B.lw_band_flux_up, 0, nb, s);   // set_band_flux_to_zero! (Fluxes.jl:191-197)
        if (e == cudaSuccess) e = cudaMemsetAsync(B.lw_band_flux_dn, 0, nb, s);
        if (e != cudaSuccess) return fail_cuda(h, e);
    }
    int e = launch_solve<FT>(mode, P, h->max_smem_optin, s);
    if (e) return fail_cuda(h, (cudaError_t)e);
    ++h->last_launches;
    return RRTMGP_B200_OK;
}

template <typename FT>
int solve_sw_t(rrtmgp_b200_handle* h, unsigned long long seed, bool fuse_net, long long c0, int count, cudaStream_t s) {
    const rrtmgp_b200_config_t& c = h->cfg;
    const rrtmgp_b200_buffers_t& B = h->buf;
    const size_t nb = (size_t)h->luts.n_bnd_sw * c.ncol * (c.nlay + 1) * sizeof(FT);
    SolveParams<FT> P;
    if (c.method == RRTMGP_B200_ALL_SKY_WITH_CLEAR) {   // update_fluxes.jl:101-128
        base_params<FT>(h, P, true, false, seed, c0, count);
        P.io.out_up = OFFL(FT, B.clear_sw_flux_up); P.io.out_dn = OFFL(FT, B.clear_sw_flux_dn); P.io.out_net = OFFL(FT, B.clear_sw_flux_net);
        P.io.out_dir = OFFL(FT, B.clear_sw_flux_dn_dir);
        P.io.aod_ext = OFFC(FT, B.aod_sw_ext); P.io.aod_sca = OFFC(FT, B.aod_sw_sca);
        if (fuse_net && B.clear_net_flux) { P.io.add_net = OFFL(const FT, B.clear_lw_flux_net); P.io.out_total_net = OFFL(FT, B.clear_net_flux); }
        int e = launch_solve<FT>(MODE_SW_2STREAM, P, h->max_smem_optin, s);
        if (e) return fail_cuda(h, (cudaError_t)e);
        ++h->last_launches;
    }
    base_params<FT>(h, P, true, c.method >= RRTMGP_B200_ALL_SKY, seed, c0, count);
    P.io.out_up = OFFL(FT, B.sw_flux_up); P.io.out_dn = OFFL(FT, B.sw_flux_dn); P.io.out_net = OFFL(FT, B.sw_flux_net);
    P.io.out_dir = OFFL(FT, B.sw_flux_dn_dir);
    P.io.cld_cover = OFFC(FT, B.cld_cover_sw);
    P.io.aod_ext = OFFC(FT, B.aod_sw_ext); P.io.aod_sca = OFFC(FT, B.aod_sw_sca);
    if (fuse_net && B.net_flux) { P.io.add_net = OFFL(const FT, B.lw_flux_net); P.io.out_total_net = OFFL(FT, B.net_flux); }
    if (c.spectral_fluxes) {
        P.io.band_up = OFFL(FT, B.sw_band_flux_up); P.io.band_dn = OFFL(FT, B.sw_band_flux_dn); P.io.band_net = OFFL(FT, B.sw_band_flux_net);
        cudaError_t e = cudaMemsetAsync(B.sw_band_flux_up, 0, nb, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(B.sw_band_flux_dn, 0, nb, s);
        if (e != cudaSuccess) return fail_cuda(h, e);
    }
    int e = launch_solve<FT>(MODE_SW_2STREAM, P, h->max_smem_optin, s);
    if (e) return fail_cuda(h, (cudaError_t)e);
    ++h->last_launches;
    return RRTMGP_B200_OK;
}

constexpr unsigned kAllPrepareSteps = RRTMGP_B200_STEP_INTERPOLATE_LEVELS | RRTMGP_B200_STEP_BOUNDARY_LAYER |
                                      RRTMGP_B200_STEP_CLIP | RRTMGP_B200_STEP_CONCENTRATIONS;

template <typename FT> int prepare_t(rrtmgp_b200_handle* h, long long c0, int count, cudaStream_t s, unsigned steps = kAllPrepareSteps) {
    const rrtmgp_b200_config_t& c = h->cfg;
    const LutStore& L = h->luts;
    const int idx_h2o = luts_of<FT>(h).lw.idx_h2o;
    if ((steps & RRTMGP_B200_STEP_INTERPOLATE_LEVELS) && h->interpolation != RRTMGP_B200_NO_INTERPOLATION) {   // update_fluxes.jl:256-264
        const int nlay_dom = c.nlay - (c.isothermal_boundary_layer ? 1 : 0);
        const long long nf = (long long)count * (nlay_dom + 1);
        interpolate_levels_kernel<FT><<<(unsigned)((nf + 255) / 256), 256, 0, s>>>(
            h->buf, c0, count, c.nlay, nlay_dom, h->interpolation, h->bottom_extrapolation, (const FT*)h->center_z,
            (const FT*)h->face_z, (FT)c.grav, (FT)h->cp_d, (FT)h->r_d);
        ++h->last_launches;
    }
    if ((steps & RRTMGP_B200_STEP_BOUNDARY_LAYER) && c.isothermal_boundary_layer) {
        boundary_layer_kernel<FT><<<(count + 127) / 128, 128, 0, s>>>(h->buf, c0, count, c.nlay, c.ngas, c.vmr_kind, (FT)L.p_ref_min);
        ++h->last_launches;
    }
    const int do_clip = (steps & RRTMGP_B200_STEP_CLIP) ? 1 : 0, do_col = (steps & RRTMGP_B200_STEP_CONCENTRATIONS) ? 1 : 0;
    if (do_clip || do_col) {
        const long long n = (long long)count * (c.nlay + 1);
        prepare_kernel<FT><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->buf, c0, count, c.nlay, c.ngas, c.vmr_kind, idx_h2o,
                                                                      luts_of<FT>(h).lw.p_ref_min, luts_of<FT>(h).lw.t_ref_min,
                                                                      luts_of<FT>(h).lw.t_ref_max, (FT)c.grav, (FT)c.molmass_dryair,
                                                                      (FT)c.molmass_water, (FT)c.avogad, do_clip, do_col);
        ++h->last_launches;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? RRTMGP_B200_OK : fail_cuda(h, e);
}

template <typename FT> int net_t(rrtmgp_b200_handle* h, cudaStream_t s) {
    const rrtmgp_b200_config_t& c = h->cfg;
    const rrtmgp_b200_buffers_t& B = h->buf;
    const long long n = (long long)c.ncol * (c.nlay + 1);
    if (B.net_flux) {
        add_kernel<FT><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const FT*)B.lw_flux_net, (const FT*)B.sw_flux_net, (FT*)B.net_flux, n);
        ++h->last_launches;
    }
    if (c.method == RRTMGP_B200_ALL_SKY_WITH_CLEAR && B.clear_net_flux) {
        add_kernel<FT><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const FT*)B.clear_lw_flux_net, (const FT*)B.clear_sw_flux_net,
                                                                  (FT*)B.clear_net_flux, n);
        ++h->last_launches;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? RRTMGP_B200_OK : fail_cuda(h, e);
}

int ready(const rrtmgp_b200_handle* h) {
    if (!h) return RRTMGP_B200_ERR_INVALID_ARG;
    if (!h->luts.loaded || !h->bound) return RRTMGP_B200_ERR_NOT_READY;
    return RRTMGP_B200_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

template <typename FT> static int validate_t(rrtmgp_b200_handle* h, uint32_t* failed, cudaStream_t s) {
    const rrtmgp_b200_config_t& c = h->cfg;
    const rrtmgp_b200_buffers_t& B = h->buf;
    unsigned* mask = h->work_counters + 255;   // the last slot of the counter ring is never handed to a kernel launch
    if (h->work_counters == nullptr) return RRTMGP_B200_ERR_NOT_READY;
    cudaError_t e = cudaMemsetAsync(mask, 0, sizeof(unsigned), s);
    if (e != cudaSuccess) return fail_cuda(h, e);
    h->last_launches = 0;
    auto run = [&](const void* p, long long n, int inner, long long stride, long long offset, int mode, unsigned bit) {
        if (p == nullptr || n <= 0) return;   // `_check(f, ::Nothing, name) = true`
        validate_kernel<FT><<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const FT*)p, n, inner, stride, offset, mode, bit, mask);
        ++h->last_launches;
    };
    const long long ncol = c.ncol, nlay = c.nlay, nlev = c.nlay + 1;
    const int nb_lw = h->luts.n_bnd_lw, nb_sw = h->luts.n_bnd_sw;
    run(B.p_lev, ncol * nlev, 1, 1, 0, 0, RRTMGP_B200_BAD_LEVEL_PRESSURE);
    run(B.t_lev, ncol * nlev, 1, 1, 0, 0, RRTMGP_B200_BAD_LEVEL_TEMPERATURE);
    run(B.layerdata, ncol * nlay, 1, 4, 1, 0, RRTMGP_B200_BAD_LAYER_PRESSURE);        // layerdata[..][1] = p_lay
    run(B.layerdata, ncol * nlay, 1, 4, 2, 0, RRTMGP_B200_BAD_LAYER_TEMPERATURE);     // layerdata[..][2] = t_lay
    run(B.t_sfc, ncol, 1, 1, 0, 0, RRTMGP_B200_BAD_SURFACE_TEMPERATURE);
    run(B.cos_zenith, ncol, 1, 1, 0, 3, RRTMGP_B200_BAD_COS_ZENITH);
    run(B.toa_flux, ncol, 1, 1, 0, 1, RRTMGP_B200_BAD_TOA_SW_FLUX_DN);
    run(B.sfc_emis, ncol * nb_lw, 1, 1, 0, 2, RRTMGP_B200_BAD_SURFACE_EMISSIVITY);
    run(B.sfc_alb_direct, ncol * nb_sw, 1, 1, 0, 2, RRTMGP_B200_BAD_DIRECT_SW_SURFACE_ALBEDO);
    run(B.sfc_alb_diffuse, ncol * nb_sw, 1, 1, 0, 2, RRTMGP_B200_BAD_DIFFUSE_SW_SURFACE_ALBEDO);
    if (c.vmr_kind == RRTMGP_B200_VMR_GM) {   // _check_vmr (validation.jl:39-45)
        run(B.vmr_h2o, ncol * nlay, 1, 1, 0, 1, RRTMGP_B200_BAD_VMR_H2O);
        run(B.vmr_o3, ncol * nlay, 1, 1, 0, 1, RRTMGP_B200_BAD_VMR_O3);
        run(B.vmr, c.ngas, 1, 1, 0, 1, RRTMGP_B200_BAD_VMR);
    } else {
        run(B.vmr, ncol * nlay * c.ngas, 1, 1, 0, 1, RRTMGP_B200_BAD_VMR);
    }
    unsigned host = 0;
    e = cudaMemcpyAsync(&host, mask, sizeof(unsigned), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail_cuda(h, e);
    *failed = host;
    return RRTMGP_B200_OK;
}

extern "C" {

int rrtmgp_b200_abi_version(void) { return RRTMGP_B200_ABI_VERSION; }

const char* rrtmgp_b200_strerror(int status) {
    switch (status) {
        case RRTMGP_B200_OK: return "ok";
        case RRTMGP_B200_ERR_INVALID_ARG: return "invalid argument or unsupported option combination";
        case RRTMGP_B200_ERR_BAD_LUT_PACK: return "malformed or incomplete LUT pack";
        case RRTMGP_B200_ERR_NOT_READY: return "lookup tables not loaded or buffers not bound";
        case RRTMGP_B200_ERR_CUDA: return "CUDA runtime error";
        case RRTMGP_B200_ERR_UNSUPPORTED: return "configuration not supported by this build";
    }
    return "unknown status";
}

int rrtmgp_b200_create(const rrtmgp_b200_config_t* cfg, rrtmgp_b200_handle_t** out) {
    if (!cfg || !out) return RRTMGP_B200_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->abi_version != RRTMGP_B200_ABI_VERSION) return RRTMGP_B200_ERR_INVALID_ARG;
    if (cfg->dtype != 0 && cfg->dtype != 1) return RRTMGP_B200_ERR_INVALID_ARG;
    if (cfg->ncol <= 0 || cfg->nlay < 2 || cfg->ngas <= 0) return RRTMGP_B200_ERR_INVALID_ARG;
    if (cfg->method < 0 || cfg->method > 2) return RRTMGP_B200_ERR_INVALID_ARG;
    if (cfg->vmr_kind != RRTMGP_B200_VMR_GM && cfg->vmr_kind != RRTMGP_B200_VMR_FULL) return RRTMGP_B200_ERR_INVALID_ARG;
    if (cfg->op_lw != RRTMGP_B200_TWO_STREAM && cfg->op_lw != RRTMGP_B200_ONE_SCALAR) return RRTMGP_B200_ERR_INVALID_ARG;
    if (cfg->n_gauss_angles < 1 || cfg->n_gauss_angles > 4) return RRTMGP_B200_ERR_INVALID_ARG;   // AngularDiscretizations.jl:40
    // solver.jl:159-171: n_gauss_angles only applies to the non-scattering longwave solver
    if (cfg->n_gauss_angles != 1 && cfg->op_lw != RRTMGP_B200_ONE_SCALAR) return RRTMGP_B200_ERR_INVALID_ARG;
    if (cfg->ice_rgh < 1 || cfg->ice_rgh > 3) return RRTMGP_B200_ERR_INVALID_ARG;
    if (cfg->nlay + 1 > 32 * kMaxLevPerLane) return RRTMGP_B200_ERR_UNSUPPORTED;
    if (!(cfg->grav > 0) || !(cfg->molmass_dryair > 0) || !(cfg->avogad > 0)) return RRTMGP_B200_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) return RRTMGP_B200_ERR_CUDA;
    rrtmgp_b200_handle* h = new (std::nothrow) rrtmgp_b200_handle();
    if (!h) return RRTMGP_B200_ERR_INVALID_ARG;
    h->cfg = *cfg;
    std::memset(&h->buf, 0, sizeof(h->buf));
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
    {
        DeviceGuard g(cfg->device);
        if (cudaMalloc(&h->work_counters, 256 * sizeof(unsigned int)) != cudaSuccess) { delete h; return RRTMGP_B200_ERR_CUDA; }
        if (cfg->dtype == 0 && cfg->ncol < kSplitMaxCols) {      // small Float32 shards: scratch of the split columns
            const size_t nscr = (size_t)cfg->ncol * 4 * (3 * (96 + 4) + 4);   // 4 shares, the larger accumulator stride (8-warp geometry)
            if (cudaMalloc(&h->split_scratch, nscr * sizeof(float)) != cudaSuccess ||
                cudaMalloc(&h->split_flags, (size_t)cfg->ncol * sizeof(unsigned int)) != cudaSuccess ||
                cudaMemset(h->split_flags, 0, (size_t)cfg->ncol * sizeof(unsigned int)) != cudaSuccess) {
                if (h->split_scratch) cudaFree(h->split_scratch);
                if (h->split_flags) cudaFree(h->split_flags);
                cudaFree(h->work_counters);
                delete h;
                return RRTMGP_B200_ERR_CUDA;
            }
        }
    }
    *out = h;
    return RRTMGP_B200_OK;
}

void rrtmgp_b200_destroy(rrtmgp_b200_handle_t* h) {
    if (!h) return;
    rrtmgp_b200_comm_destroy(h);
    DeviceGuard g(h->cfg.device);
    free_lut_store(h->luts);
    if (h->work_counters) cudaFree(h->work_counters);
    if (h->split_scratch) cudaFree(h->split_scratch);
    if (h->split_flags) cudaFree(h->split_flags);
    delete h;
}

int rrtmgp_b200_load_luts(rrtmgp_b200_handle_t* h, const void* pack, size_t nbytes) {
    if (!h || !pack) return RRTMGP_B200_ERR_INVALID_ARG;
    DeviceGuard g(h->cfg.device);
    const char* err = nullptr;
    int st = load_lut_pack(h->luts, pack, nbytes, h->cfg.dtype == 1, &err);
    if (st == RRTMGP_B200_ERR_CUDA && err) std::snprintf(h->cuda_err, sizeof(h->cuda_err), "%s", err);
    if (st != RRTMGP_B200_OK) return st;
    // Every failure below leaves the handle WITHOUT tables (`ready()` keeps refusing), so a caller that ignores the
    // status cannot reach the kernels with tables they would index out of bounds.
    auto refuse = [&](int code) { free_lut_store(h->luts); return code; };
    if (h->luts.ngas > h->cfg.ngas) return refuse(RRTMGP_B200_ERR_INVALID_ARG);   // vmr gas axis shorter than the tables'
    // a pack without the cloud / aerosol sections only serves the methods that never read them
    // (lookup_tables, ext/RRTMGPNCDatasetsExt.jl:26-133, loads exactly what the method needs)
    const bool f64 = h->cfg.dtype == 1;
    const bool has_cld = f64 ? h->luts.f64.cld_lw.liqdata != nullptr : h->luts.f32.cld_lw.liqdata != nullptr;
    const bool has_aer = f64 ? h->luts.f64.aero_lw.dust != nullptr : h->luts.f32.aero_lw.dust != nullptr;
    if ((h->cfg.method >= RRTMGP_B200_ALL_SKY && !has_cld) || (h->cfg.aerosol_radiation && !has_aer))
        return refuse(RRTMGP_B200_ERR_BAD_LUT_PACK);
    // limits of the fields the kernels bit-pack (solver.cuh): eta cell in 4 bits, T / p node in 8 bits each,
    // cloud size interval in 8 bits, MERRA size bin in 3 bits
    auto limits_ok = [&](auto& L) {
        for (auto* g : {&L.lw, &L.sw})
            if (g->n_eta > 16 || g->n_eta < 2 || g->n_t > 255 || g->n_t < 2 || g->n_p > 255 || g->n_p_ref < 2) return false;
        for (auto* c : {&L.cld_lw, &L.cld_sw})
            if (c->liqdata != nullptr && (c->nsize_liq > 255 || c->nsize_ice > 255 || c->nsize_liq < 2 || c->nsize_ice < 2 ||
                                          c->nrghice < h->cfg.ice_rgh)) return false;
        for (auto* a : {&L.aero_lw, &L.aero_sw})
            if (a->dust != nullptr && (a->nbin > 7 || a->nbin < 1 || a->nrh < 2)) return false;
        return true;
    };
    if (!(f64 ? limits_ok(h->luts.f64) : limits_ok(h->luts.f32))) return refuse(RRTMGP_B200_ERR_UNSUPPORTED);
    return RRTMGP_B200_OK;
}

int rrtmgp_b200_lut_info(const rrtmgp_b200_handle_t* h, rrtmgp_b200_lut_info_t* o) {
    if (!h || !o) return RRTMGP_B200_ERR_INVALID_ARG;
    if (!h->luts.loaded) return RRTMGP_B200_ERR_NOT_READY;
    const LutStore& L = h->luts;
    o->n_gpt_lw = L.n_gpt_lw; o->n_bnd_lw = L.n_bnd_lw; o->n_gpt_sw = L.n_gpt_sw; o->n_bnd_sw = L.n_bnd_sw;
    o->ngas = L.ngas; o->iband_550nm = L.iband_550nm;
    o->p_ref_min = L.p_ref_min; o->t_ref_min = L.t_ref_min; o->t_ref_max = L.t_ref_max; o->solar_src_tot = L.solar_src_tot;
    return RRTMGP_B200_OK;
}

int rrtmgp_b200_bind(rrtmgp_b200_handle_t* h, const rrtmgp_b200_buffers_t* b) {
    if (!h || !b) return RRTMGP_B200_ERR_INVALID_ARG;
    const rrtmgp_b200_config_t& c = h->cfg;
    if (!b->layerdata || !b->p_lev || !b->t_lev || !b->t_sfc || !b->vmr) return RRTMGP_B200_ERR_INVALID_ARG;
    if (c.vmr_kind == RRTMGP_B200_VMR_GM && (!b->vmr_h2o || !b->vmr_o3)) return RRTMGP_B200_ERR_INVALID_ARG;
    if (!b->sfc_emis || !b->cos_zenith || !b->toa_flux || !b->sfc_alb_direct || !b->sfc_alb_diffuse) return RRTMGP_B200_ERR_INVALID_ARG;
    if (!b->lw_flux_up || !b->lw_flux_dn || !b->lw_flux_net) return RRTMGP_B200_ERR_INVALID_ARG;
    if (!b->sw_flux_up || !b->sw_flux_dn || !b->sw_flux_net || !b->sw_flux_dn_dir) return RRTMGP_B200_ERR_INVALID_ARG;
    if (c.method >= RRTMGP_B200_ALL_SKY &&
        (!b->cld_frac || !b->cld_r_eff_liq || !b->cld_r_eff_ice || !b->cld_path_liq || !b->cld_path_ice))
        return RRTMGP_B200_ERR_INVALID_ARG;
    if (c.aerosol_radiation && (!b->aero_mass || !b->aero_size)) return RRTMGP_B200_ERR_INVALID_ARG;
    if ((b->aod_sw_ext == nullptr) != (b->aod_sw_sca == nullptr)) return RRTMGP_B200_ERR_INVALID_ARG;
    if (c.method == RRTMGP_B200_ALL_SKY_WITH_CLEAR &&
        (!b->clear_lw_flux_up || !b->clear_lw_flux_dn || !b->clear_lw_flux_net || !b->clear_sw_flux_up ||
         !b->clear_sw_flux_dn || !b->clear_sw_flux_net || !b->clear_sw_flux_dn_dir))
        return RRTMGP_B200_ERR_INVALID_ARG;
    if (c.spectral_fluxes && (!b->lw_band_flux_up || !b->lw_band_flux_dn || !b->lw_band_flux_net || !b->sw_band_flux_up ||
                              !b->sw_band_flux_dn || !b->sw_band_flux_net))
        return RRTMGP_B200_ERR_INVALID_ARG;
    h->buf = *b;
    h->bound = true;
    return RRTMGP_B200_OK;
}

#define RB_DISPATCH(h, call) ((h)->cfg.dtype == 1 ? call<double> : call<float>)

int rrtmgp_b200_prepare_atmosphere(rrtmgp_b200_handle_t* h, void* stream) {
    int st = ready(h);
    if (st) return st;
    DeviceGuard g(h->cfg.device);
    h->last_launches = 0;
    return h->cfg.dtype == 1 ? prepare_t<double>(h, 0, h->cfg.ncol, (cudaStream_t)stream) : prepare_t<float>(h, 0, h->cfg.ncol, (cudaStream_t)stream);
}

int rrtmgp_b200_prepare_steps(rrtmgp_b200_handle_t* h, uint32_t steps, void* stream) {
    int st = ready(h);
    if (st) return st;
    if (steps == 0 || (steps & ~kAllPrepareSteps)) return RRTMGP_B200_ERR_INVALID_ARG;
    DeviceGuard g(h->cfg.device);
    h->last_launches = 0;
    return h->cfg.dtype == 1 ? prepare_t<double>(h, 0, h->cfg.ncol, (cudaStream_t)stream, steps)
                             : prepare_t<float>(h, 0, h->cfg.ncol, (cudaStream_t)stream, steps);
}

int rrtmgp_b200_update_lw_fluxes(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, void* stream) {
    int st = ready(h);
    if (st) return st;
    DeviceGuard g(h->cfg.device);
    h->last_launches = 0;
    unsigned long long s = effective_seed(h, seed, have_seed);
    return h->cfg.dtype == 1 ? solve_lw_t<double>(h, s, 0, h->cfg.ncol, (cudaStream_t)stream) : solve_lw_t<float>(h, s, 0, h->cfg.ncol, (cudaStream_t)stream);
}

int rrtmgp_b200_update_sw_fluxes(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, void* stream) {
    int st = ready(h);
    if (st) return st;
    DeviceGuard g(h->cfg.device);
    h->last_launches = 0;
    unsigned long long s = effective_seed(h, seed, have_seed);
    return h->cfg.dtype == 1 ? solve_sw_t<double>(h, s, false, 0, h->cfg.ncol, (cudaStream_t)stream)
                             : solve_sw_t<float>(h, s, false, 0, h->cfg.ncol, (cudaStream_t)stream);
}

int rrtmgp_b200_update_net_fluxes(rrtmgp_b200_handle_t* h, void* stream) {
    int st = ready(h);
    if (st) return st;
    DeviceGuard g(h->cfg.device);
    h->last_launches = 0;
    return h->cfg.dtype == 1 ? net_t<double>(h, (cudaStream_t)stream) : net_t<float>(h, (cudaStream_t)stream);
}

static int update_range(rrtmgp_b200_handle_t* h, unsigned long long sd, long long c0, int count, cudaStream_t s) {
    const bool f64 = h->cfg.dtype == 1;
    int st = f64 ? prepare_t<double>(h, c0, count, s) : prepare_t<float>(h, c0, count, s);
    if (st) return st;
    st = f64 ? solve_lw_t<double>(h, sd, c0, count, s) : solve_lw_t<float>(h, sd, c0, count, s);
    if (st) return st;
    // net = lw_net + sw_net is folded into the shortwave epilogue (update_fluxes.jl:165-194)
    return f64 ? solve_sw_t<double>(h, sd, true, c0, count, s) : solve_sw_t<float>(h, sd, true, c0, count, s);
}

int rrtmgp_b200_update_fluxes(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, void* stream) {
    int st = ready(h);
    if (st) return st;
    DeviceGuard g(h->cfg.device);
    h->last_launches = 0;
    return update_range(h, effective_seed(h, seed, have_seed), 0, h->cfg.ncol, (cudaStream_t)stream);
}

int rrtmgp_b200_update_fluxes_range(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, int64_t col_begin,
                                    int32_t col_count, void* stream) {
    int st = ready(h);
    if (st) return st;
    if (col_begin < 0 || col_count <= 0 || col_begin + col_count > h->cfg.ncol) return RRTMGP_B200_ERR_INVALID_ARG;
    if (!have_seed) return RRTMGP_B200_ERR_INVALID_ARG;          // every range of one step must share the seed
    if (h->cfg.spectral_fluxes) return RRTMGP_B200_ERR_UNSUPPORTED;
    DeviceGuard g(h->cfg.device);
    h->last_launches = 0;
    return update_range(h, (unsigned long long)seed, col_begin, col_count, (cudaStream_t)stream);
}

int rrtmgp_b200_set_level_interpolation(rrtmgp_b200_handle_t* h, int32_t interpolation, int32_t bottom_extrapolation,
                                        const void* center_z, const void* face_z, double cp_d, double R_d) {
    if (!h) return RRTMGP_B200_ERR_INVALID_ARG;
    if (interpolation < RRTMGP_B200_NO_INTERPOLATION || interpolation > RRTMGP_B200_BEST_FIT) return RRTMGP_B200_ERR_INVALID_ARG;
    if (bottom_extrapolation < RRTMGP_B200_SAME_AS_INTERPOLATION || bottom_extrapolation > RRTMGP_B200_HYDROSTATIC_BOTTOM)
        return RRTMGP_B200_ERR_INVALID_ARG;
    // solver.jl:183-193: the z-based schemes read the altitudes at solve time
    const bool needs_z = interpolation == RRTMGP_B200_BEST_FIT ||
                         (interpolation != RRTMGP_B200_NO_INTERPOLATION && bottom_extrapolation == RRTMGP_B200_HYDROSTATIC_BOTTOM);
    if (needs_z && (!center_z || !face_z)) return RRTMGP_B200_ERR_INVALID_ARG;
    if (interpolation != RRTMGP_B200_NO_INTERPOLATION && bottom_extrapolation != RRTMGP_B200_SAME_AS_INTERPOLATION &&
        (!(cp_d > 0) || !(R_d > 0)))
        return RRTMGP_B200_ERR_INVALID_ARG;
    if (interpolation != RRTMGP_B200_NO_INTERPOLATION && h->cfg.nlay - (h->cfg.isothermal_boundary_layer ? 1 : 0) < 2)
        return RRTMGP_B200_ERR_INVALID_ARG;
    h->interpolation = interpolation; h->bottom_extrapolation = bottom_extrapolation;
    h->center_z = center_z; h->face_z = face_z; h->cp_d = cp_d; h->r_d = R_d;
    return RRTMGP_B200_OK;
}

int rrtmgp_b200_heating_rate(rrtmgp_b200_handle_t* h, const void* flux_net, void* heating_rate, double cp_d, void* stream) {
    int st = ready(h);
    if (st) return st;
    if (!flux_net || !heating_rate || !(cp_d > 0)) return RRTMGP_B200_ERR_INVALID_ARG;
    DeviceGuard g(h->cfg.device);
    h->last_launches = 1;
    const rrtmgp_b200_config_t& c = h->cfg;
    const int nlay_dom = c.nlay - (c.isothermal_boundary_layer ? 1 : 0);
    const long long n = (long long)c.ncol * nlay_dom;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (c.dtype == 1)
        heating_rate_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)flux_net, (const double*)h->buf.p_lev,
                                                                            (double*)heating_rate, c.ncol, c.nlay, nlay_dom, c.grav, cp_d);
    else
        heating_rate_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)flux_net, (const float*)h->buf.p_lev,
                                                                           (float*)heating_rate, c.ncol, c.nlay, nlay_dom, (float)c.grav, (float)cp_d);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? RRTMGP_B200_OK : fail_cuda(h, e);
}

int rrtmgp_b200_compute_relative_humidity(rrtmgp_b200_handle_t* h, void* stream) {
    int st = ready(h);
    if (st) return st;
    DeviceGuard g(h->cfg.device);
    h->last_launches = 1;
    const rrtmgp_b200_config_t& c = h->cfg;
    const long long n = (long long)c.ncol * c.nlay;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (c.dtype == 1)
        rel_hum_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(h->buf, c.ncol, c.nlay, c.ngas, c.vmr_kind, h->luts.f64.lw.idx_h2o,
                                                                       c.molmass_water / c.molmass_dryair);
    else
        rel_hum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(h->buf, c.ncol, c.nlay, c.ngas, c.vmr_kind, h->luts.f32.lw.idx_h2o,
                                                                      (float)c.molmass_water / (float)c.molmass_dryair);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? RRTMGP_B200_OK : fail_cuda(h, e);
}

int rrtmgp_b200_validate_inputs(rrtmgp_b200_handle_t* h, uint32_t* failed, void* stream) {
    if (failed == nullptr) return RRTMGP_B200_ERR_INVALID_ARG;
    int st = ready(h);
    if (st) return st;
    DeviceGuard g(h->cfg.device);
    return h->cfg.dtype == 1 ? validate_t<double>(h, failed, (cudaStream_t)stream) : validate_t<float>(h, failed, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------------
// multi-GPU (include/rrtmgp_b200.h "multi-GPU")
// ------------------------------------------------------------------------------------------------------------------
static bool load_nccl(NcclApi& n) {
    if (n.lib) return true;
    // libnccl.so.2 is already mapped when the host uses torch.distributed / NCCL.jl; otherwise the loader path decides
    n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) n.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) return false;
#define RB_NCCL_SYM(field, name) n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.lib, name)); if (!n.field) return false;
    RB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") RB_NCCL_SYM(CommInitRank, "ncclCommInitRank") RB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    RB_NCCL_SYM(AllGather, "ncclAllGather") RB_NCCL_SYM(AllReduce, "ncclAllReduce") RB_NCCL_SYM(GroupStart, "ncclGroupStart")
    RB_NCCL_SYM(GroupEnd, "ncclGroupEnd") RB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef RB_NCCL_SYM
    return true;
}

static int fail_nccl(rrtmgp_b200_handle* h, const NcclApi& n, int rc) {
    std::snprintf(h->cuda_err, sizeof(h->cuda_err), "NCCL: %s", n.GetErrorString ? n.GetErrorString(rc) : "error");
    return RRTMGP_B200_ERR_CUDA;
}

// the eight local flux views in the order of rrtmgp_b200_gathered_t
static void local_views(const rrtmgp_b200_handle* h, void* v[kGatheredViews]) {
    const rrtmgp_b200_buffers_t& B = h->buf;
    void* t[kGatheredViews] = {B.lw_flux_up, B.lw_flux_dn, B.lw_flux_net, B.sw_flux_up, B.sw_flux_dn, B.sw_flux_net, B.sw_flux_dn_dir, B.net_flux};
    for (int i = 0; i < kGatheredViews; ++i) v[i] = t[i];
}

int rrtmgp_b200_comm_unique_id(void* id_out, size_t nbytes) {
    if (!id_out || nbytes < RRTMGP_B200_UNIQUE_ID_BYTES) return RRTMGP_B200_ERR_INVALID_ARG;
    NcclApi n;
    if (!load_nccl(n)) return RRTMGP_B200_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != 0) return RRTMGP_B200_ERR_CUDA;
    std::memcpy(id_out, id.internal, sizeof(id.internal));
    return RRTMGP_B200_OK;
}

int rrtmgp_b200_comm_destroy(rrtmgp_b200_handle_t* h) {
    if (!h) return RRTMGP_B200_ERR_INVALID_ARG;
    CommState* c = h->comm;
    if (!c) return RRTMGP_B200_OK;
    DeviceGuard g(h->cfg.device);
    cudaDeviceSynchronize();
    for (int p = 0; p < c->nranks; ++p)
        if (p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
    // exported memory must outlive every importer's mapping: all ranks close first, then everybody frees
    if (c->ready && c->comm && c->token && c->copy_stream && c->nranks > 1 &&
        c->nccl.AllReduce(c->token, c->token, 1, kNcclFloat32, kNcclSum, c->comm, c->copy_stream) == 0)
        cudaStreamSynchronize(c->copy_stream);
    if (c->comm) c->nccl.CommDestroy(c->comm);
    if (c->gathered) cudaFree(c->gathered);
    if (c->token) cudaFree(c->token);
    for (auto& e : c->ev_kernel) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_copied) if (e) cudaEventDestroy(e);
    for (auto& x : c->copy_extra) if (x) cudaStreamDestroy(x);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
    h->comm = nullptr;
    return RRTMGP_B200_OK;
}

int rrtmgp_b200_comm_init(rrtmgp_b200_handle_t* h, const void* unique_id, int32_t rank, int32_t nranks) {
    if (!h || !unique_id || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return RRTMGP_B200_ERR_INVALID_ARG;
    if (h->comm) return RRTMGP_B200_ERR_INVALID_ARG;   // one communicator per handle
    DeviceGuard g(h->cfg.device);
    CommState* c = new (std::nothrow) CommState();
    if (!c) return RRTMGP_B200_ERR_INVALID_ARG;
    h->comm = c;
    c->rank = rank; c->nranks = nranks;
    auto bail = [&](int code) { rrtmgp_b200_comm_destroy(h); return code; };
    if (!load_nccl(c->nccl)) { std::snprintf(h->cuda_err, sizeof(h->cuda_err), "libnccl.so.2 not found"); return bail(RRTMGP_B200_ERR_UNSUPPORTED); }
    ncclUniqueId id;
    std::memcpy(id.internal, unique_id, sizeof(id.internal));
    int rc = c->nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (rc != 0) { fail_nccl(h, c->nccl, rc); return bail(RRTMGP_B200_ERR_CUDA); }
    cudaError_t e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (auto& ev : c->ev_kernel) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto& ev : c->ev_copied) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto& x : c->copy_extra) if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
    const size_t esz = h->cfg.dtype == 1 ? 8 : 4;
    c->view_bytes = (size_t)nranks * h->cfg.ncol * (h->cfg.nlay + 1) * esz;
    // an allocation of its own (cudaMalloc, not a pool): CUDA IPC exports whole allocations
    if (e == cudaSuccess) e = cudaMalloc(&c->gathered, kGatheredViews * c->view_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->gathered, 0, kGatheredViews * c->view_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->token, 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(c->token, 0, 2 * sizeof(float));
    if (e != cudaSuccess) { fail_cuda(h, e); return bail(RRTMGP_B200_ERR_CUDA); }
    // exchange (IPC handle, ncol, nlay, dtype) of every rank through the communicator itself
    struct Card { cudaIpcMemHandle_t handle; int32_t ncol, nlay, dtype, pad; };
    static_assert(sizeof(Card) % 8 == 0, "card is exchanged as bytes");
    Card mine{};
    e = cudaIpcGetMemHandle(&mine.handle, c->gathered);
    mine.ncol = h->cfg.ncol; mine.nlay = h->cfg.nlay; mine.dtype = h->cfg.dtype;
    Card* d_cards = nullptr;
    Card cards[kMaxRanks];
    if (e == cudaSuccess) e = cudaMalloc(&d_cards, nranks * sizeof(Card));
    if (e == cudaSuccess) e = cudaMemcpy(d_cards + rank, &mine, sizeof(Card), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { if (d_cards) cudaFree(d_cards); fail_cuda(h, e); return bail(RRTMGP_B200_ERR_CUDA); }
    rc = c->nccl.AllGather(d_cards + rank, d_cards, sizeof(Card), kNcclChar, c->comm, c->copy_stream);
    if (rc == 0) e = cudaStreamSynchronize(c->copy_stream);
    if (rc == 0 && e == cudaSuccess) e = cudaMemcpy(cards, d_cards, nranks * sizeof(Card), cudaMemcpyDeviceToHost);
    cudaFree(d_cards);
    if (rc != 0) { fail_nccl(h, c->nccl, rc); return bail(RRTMGP_B200_ERR_CUDA); }
    if (e != cudaSuccess) { fail_cuda(h, e); return bail(RRTMGP_B200_ERR_CUDA); }
    for (int p = 0; p < nranks; ++p) {
        if (cards[p].ncol != mine.ncol || cards[p].nlay != mine.nlay || cards[p].dtype != mine.dtype) {
            std::snprintf(h->cuda_err, sizeof(h->cuda_err), "rank %d has ncol/nlay/dtype %d/%d/%d, this rank %d/%d/%d", p, cards[p].ncol,
                          cards[p].nlay, cards[p].dtype, mine.ncol, mine.nlay, mine.dtype);
            return bail(RRTMGP_B200_ERR_INVALID_ARG);
        }
        if (p == rank) { c->peer[p] = c->gathered; continue; }
        void* mapped = nullptr;
        e = cudaIpcOpenMemHandle(&mapped, cards[p].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { fail_cuda(h, e); return bail(RRTMGP_B200_ERR_CUDA); }
        c->peer[p] = (unsigned char*)mapped;
    }
    c->ready = true;
    return RRTMGP_B200_OK;
}

int rrtmgp_b200_gathered_buffers(const rrtmgp_b200_handle_t* h, rrtmgp_b200_gathered_t* out) {
    if (!h || !out) return RRTMGP_B200_ERR_INVALID_ARG;
    if (!h->comm || !h->comm->gathered) return RRTMGP_B200_ERR_NOT_READY;
    void** o = reinterpret_cast<void**>(out);
    for (int v = 0; v < kGatheredViews; ++v) o[v] = h->comm->gathered + v * h->comm->view_bytes;
    return RRTMGP_B200_OK;
}

// one-element all-reduce on `s`: a rank passes it only when every rank has reached it (stream order)
static int comm_fence(rrtmgp_b200_handle* h, int which, cudaStream_t s) {
    CommState* c = h->comm;
    int rc = c->nccl.AllReduce(c->token + which, c->token + which, 1, kNcclFloat32, kNcclSum, c->comm, s);
    return rc == 0 ? RRTMGP_B200_OK : fail_nccl(h, c->nccl, rc);
}

// push rows [c0, c0 + count) of the local views [first, last) into every rank's gathered arrays (copy engines)
static cudaError_t push_views(rrtmgp_b200_handle* h, int first, int last, long long c0, int count) {
    CommState* c = h->comm;
    const size_t esz = h->cfg.dtype == 1 ? 8 : 4, row = (size_t)(h->cfg.nlay + 1) * esz;
    void* v[kGatheredViews];
    local_views(h, v);
    for (int p = 0; p < c->nranks; ++p) {
        const int peer = (c->rank + p) % c->nranks;      // own copy first, then a different first peer per rank
        for (int i = first; i < last; ++i) {
            if (!v[i]) continue;
            unsigned char* dst = c->peer[peer] + i * c->view_bytes + ((size_t)c->rank * h->cfg.ncol + c0) * row;
            const unsigned char* src = (const unsigned char*)v[i] + (size_t)c0 * row;
            cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)count * row, cudaMemcpyDeviceToDevice, c->push_stream(p % kCopyStreams));
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

int rrtmgp_b200_update_fluxes_gathered(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, void* stream) {
    int st = ready(h);
    if (st) return st;
    if (!h->comm) return RRTMGP_B200_ERR_NOT_READY;
    if (!h->buf.net_flux) return RRTMGP_B200_ERR_INVALID_ARG;
    DeviceGuard g(h->cfg.device);
    CommState* c = h->comm;
    cudaStream_t s = (cudaStream_t)stream;
    const bool f64 = h->cfg.dtype == 1;
    const unsigned long long sd = effective_seed(h, seed, have_seed);
    const int ncol = h->cfg.ncol;
    h->last_launches = 0;
    // nobody may still be reading the arrays of the previous step when the first push lands
    if ((st = comm_fence(h, 0, s))) return st;
    st = f64 ? prepare_t<double>(h, 0, ncol, s) : prepare_t<float>(h, 0, ncol, s);
    if (!st) st = f64 ? solve_lw_t<double>(h, sd, 0, ncol, s) : solve_lw_t<float>(h, sd, 0, ncol, s);
    if (st) return st;
    cudaError_t e = cudaEventRecord(c->ev_kernel[0], s);
    for (int j = 0; j < kCopyStreams && e == cudaSuccess; ++j) e = cudaStreamWaitEvent(c->push_stream(j), c->ev_kernel[0], 0);
    if (e == cudaSuccess) e = push_views(h, 0, 3, 0, ncol);                 // longwave views travel under the shortwave kernel
    if (e != cudaSuccess) return fail_cuda(h, e);
    // shortwave (+ net) in column chunks of whole waves of the persistent kernel; a chunk travels while the next one runs,
    // and the chunks shrink (45 / 30 / 17 / 8 %) because only the last one's pushes are exposed
    const int wave = 12 * (h->sm_count > 0 ? h->sm_count : 148);
    int nchunk = 1;
    if (!h->cfg.spectral_fluxes && ncol >= 4 * wave) nchunk = ncol >= 16 * wave ? 4 : 3;
    if (const char* env = std::getenv("RRTMGP_B200_GATHER_CHUNKS")) {   // experiments: 1, 3 (equal) or 4 (shrinking)
        const int n = std::atoi(env);
        if ((n == 1 || n == 3 || n == 4) && !h->cfg.spectral_fluxes && ncol >= 4 * wave) nchunk = n;
    }
    static const double kEnd4[4] = {0.45, 0.75, 0.92, 1.0};
    long long c0 = 0;
    for (int i = 0; i < nchunk; ++i) {
        const double end = nchunk == 4 ? kEnd4[i] : (double)(i + 1) / nchunk;
        long long c1 = i + 1 == nchunk ? ncol : ((long long)(ncol * end) / wave) * wave;
        if (c1 <= c0) continue;
        st = f64 ? solve_sw_t<double>(h, sd, true, c0, (int)(c1 - c0), s) : solve_sw_t<float>(h, sd, true, c0, (int)(c1 - c0), s);
        if (st) return st;
        e = cudaEventRecord(c->ev_kernel[1 + i], s);
        for (int j = 0; j < kCopyStreams && e == cudaSuccess; ++j) e = cudaStreamWaitEvent(c->push_stream(j), c->ev_kernel[1 + i], 0);
        if (e == cudaSuccess) e = push_views(h, 3, kGatheredViews, c0, (int)(c1 - c0));
        if (e != cudaSuccess) return fail_cuda(h, e);
        c0 = c1;
    }
    for (int j = 0; j < kCopyStreams && e == cudaSuccess; ++j) {
        e = cudaEventRecord(c->ev_copied[j], c->push_stream(j));
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s, c->ev_copied[j], 0);
    }
    if (e != cudaSuccess) return fail_cuda(h, e);
    return comm_fence(h, 1, s);                                             // everybody's pushes have landed
}

int rrtmgp_b200_all_gather_fluxes(rrtmgp_b200_handle_t* h, void* stream) {
    int st = ready(h);
    if (st) return st;
    if (!h->comm) return RRTMGP_B200_ERR_NOT_READY;
    DeviceGuard g(h->cfg.device);
    CommState* c = h->comm;
    void* v[kGatheredViews];
    local_views(h, v);
    const size_t n = (size_t)h->cfg.ncol * (h->cfg.nlay + 1);
    int rc = c->nccl.GroupStart();
    for (int i = 0; i < kGatheredViews && rc == 0; ++i)
        if (v[i]) rc = c->nccl.AllGather(v[i], c->gathered + i * c->view_bytes, n, h->cfg.dtype == 1 ? kNcclFloat64 : kNcclFloat32, c->comm,
                                         (cudaStream_t)stream);
    const int rc2 = c->nccl.GroupEnd();
    if (rc != 0 || rc2 != 0) return fail_nccl(h, c->nccl, rc != 0 ? rc : rc2);
    return RRTMGP_B200_OK;
}

int rrtmgp_b200_last_launch_count(const rrtmgp_b200_handle_t* h) { return h ? h->last_launches : 0; }
const char* rrtmgp_b200_last_cuda_error(const rrtmgp_b200_handle_t* h) { return h ? h->cuda_err : ""; }

}  // extern "C"

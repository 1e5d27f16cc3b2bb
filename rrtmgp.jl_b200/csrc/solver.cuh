// Fused column-radiation kernels: one warp owns one column, the 32 lanes are 32 consecutive
// g-points (two 16-g-point bands of the real tables), and the warp walks the column's
// vertical recurrences with the per-level state of its 32 (column, g-point) problems on chip.
// Replaces the reference's one-thread-per-column kernels
// (ext/cuda/rte_longwave_2stream.jl:86-141, rte_shortwave_2stream.jl:96-157,
// rte_longwave_noscat.jl:91-146), which round-trip tau/ssa/g, sources and per-g-point fluxes
// through global scratch for every g-point (src/rte/RTE.jl:121-128).
//
// Per column:
//   phase 0 (lane = layer)   everything that depends on (column, layer) only: T/p interpolation
//            indices and fractions, aerosol size bins and RH interval, cloud-LUT and Planck
//            interpolation positions.  The reference recomputes these for every g-point.
// and per block of 32 g-points:
//   phase 1 (lane = layer, loop over the block's bands)  everything that depends on (layer, band):
//            eta interpolation, minor-gas scalings, Planck functions, LUT cloud optics, MERRA
//            aerosol optics                                    -> shared-memory "band records"
//   McICA    (lane = g-point)  max-random cloud mask, one bit per layer        -> registers
//   phase 2  (lane = g-point)  LUT gathers (one 64 B segment per band per corner), tau/ssa/g,
//            two-stream coefficients, adding recurrences with 4-5 values per level per lane in
//            the on-chip "level store"
//   reduce   g-point sum into per-lane broadband accumulators (lane = level)
// and once per column the (nlev, ncol) epilogue with net flux, metric scaling, night zeroing,
// cloud cover and AOD diagnostics folded in.
#pragma once
#include "device_math.cuh"
#include "lut.cuh"

namespace rb {

enum SolveMode { MODE_LW_2STREAM = 0, MODE_LW_NOSCAT = 1, MODE_SW_2STREAM = 2 };

constexpr int kMaxLevPerLane = 3;   // nlev <= 96

template <typename FT>
struct ColumnIO {
    const FT* layerdata; const FT* p_lev; const FT* t_lev; const FT* t_sfc;
    const FT* vmr_h2o; const FT* vmr_o3; const FT* vmr;
    const FT *cld_r_eff_liq, *cld_r_eff_ice, *cld_path_liq, *cld_path_ice, *cld_frac;
    const FT *aero_mass, *aero_size;
    const FT *sfc_emis, *inc_flux_lw, *cos_zenith, *toa_flux, *sfc_alb_direct, *sfc_alb_diffuse, *metric_scaling;
    FT *out_up, *out_dn, *out_net, *out_dir;   // [ncol][nlev]
    const FT* add_net;                         // lw net to add (SW pass) or null
    FT* out_total_net;                         // net_flux = lw_net + sw_net or null
    FT* cld_cover;                             // [ncol] or null
    FT *aod_ext, *aod_sca;                     // [ncol] or null (SW)
    FT *band_up, *band_dn, *band_net;          // [n_bnd][ncol][nlev] or null
};

template <typename FT>
struct SolveParams {
    ColumnIO<FT> io;
    GasLut<FT> lut;
    CldLut<FT> cld;
    AeroLut<FT> aero;
    int ncol, nlay, ngas, vmr_kind, ice_rgh;
    int ncol_total;   // column stride of the [n_gpt][ncol] / [n_bnd][ncol][nlev] arrays (>= ncol for column ranges)
    int use_cloud, use_aero, n_mu;
    long long col_offset;
    unsigned long long seed;
    unsigned int* work_counter;   // fast kernels: next column to hand out (zeroed before the launch)
    // fast kernels, small shards: a column's g-point blocks split over `split` (1, 2 or 4) work items (solver_fast.cuh)
    int split;
    float* split_scratch;         // [ncol][split][3 * acc_stride + 4] partial sums, or null
    unsigned int* split_flags;    // [ncol] arrival counters, zero between launches
    FT Ds[4], wts[4];
    // per-warp shared-memory layout, in bytes from the warp's base
    int off_colj, off_colp, off_recj, off_rec, off_plk, off_store, warp_bytes;
    int rec_words;   // FT words per band record
    int rec_row;     // FT words per layer row of band records (>= maxb * rec_words; padded in the fast kernels)
};

// ---------------------------------------------------------------------------------------------
// VolumeMixingRatios.jl:91-129; ig 1-based gas index (0 = dry air)
// ---------------------------------------------------------------------------------------------
// `svmr`: optional on-chip copy of the global-mean array (fast kernels)
template <typename FT>
__device__ __forceinline__ FT get_vmr(const SolveParams<FT>& P, int ig, int lay, long long col, const FT* svmr = nullptr) {
    if (ig == 0) return FT(1);
    size_t k = (size_t)col * P.nlay + lay;
    if (P.vmr_kind == 0) {
        if (ig == 1) return __ldg(P.io.vmr_h2o + k);
        if (ig == 3) return __ldg(P.io.vmr_o3 + k);
        return svmr != nullptr ? svmr[ig - 1] : __ldg(P.io.vmr + ig - 1);
    }
    return __ldg(P.io.vmr + k * P.ngas + ig - 1);
}

template <typename FT> __device__ __forceinline__ FT int_as_ft(int i);
template <> __device__ __forceinline__ float int_as_ft<float>(int i) { return __int_as_float(i); }
template <> __device__ __forceinline__ double int_as_ft<double>(int i) { return (double)i; }

// cloud_optics.jl:154-192 / :207-244, split into "locate" (per layer) and "evaluate" (per band)
template <typename FT>
__device__ __forceinline__ void cld_locate(int nsize, FT lwr, FT upr, FT re, int& loc, FT& fac) {
    FT dr = pdiv(upr - lwr, FT(nsize - 1));
    re = rmax(rmin(re, upr), lwr);
    loc = (int)pdiv(re - lwr, dr) + 1;
    loc = loc < nsize - 1 ? loc : nsize - 1;
    loc = loc > 1 ? loc : 1;
    fac = pdiv(re - lwr - (loc - 1) * dr, dr);
}
template <bool S = false, typename FT>
__device__ __forceinline__ void cld_eval(int nsize, const FT* __restrict__ tbl, int loc, FT fac, FT path, FT& tau,
                                         FT& tau_ssa, FT& tau_ssag) {
    tau = tau_ssa = tau_ssag = FT(0);
    if (path > Num<FT>::eps()) {
        FT fc1 = FT(1) - fac;
        tau = rmax((fc1 * ldt<S>(tbl + loc - 1) + fac * ldt<S>(tbl + loc)) * path, FT(0));
        tau_ssa = (fc1 * ldt<S>(tbl + nsize + loc - 1) + fac * ldt<S>(tbl + nsize + loc)) * tau;
        tau_ssag = (fc1 * ldt<S>(tbl + 2 * nsize + loc - 1) + fac * ldt<S>(tbl + 2 * nsize + loc)) * tau_ssa;
    }
}

// aerosol_optics.jl:438-451
template <bool S = false, typename FT> __device__ __forceinline__ int merra_size_bin(const FT* __restrict__ lims, int nbins, FT size) {
    int bin = 1;
    for (int ib = 1; ib <= nbins; ++ib) {
        if (ldt<S>(lims + 2 * (ib - 1)) <= size && size <= ldt<S>(lims + 2 * (ib - 1) + 1)) { bin = ib; break; }
        bin = nbins;
    }
    return bin;
}

// Per-(column, layer) aerosol bookkeeping, hoisted out of the band loop (aerosol_optics.jl:141-235): `active` has
// one bit per species present (mass > 0), in the ORDER the reference sums them -- positions 0-4 dust (species
// 1, 8..11), 5-9 sea salt (2, 12..15), 10 sulfate, 11 BC-rh, 12 BC, 13 OC-rh, 14 OC -- `bins` the MERRA size bin
// of the ten sized species (3 bits per position) and `rh_loc` the RH interval.  The usual input has one species
// per layer, so the table row of the FIRST active species is resolved here once per layer: `off0` = element
// offset from AeroLut::dust of its (bin, RH interval) row for band 0, `bs0` = band stride | sea salt << 23 |
// rh-interpolated << 24 | 0-based species index << 25.
struct AeroLayer { unsigned active; unsigned bins; int rh_loc; int off0; int bs0; };

__device__ __forceinline__ int aero_pos_of_species(int i) {   // 0-based species -> summation position
    return i == 0 ? 0 : (i == 1 ? 5 : (i <= 6 ? i + 8 : (i <= 10 ? i - 6 : i - 5)));
}

// table row of the species at summation position `pos` (species functions aerosol_optics.jl:243-431)
template <typename FT>
__device__ __forceinline__ void aero_entry(const AeroLut<FT>& A, int pos, unsigned bins, int loc, int& off, int& bs) {
    const int nrh = A.nrh, nbin = A.nbin;
    if (pos < 5) {             // dust (3, nbin, nband)
        const int bin = (int)((bins >> (3 * pos)) & 7u);
        off = 3 * (bin - 1);
        bs = (3 * nbin) | ((pos == 0 ? 0 : 6 + pos) << 25);
    } else if (pos < 10) {     // sea salt (3, nrh, nbin, nband)
        const int k = pos - 5;
        const int bin = (int)((bins >> (3 * pos)) & 7u);
        off = (int)(A.sea_salt - A.dust) + 3 * nrh * (bin - 1) + 3 * (loc - 1);
        bs = (3 * nrh * nbin) | (1 << 23) | (1 << 24) | ((k == 0 ? 1 : 10 + k) << 25);
    } else {
        const FT* t = pos == 10 ? A.sulfate : (pos == 11 ? A.black_carbon_rh : (pos == 12 ? A.black_carbon
                                : (pos == 13 ? A.organic_carbon_rh : A.organic_carbon)));
        const bool rh = pos != 12 && pos != 14;
        off = (int)(t - A.dust) + (rh ? 3 * (loc - 1) : 0);
        bs = (rh ? (3 * nrh) | (1 << 24) : 3) | ((pos - 8) << 25);
    }
}

// aerosol_optics.jl:141-235; ibnd 0-based.  Dry species read one (ext, ssa, asy) row, RH-dependent ones
// interpolate between two adjacent rows; one branch-free form serves both (f_eff = 0, second row = first).
// `small` = where AeroLut::dust (and, at the same relative offsets, every aerosol table but sea salt) is read from:
// the global array, or the fast kernels' shared-memory copy (S = true; sea salt is always read from global memory).
template <bool S = false, typename FT>
__device__ __forceinline__ void lookup_aerosol(const AeroLut<FT>& A, const FT* small, int ibnd, const FT* __restrict__ mass,
                                               const AeroLayer& al, FT f, FT& tc, FT& tsc, FT& tsgc) {
    tc = tsc = tsgc = FT(0);
    auto species = [&](int off, int bs) {
        const FT* p = (((bs >> 23) & 1) ? A.dust : small) + (off + (bs & 0x7fffff) * ibnd);
        const bool rh = (bs >> 24) & 1;
        const int d = rh ? 3 : 0;
        const FT fe = rh ? f : FT(0);
        const FT m = __ldg(mass + (bs >> 25));
        FT t = m * (ldt<S>(p) * (FT(1) - fe) + ldt<S>(p + d) * fe);
        FT ts = t * (ldt<S>(p + 1) * (FT(1) - fe) + ldt<S>(p + d + 1) * fe);
        FT tsg = ts * (ldt<S>(p + 2) * (FT(1) - fe) + ldt<S>(p + d + 2) * fe);
        tc += t; tsc += ts; tsgc += tsg;
    };
    species(al.off0, al.bs0);
    unsigned rest = al.active & (al.active - 1u);
    while (rest) {   // layers with several species: resolve the remaining rows here
        const int pos = __ffs((int)rest) - 1;
        rest &= rest - 1u;
        int off, bs;
        aero_entry(A, pos, al.bins, al.rh_loc, off, bs);
        species(off, bs);
    }
}

// Small tables a fast kernel reads in phases 0 / 1: where each one lives (its copy inside the staged part of the
// small-table block, or the global array) is resolved ONCE per CTA into a shared-memory pointer table, so a use is
// one 64-bit shared load instead of the offset / range-check / select sequence (profiles/r2p: 50 of the 620
// instructions of a band pass).
enum SmallTable { TB_T_REF = 0, TB_LN_P_REF, TB_T_PLANCK, TB_TOT_PLANCK, TB_GPT2BND, TB_KEY_SPECIES, TB_VMR_REF,
                  TB_MINOR_BND_ST, TB_MINOR_BND_ST_UPPER, TB_MINOR_GASDATA, TB_MINOR_GASDATA_UPPER, TB_LIQDATA, TB_ICEDATA,
                  TB_AERO_DUST, TB_AERO_BIN_LIMS, TB_AERO_RH, TB_COUNT };

template <typename FT>
__device__ __forceinline__ void fill_small_table_pointers(const void** tp, const SolveParams<FT>& P, const unsigned char* sblob,
                                                          int staged, int tid) {
    if (tid >= TB_COUNT) return;
    const GasLut<FT>& L = P.lut;
    const void* g = nullptr;
    switch (tid) {
        case TB_T_REF: g = L.t_ref; break;
        case TB_LN_P_REF: g = L.ln_p_ref; break;
        case TB_T_PLANCK: g = L.t_planck; break;
        case TB_TOT_PLANCK: g = L.tot_planck; break;
        case TB_GPT2BND: g = L.gpt2bnd; break;
        case TB_KEY_SPECIES: g = L.key_species; break;
        case TB_VMR_REF: g = L.vmr_ref; break;
        case TB_MINOR_BND_ST: g = L.minor_bnd_st[0]; break;
        case TB_MINOR_BND_ST_UPPER: g = L.minor_bnd_st[1]; break;
        case TB_MINOR_GASDATA: g = L.minor_gasdata[0]; break;
        case TB_MINOR_GASDATA_UPPER: g = L.minor_gasdata[1]; break;
        case TB_LIQDATA: g = P.cld.liqdata; break;
        case TB_ICEDATA: g = P.cld.icedata; break;
        case TB_AERO_DUST: g = P.aero.dust; break;
        case TB_AERO_BIN_LIMS: g = P.aero.size_bin_limits; break;
        default: g = P.aero.rh_levels; break;
    }
    const long long off = reinterpret_cast<const unsigned char*>(g) - L.blob;
    tp[tid] = (g != nullptr && off >= 0 && off < staged) ? static_cast<const void*>(sblob + off) : g;
}

// ---------------------------------------------------------------------------------------------
// Per-warp context shared by the kernels below
// ---------------------------------------------------------------------------------------------
template <typename FT, int MODE, int NOWN, bool FUSED = false>
struct Warp {
    static constexpr bool LW = MODE != MODE_SW_2STREAM;
    static constexpr bool NOSCAT = MODE == MODE_LW_NOSCAT;

    const SolveParams<FT>& P;
    const GasLut<FT>& L;
    const int lane;
    const long long col;
    const int nlay, nlev;
    // shared memory
    int* colj;   // [nlay] jt | jp<<8 | tropo<<16 | aero<<17
    FT* colp;    // [nlay][4] ft, fp, col_dry, vmr_h2o + 1
    int* recj;   // [nlay][maxb] je1 | je2<<4 | nminor<<8
    FT* rec;     // [nlay][maxb][RW]: fe1, fe2, s1, s2, minor scalings, cloud(3), aerosol(3)
    FT* plk;     // LW: [maxb][2*nlev] B(t_lev) | B(t_lay) (noscat) | B(t_sfc)
    const int RW, maxb;
    bool rec_by_part = false;   // generic records built 32 layers at a time (solver_tm.cuh): row of layer k is k & 31
    // fast kernels: shared-memory copy of the small-table block (GasLut::blob) and of the global-mean vmr array
    const unsigned char* sblob;
    const int staged;   // bytes of the block that are staged (a prefix ending at a table boundary)
    const FT* svmr;     // element -1 holds 1 (dry air, gas index 0): svmr[ig - 1] serves every ig >= 0
    const void* const* tptr = nullptr;   // fast kernels: the CTA's small-table pointer table (fill_small_table_pointers)
    // per-lane registers for the layers this lane owns in phase 0/1: lane, lane+32, lane+64
    FT own_h2o[NOWN], own_g3[NOWN], own_dens[NOWN], own_cdry[NOWN];   // own_g3: vmr of gas 3 (ozone in VmrGM)
    AeroLayer own_aero[NOWN];
    FT own_rh_f[NOWN];
    int own_cld[NOWN];          // loc_liq | loc_ice << 8 | cloudy << 16
    FT own_cld_fl[NOWN], own_cld_fi[NOWN];
    int own_pl_loc[NOWN], own_py_loc[NOWN];   // Planck positions: t_lev[k+1], t_lay[k]
    FT own_pl_f[NOWN], own_py_f[NOWN];
    int p0_loc, psfc_loc; FT p0_f, psfc_f;                        // t_lev[0], t_sfc (lane 0)
    // current g-point block
    int gpt, ibnd, bl, b_first, nb;
    bool lane_on;
    unsigned mask[NOWN];
    // set once per column by kernels that scan cld_frac anyway: every cloud fraction of the column is exactly 0 or 1, so the
    // McICA sample is the same for every g-point -- mask = (cld_frac > 0), no draws (see mcica)
    bool cf_binary = false;
    unsigned cf_words[NOWN];

    __device__ __forceinline__ Warp(const SolveParams<FT>& P_, unsigned char* wbase, int lane_, long long col_,
                                    const unsigned char* sblob_ = nullptr, int staged_ = 0, const FT* svmr_ = nullptr)
        : P(P_), L(P_.lut), lane(lane_), col(col_), nlay(P_.nlay), nlev(P_.nlay + 1), sblob(sblob_), staged(staged_), svmr(svmr_),
          colj(reinterpret_cast<int*>(wbase + P_.off_colj)), colp(reinterpret_cast<FT*>(wbase + P_.off_colp)),
          recj(reinterpret_cast<int*>(wbase + P_.off_recj)), rec(reinterpret_cast<FT*>(wbase + P_.off_rec)),
          plk(reinterpret_cast<FT*>(wbase + P_.off_plk)), RW(P_.rec_words), maxb(P_.lut.maxb) {}

    // words per band of the Planck buffer
    __device__ __forceinline__ int plk_stride() const { return FUSED ? (NOSCAT ? 2 * nlev : nlev + 1) : 2 * nlev; }

    // A small table: the global array, or (fast kernels) wherever the CTA's pointer table says it lives -- its copy
    // inside the staged part of the block, or the global array.  `sel` (0 / 1) picks the upper-atmosphere twin.
    template <class T> __device__ __forceinline__ const T* tb(const T* g, int idx, int sel = 0) const {
        if (FUSED) return reinterpret_cast<const T*>(tptr[idx + sel]);
        return g;
    }
    // vmr of gas ig at layer k = lane + 32 j.  Fast kernels: the two gases that vary per layer in VmrGM (1 = h2o,
    // 3 = o3, VolumeMixingRatios.jl:91-106) come from the registers phase 0 filled and the global means from the
    // staged copy, so the dependent chains of phase 1 (key species -> gas -> vmr) never wait on global memory.
    __device__ __forceinline__ FT vmr_of(int ig, int k, int j) const {
        if (FUSED) {
            if (ig == 1 && L.idx_h2o == 1) return pick(own_h2o, j);
            if (ig == 3) return pick(own_g3, j);
        }
        return get_vmr(P, ig, k, col, FUSED ? svmr : nullptr);
    }
    // own_x[j] for a warp-uniform RUNTIME j as selects (a dynamically indexed register array would go to local
    // memory); with a compile-time j (generic kernel: unrolled loop) it folds to the element
    template <class T> __device__ __forceinline__ T pick(const T (&a)[NOWN], int j) const {
        T r = a[0];
        if (NOWN > 1 && j == 1) r = a[1];
        if (NOWN > 2 && j == 2) r = a[2];
        return r;
    }

    // ---------------- phase 0 (gas_optics.jl:87-115,188 and the hoisted per-layer searches) ----------------
    __device__ __forceinline__ void phase0() {
        const FT* ld = P.io.layerdata + (size_t)col * nlay * 4;
        const bool use_cloud = P.use_cloud != 0, use_aero = P.use_aero != 0;
        const int n_t = L.n_t;
        const FT* t_ref = tb(L.t_ref, TB_T_REF);
        const FT* ln_p_ref = tb(L.ln_p_ref, TB_LN_P_REF);
        const FT* t_planck = tb(L.t_planck, TB_T_PLANCK);
#pragma unroll
        for (int j = 0; j < NOWN; ++j) {
            const int k = lane + 32 * j;
            own_h2o[j] = FT(0); own_g3[j] = FT(0); own_dens[j] = FT(0); own_cdry[j] = FT(0); own_aero[j] = AeroLayer{0u, 0u, 1, 0, 0}; own_rh_f[j] = FT(0);
            own_cld[j] = 0; own_cld_fl[j] = own_cld_fi[j] = FT(0);
            own_pl_loc[j] = own_py_loc[j] = 0; own_pl_f[j] = own_py_f[j] = FT(0);
            if (k >= nlay) continue;
            FT col_dry = __ldg(ld + 4 * k + 0), p_lay = __ldg(ld + 4 * k + 1), t_lay = __ldg(ld + 4 * k + 2);
            int tropo = p_lay > L.p_ref_tropo ? 1 : 2;
            FT dT = ldt<FUSED>(t_ref + 1) - ldt<FUSED>(t_ref);
            int jt = loc_lower_eq<FUSED>(t_lay, dT, n_t, t_ref);
            FT ft = pdiv(t_lay - ldt<FUSED>(t_ref + jt - 1), dT);
            FT dlnp = ldt<FUSED>(ln_p_ref) - ldt<FUSED>(ln_p_ref + 1);
            FT lp = rlog(p_lay);
            int jpress = (int)pdiv(ldt<FUSED>(ln_p_ref) - lp, dlnp) + 1;
            jpress = jpress > 1 ? jpress : 1;
            jpress = (jpress < L.n_p_ref - 1 ? jpress : L.n_p_ref - 1) + 1;
            FT fp = pdiv(ldt<FUSED>(ln_p_ref + jpress - 2) - lp, dlnp);
            int jp = jpress + tropo - 1;
            FT h2o = get_vmr(P, L.idx_h2o, k, col, FUSED ? svmr : nullptr);
            own_h2o[j] = h2o;
            if (FUSED && P.ngas >= 3) own_g3[j] = get_vmr(P, 3, k, col);
            own_cdry[j] = col_dry;
            own_dens[j] = pdiv(FT(0.01) * p_lay, t_lay);
            int aero_on = 0;
            if (use_aero) {   // aerosol_optics.jl:464-483, :438-451, optics_utils.jl:51-62
                const FT* am = P.io.aero_mass + ((size_t)col * nlay + k) * 15;
                const FT* as = P.io.aero_size + ((size_t)col * nlay + k) * 15;
                unsigned act = 0u, bins = 0u;
                for (int i = 0; i < 15; ++i) act |= (__ldg(am + i) > FT(0)) ? (1u << aero_pos_of_species(i)) : 0u;
                if (act) {
                    unsigned sized = act & 0x3ffu;   // sized species: dust 1, 8..11 (positions 0-4) then sea salt 2, 12..15
                    while (sized) {                  // one pass per species present, not per species slot
                        const int pos = __ffs((int)sized) - 1;
                        sized &= sized - 1u;
                        const int i = pos < 5 ? (pos == 0 ? 0 : 6 + pos) : (pos == 5 ? 1 : 5 + pos);
                        bins |= (unsigned)merra_size_bin<FUSED>(tb(P.aero.size_bin_limits, TB_AERO_BIN_LIMS), P.aero.nbin, __ldg(as + i)) << (3 * pos);
                    }
                    int loc; FT f;
                    interp1d_loc_factor<FUSED>(__ldg(ld + 4 * k + 3), tb(P.aero.rh_levels, TB_AERO_RH), P.aero.nrh, loc, f);
                    int off0, bs0;
                    aero_entry(P.aero, __ffs((int)act) - 1, bins, loc, off0, bs0);
                    own_aero[j] = AeroLayer{act, bins, loc, off0, bs0};
                    own_rh_f[j] = f;
                    aero_on = 1;
                }
            }
            if (use_cloud) {
                size_t kk = (size_t)col * nlay + k;
                if (__ldg(P.io.cld_frac + kk) > FT(0)) {
                    const CldLut<FT>& C = P.cld;
                    int ll, li;
                    cld_locate(C.nsize_liq, C.radliq_lwr, C.radliq_upr, __ldg(P.io.cld_r_eff_liq + kk), ll, own_cld_fl[j]);
                    cld_locate(C.nsize_ice, C.radice_lwr, C.radice_upr, __ldg(P.io.cld_r_eff_ice + kk), li, own_cld_fi[j]);
                    own_cld[j] = ll | (li << 8) | (1 << 16);
                }
            }
            if (LW) {
                const FT* tl = P.io.t_lev + (size_t)col * nlev;
                interp1d_eq_locate<FUSED>(__ldg(tl + k + 1), t_planck, L.n_t_plnk, own_pl_loc[j], own_pl_f[j]);
                if (NOSCAT) interp1d_eq_locate<FUSED>(t_lay, t_planck, L.n_t_plnk, own_py_loc[j], own_py_f[j]);
            }
            colj[k] = jt | (jp << 8) | ((tropo - 1) << 16) | (aero_on << 17);
            colp[4 * k + 0] = ft; colp[4 * k + 1] = fp;
            if (FUSED) {
                // fast kernels: element offsets of the (jt) node row of the packed minor table (float4 units, both
                // tropospheres in one arena) and of the (jp-1, jt) node row of the major table
                const int tr_off = tropo == 2 ? (int)((L.kminor4[1] - L.kminor4[0]) >> 2) : 0;
                colp[4 * k + 2] = int_as_ft<FT>(tr_off + (jt - 1) * L.n_eta * L.n_gpt);
                colp[4 * k + 3] = int_as_ft<FT>(((jp - 2) * n_t + (jt - 1)) * L.n_eta * L.n_gpt);
            } else {
                colp[4 * k + 2] = col_dry;
                colp[4 * k + 3] = h2o + FT(1);
            }
        }
        p0_loc = psfc_loc = 0; p0_f = psfc_f = FT(0);
        if (LW && lane == 0) {
            interp1d_eq_locate<FUSED>(__ldg(P.io.t_lev + (size_t)col * nlev), t_planck, L.n_t_plnk, p0_loc, p0_f);
            interp1d_eq_locate<FUSED>(__ldg(P.io.t_sfc + col), t_planck, L.n_t_plnk, psfc_loc, psfc_f);
        }
        __syncwarp();
    }

    __device__ __forceinline__ void set_block(int g0) {
        const int n_gpt = L.n_gpt;
        lane_on = g0 + lane < n_gpt;
        gpt = lane_on ? g0 + lane : n_gpt - 1;
        const int* gpt2bnd = tb(L.gpt2bnd, TB_GPT2BND);
        b_first = ldt<FUSED>(gpt2bnd + g0);
        const int b_last = ldt<FUSED>(gpt2bnd + (g0 + 31 < n_gpt ? g0 + 31 : n_gpt - 1));
        nb = b_last - b_first + 1;
        ibnd = ldt<FUSED>(gpt2bnd + gpt);
        bl = ibnd - b_first;
    }

    // ---------------- phase 1: band records of this block ----------------
    // returns this lane's partial (aod_ext, aod_sca) of the 550 nm band if it is in the block
    // `half` >= 0 (fast kernels): only the layers [32 half, 32 half + 32) and records indexed by (k & 31),
    // so the band records of half a column fit in 3.6 KB; `half` < 0: all layers
    __device__ __forceinline__ void phase1(FT& aod_e, FT& aod_s, int half = -1) {
        const bool use_cloud = P.use_cloud != 0, use_aero = P.use_aero != 0;
        const int n_eta = L.n_eta;
        aod_e = aod_s = FT(0);
        // Fast kernels (half >= 0): ONE copy of the body with j = half selected at run time -- unrolled over j the
        // body existed NOWN times (~15 KB each), and phase 1 + the main loop no longer fit the instruction cache
        // (profiles/r1r: no_inst 10-12 %).  Generic kernel (half < 0): all layers, unrolled.
#pragma unroll
        for (int jj = 0; jj < (FUSED ? 1 : NOWN); ++jj) {
            const int j = FUSED ? half : jj;
            const int k = lane + 32 * j;
            if (k >= nlay || (!FUSED && half >= 0 && j != half)) continue;
            const int kr = half >= 0 ? lane : k;   // record row
            const int cj = colj[k];
            const int jt = cj & 0xff, tropo = ((cj >> 16) & 1) + 1;
            const FT col_dry = pick(own_cdry, j);
            const FT vmr_h2o = pick(own_h2o, j);
            const FT dens_j = pick(own_dens, j);
            const int cld_j = pick(own_cld, j);
            const FT cld_fl_j = pick(own_cld_fl, j), cld_fi_j = pick(own_cld_fi, j);
            const AeroLayer aero_j = pick(own_aero, j);
            const FT rh_f_j = pick(own_rh_f, j);
            const int pl_loc_j = pick(own_pl_loc, j), py_loc_j = pick(own_py_loc, j);
            const FT pl_f_j = pick(own_pl_f, j), py_f_j = pick(own_py_f, j);
            const FT dry_fact = hdiv(FT(1), FT(1) + vmr_h2o);
            // vmr of gas ig, branch-free for the fast kernels with VmrGM storage (the gas index depends on (tropo, band), so
            // lanes disagree and the if-chain of vmr_of diverged: 6 % of the longwave kernel's time in profiles/r2h)
            const bool gm_fast = FUSED && P.vmr_kind == 0 && svmr != nullptr && L.idx_h2o == 1 && P.ngas >= 3;   // warp-uniform
            const FT g3_j = pick(own_g3, j);
            auto vmrq = [&](int ig) -> FT {
                if (gm_fast) {
                    FT v = svmr[ig - 1];              // svmr[-1] = 1: dry air
                    v = ig == 1 ? vmr_h2o : v;
                    v = ig == 3 ? g3_j : v;
                    return v;
                }
                return vmr_of(ig, k, j);
            };
            for (int b = 0; b < nb; ++b) {
                const int ib = b_first + b;
                FT* r = rec + (kr * P.rec_row + b * RW);
                // gas_optics.jl:129-170
                const int* ksp = tb(L.key_species, TB_KEY_SPECIES) + 2 * ((tropo - 1) + 2 * ib);
                const int ig1 = ldt<FUSED>(ksp), ig2 = ldt<FUSED>(ksp + 1);
                const FT vmr1 = vmrq(ig1), vmr2 = vmrq(ig2);
                int je[2];
                FT fe[2], smix[2];
                // fast kernels (FUSED): record = 8 corner weights | s1, s2, major-table offsets of the two T nodes |
                // slot scalings | {aerosol-only products, minor-table offset} | {cloud+aerosol products, same offset}
                const int sc0 = FUSED ? 12 : 4;
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const FT* vr = tb(L.vmr_ref, TB_VMR_REF) + (2 * L.ngas1 * (jt - 1 + it) + (tropo - 1));
                    FT eta_half = hdiv(ldt<FUSED>(vr + 2 * ig1), ldt<FUSED>(vr + 2 * ig2));
                    FT col_mix = vmr1 + eta_half * vmr2;
                    FT eta = vmr1 * hdiv(FT(1), col_mix);
                    if (col_mix <= FT(0)) eta = FT(0.5);
                    FT loc_eta = eta * FT(n_eta - 1);
                    int jj = (int)loc_eta + 1;
                    jj = jj < n_eta - 1 ? jj : n_eta - 1;
                    je[it] = jj; fe[it] = loc_eta - FT(jj - 1);
                    smix[it] = FUSED ? col_mix * col_dry : col_mix;   // fast kernels fold col_dry in here
                    if (!FUSED) { r[it] = fe[it]; r[2 + it] = smix[it]; }
                }
                // gas_optics.jl:344-412: per-absorber scalings.  Fast kernels (FUSED) gather four slots per
                // 128-bit load: slots are zero padded to a multiple of four and SW slot 0 is Rayleigh
                // (gas_optics.jl:430-444: (vmr_h2o + 1) * col_dry).
                const int soff = (FUSED && !LW) ? 1 : 0;
                const int nslots = FUSED ? 4 * L.n_minor_groups : L.nminor_max;
                const int* bst = tb(L.minor_bnd_st[tropo - 1], TB_MINOR_BND_ST, tropo - 1);
                const int4* gdt = reinterpret_cast<const int4*>(tb(L.minor_gasdata[tropo - 1], TB_MINOR_GASDATA, tropo - 1));
                const int m0 = ldt<FUSED>(bst + ib), nmin = ldt<FUSED>(bst + ib + 1) - m0;
                auto minor_scaling = [&](int i) -> FT {      // gas_optics.jl:344-412, absorber i of this (band, tropo)
                    const int4 gd = ldt<FUSED>(gdt + (m0 + i));
                    const FT vmr_i = vmrq(gd.x);
                    FT scaling = FT(0);
                    if (vmr_i > FT(0)) {
                        scaling = vmr_i * col_dry;
                        if (gd.z == 1) {
                            scaling *= dens_j;
                            if (gd.y > 0) {
                                if (gd.w == 1) scaling *= (FT(1) - vmrq(gd.y) * dry_fact);
                                else scaling *= vmrq(gd.y) * dry_fact;
                            }
                        }
                    }
                    return scaling;
                };
                if constexpr (FUSED) {   // unused slots are zero: clear the groups first (128-bit stores), then fill
                    reinterpret_cast<float4*>(r + sc0)[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (L.n_minor_groups > 1) reinterpret_cast<float4*>(r + sc0)[1] = make_float4(0.f, 0.f, 0.f, 0.f);   // (fast kernels: one or two groups)
                    if (!LW) r[sc0] = (vmr_h2o + FT(1)) * col_dry;
                }
                for (int i = 0; i < nmin; ++i) r[sc0 + soff + i] = minor_scaling(i);
                FT* rc = r + sc0 + nslots;   // cloud (3) then aerosol (3)  [FUSED: aerosol-only / cloud+aerosol products]
                FT tc = FT(0), sc = FT(0), gc = FT(0);
                // cloud_optics.jl:70-138 (2-stream) / :1-50 (1-scalar), for layers that can be cloudy
                if (use_cloud) {
                    if ((cld_j >> 16) & 1) {
                        const CldLut<FT>& C = P.cld;
                        size_t kk = (size_t)col * nlay + k;
                        const FT* liq = tb(C.liqdata, TB_LIQDATA) + 3 * C.nsize_liq * ib;
                        const FT* ice = tb(C.icedata, TB_ICEDATA) + 3 * C.nsize_ice * (ib + C.nband * (P.ice_rgh - 1));
                        FT tl, tls, tlsg, ti, tis, tisg;
                        cld_eval<FUSED>(C.nsize_liq, liq, cld_j & 0xff, cld_fl_j, __ldg(P.io.cld_path_liq + kk), tl, tls, tlsg);
                        cld_eval<FUSED>(C.nsize_ice, ice, (cld_j >> 8) & 0xff, cld_fi_j, __ldg(P.io.cld_path_ice + kk), ti, tis, tisg);
                        if (NOSCAT) {
                            tc = (tl - tls) + (ti - tis);
                        } else {
                            tc = tl + ti;
                            sc = tls + tis;
                            gc = hdiv(tlsg + tisg, rmax(Num<FT>::eps(), sc));
                            sc = hdiv(sc, rmax(Num<FT>::eps(), tc));
                            if (!LW) delta_scale(tc, sc, gc);
                        }
                    }
                }
                // aerosol_optics.jl:80-133 (2-stream) / :18-61 (1-scalar)
                FT ta = FT(0), sa = FT(0), ga = FT(0);
                if (use_aero) {
                    if ((cj >> 17) & 1) {
                        size_t kk = ((size_t)col * nlay + k) * 15;
                        FT tsa, tsga;
                        lookup_aerosol<FUSED>(P.aero, tb(P.aero.dust, TB_AERO_DUST), ib, P.io.aero_mass + kk, aero_j, rh_f_j, ta, tsa, tsga);
                        if (!LW && ib + 1 == P.aero.iband_550nm) { aod_e += ta; aod_s += tsa; }   // :96-116
                        if (NOSCAT) {
                            ta = ta - tsa;
                        } else {
                            ga = hdiv(tsga, rmax(Num<FT>::eps(), tsa));
                            sa = hdiv(tsa, rmax(Num<FT>::eps(), ta));
                            if (!LW) delta_scale(ta, sa, ga);
                        }
                    }
                }
                if constexpr (FUSED) {
                    // increment_2stream (optics_utils.jl:189-202) is additive in (tau, tau ssa, tau ssa g):
                    // store those products for "aerosol only" and "cloud + aerosol" as two 16-byte groups whose
                    // fourth word is the minor-table offset, so a cell needs one 128-bit load for both
                    const FT a0 = ta, a1 = ta * sa, a2 = ta * sa * ga;
                    // trilinear corner weights (optics_utils.jl:136-181): index = T node * 4 + p node * 2 + eta node,
                    // eta cell and fraction per T node; the cell sums weight * corner and scales by s1 / s2
                    const FT ft = colp[4 * k + 0], fp = colp[4 * k + 1];
                    const FT omft = FT(1) - ft, omfp = FT(1) - fp;
                    const FT wa0 = omfp * omft, wa1 = fp * omft, wb0 = omfp * ft, wb1 = fp * ft;
                    // (128-bit stores: with lane = layer and a row of 4 * odd words they are bank-conflict free)
                    float4* r4 = reinterpret_cast<float4*>(r);
                    r4[0] = make_float4(wa0 * (FT(1) - fe[0]), wa0 * fe[0], wa1 * (FT(1) - fe[0]), wa1 * fe[0]);
                    r4[1] = make_float4(wb0 * (FT(1) - fe[1]), wb0 * fe[1], wb1 * (FT(1) - fe[1]), wb1 * fe[1]);
                    // element offsets without the g-point: (jp-1, jt, je1) and (jp-1, jt+1, je2) rows of the major
                    // table, (jt, je1) row of the packed minor table; its (jt+1, je2) row is at ma + (ib - ia)
                    const int e1 = (je[0] - 1) * L.n_gpt, e2 = (je[1] - 1) * L.n_gpt;
                    const int rowoff = __float_as_int((float)colp[4 * k + 3]), moff = __float_as_int((float)colp[4 * k + 2]);
                    r4[2] = make_float4(smix[0], smix[1], int_as_ft<FT>(rowoff + e1), int_as_ft<FT>(rowoff + L.n_eta * L.n_gpt + e2));
                    const FT ma = int_as_ft<FT>(moff + e1);
                    float4* rc4 = reinterpret_cast<float4*>(rc);
                    rc4[0] = make_float4(a0, a1, a2, ma);
                    rc4[1] = make_float4(a0 + tc, a1 + tc * sc, a2 + tc * sc * gc, ma);
                } else {
                    if (use_cloud) { rc[0] = tc; rc[1] = sc; rc[2] = gc; }
                    if (use_aero) { rc[3] = ta; rc[4] = sa; rc[5] = ga; }
                    recj[kr * maxb + b] = je[0] | (je[1] << 4) | (nmin << 8);
                }
                // Planck functions of this band (compute_optical_props.jl:157-195 / :43-82)
                if (LW) {
                    const FT* totplnk = tb(L.tot_planck, TB_TOT_PLANCK) + L.n_t_plnk * ib;
                    // per band: B(t_lev[0..nlay]), then B(t_lay) (no-scattering only), B(t_sfc) last;
                    // the fast kernels keep just the nlev + 1 values they use
                    // (fast no-scattering kernel: B(t_lev) [nlev], B(t_sfc), then B(t_lay) [nlay])
                    FT* pb = plk + b * plk_stride();
                    pb[k + 1] = interp1d_eq_eval<FUSED>(pl_loc_j, pl_f_j, totplnk, L.n_t_plnk);
                    if (k == 0) {
                        pb[0] = interp1d_eq_eval<FUSED>(p0_loc, p0_f, totplnk, L.n_t_plnk);
                        pb[FUSED ? nlev : nlev + nlay] = interp1d_eq_eval<FUSED>(psfc_loc, psfc_f, totplnk, L.n_t_plnk);
                    }
                    if (NOSCAT) pb[(FUSED ? nlev + 1 : nlev) + k] = interp1d_eq_eval<FUSED>(py_loc_j, py_f_j, totplnk, L.n_t_plnk);
                }
            }
        }
        __syncwarp();
    }

    // ---------------- McICA mask of this lane's g-point (cloud_optics.jl:264-307) ----------------
    // returns the number of this block's g-points with any cloudy layer
    __device__ __forceinline__ int mcica(uint64_t col_key, int cld_start, int cld_finish) {
#pragma unroll
        for (int i = 0; i < NOWN; ++i) mask[i] = 0u;
        if (!(P.use_cloud != 0) || cld_finish <= 0) return 0;
        if (cf_binary) {
            // cloud_optics.jl:276-301 with every fraction 0 or 1: the threshold 1 - cf of a cloudy layer is exactly 0, so
            // `r >= 1 - cf` holds for any draw r in [0, 1) (fresh, reused or rescaled) and a layer with cf = 0 is clear:
            // bit-identical to the loop below, without the draws
#pragma unroll
            for (int i = 0; i < NOWN; ++i) mask[i] = cf_words[i];
            return __popc(__ballot_sync(0xffffffffu, lane_on));
        }
        const FT* cf = P.io.cld_frac + (size_t)col * nlay;
        const int swflag = LW ? 0 : 1;
        FT cf_p1 = __ldg(cf + cld_finish - 1);
        double r_p1 = mcica_rand(col_key, swflag, gpt + 1, cld_finish);
        bool m_p1 = r_p1 >= (double)(FT(1) - cf_p1);
        set_mask(cld_finish - 1, m_p1);
        for (int ilay = cld_finish - 1; ilay >= cld_start; --ilay) {
            FT cfk = __ldg(cf + ilay - 1);
            bool m = false;
            if (cfk > FT(0)) {
                double r = m_p1 ? r_p1 : mcica_rand(col_key, swflag, gpt + 1, ilay) * (double)(FT(1) - cf_p1);
                m = r >= (double)(FT(1) - cfk);
                r_p1 = r;
            }
            set_mask(ilay - 1, m);
            cf_p1 = cfk; m_p1 = m;
        }
        unsigned anyw = 0u;
#pragma unroll
        for (int i = 0; i < NOWN; ++i) anyw |= mask[i];
        bool any = lane_on && anyw != 0u;
        return __popc(__ballot_sync(0xffffffffu, any));
    }
    __device__ __forceinline__ void set_mask(int k, bool m) {
        const unsigned bit = m ? (1u << (k & 31)) : 0u;
        if (k < 32) mask[0] |= bit; else if (NOWN < 3 || k < 64) mask[NOWN > 1 ? 1 : 0] |= bit; else mask[NOWN - 1] |= bit;
    }
    __device__ __forceinline__ bool mask_bit(int k) const {
        unsigned w = k < 32 ? mask[0] : ((NOWN < 3 || k < 64) ? mask[NOWN > 1 ? 1 : 0] : mask[NOWN - 1]);
        return (w >> (k & 31)) & 1u;
    }

    // ---------------- gas + cloud + aerosol optics of layer k for this lane's g-point ----------------
    // gas_optics.jl:176-320 with the (layer, band) work read from the band record; 32-bit table offsets
    __device__ __forceinline__ void optics(int k, FT& tau, FT& ssa, FT& g, FT& pfrac) const {
        const int n_gpt = L.n_gpt, n_eta = L.n_eta, n_t = L.n_t;
        const int cj = colj[k];
        const int jt = cj & 0xff, jp = (cj >> 8) & 0xff, tr = (cj >> 16) & 1;
        const FT ft = colp[4 * k + 0], fp = colp[4 * k + 1], col_dry = colp[4 * k + 2];
        const int kr = rec_by_part ? (k & 31) : k;   // record row
        const int rj = recj[kr * maxb + bl];
        const int je1 = rj & 0xf, je2 = (rj >> 4) & 0xf, nmin = rj >> 8;
        const FT* r = rec + kr * P.rec_row + bl * RW;
        const FT fe1 = r[0], fe2 = r[1];
        const FT omfe1 = FT(1) - fe1, omfe2 = FT(1) - fe2, omft = FT(1) - ft, omfp = FT(1) - fp;
        // element offsets (all tables < 2^31 elements)
        const int st = n_eta * n_gpt, sp = st * n_t;
        const int oa = (jp - 2) * sp + (jt - 1) * st + (je1 - 1) * n_gpt + gpt;    // (jp-1, jt,   je1)
        const int ob = (jp - 2) * sp + jt * st + (je2 - 1) * n_gpt + gpt;          // (jp-1, jt+1, je2)
        const int ma = (jt - 1) * st + (je1 - 1) * n_gpt + gpt;                    // minor / Rayleigh (jt, je1)
        const int mb = jt * st + (je2 - 1) * n_gpt + gpt;                          //                  (jt+1, je2)
        {
            const FT* km = L.kmajor;
            FT c000 = __ldg(km + oa), c100 = __ldg(km + oa + n_gpt), c010 = __ldg(km + oa + sp), c110 = __ldg(km + oa + sp + n_gpt);
            FT c001 = __ldg(km + ob), c101 = __ldg(km + ob + n_gpt), c011 = __ldg(km + ob + sp), c111 = __ldg(km + ob + sp + n_gpt);
            FT tau_major = (r[2] * (omfp * (omft * (omfe1 * c000 + fe1 * c100)) + fp * (omft * (omfe1 * c010 + fe1 * c110))) +
                            r[3] * (omfp * (ft * (omfe2 * c001 + fe2 * c101)) + fp * (ft * (omfe2 * c011 + fe2 * c111)))) * col_dry;
            tau = tau_major;
        }
        const FT w11 = omfe1 * omft, w21 = fe1 * omft, w12 = omfe2 * ft, w22 = fe2 * ft;   // optics_utils.jl:85-98
        {
            const FT* kmn = L.kminor[tr];
            FT tau_minor = FT(0);
            const int slot = n_t * st;
            for (int i = 0; i < nmin; ++i) {
                const FT* t = kmn + i * slot;
                FT v = w11 * __ldg(t + ma) + w21 * __ldg(t + ma + n_gpt) + w12 * __ldg(t + mb) + w22 * __ldg(t + mb + n_gpt);
                tau_minor += v * r[4 + i];
            }
            tau += tau_minor;
        }
        if (LW) {
            const FT* pf = L.pfrac;
            FT c000 = __ldg(pf + oa), c100 = __ldg(pf + oa + n_gpt), c010 = __ldg(pf + oa + sp), c110 = __ldg(pf + oa + sp + n_gpt);
            FT c001 = __ldg(pf + ob), c101 = __ldg(pf + ob + n_gpt), c011 = __ldg(pf + ob + sp), c111 = __ldg(pf + ob + sp + n_gpt);
            pfrac = (omfp * (omft * (omfe1 * c000 + fe1 * c100)) + fp * (omft * (omfe1 * c010 + fe1 * c110))) +
                    (omfp * (ft * (omfe2 * c001 + fe2 * c101)) + fp * (ft * (omfe2 * c011 + fe2 * c111)));
            tau = rmax(tau, FT(0));
            ssa = FT(0); g = FT(0);
        } else {
            const FT* t = L.rayl + tr * (n_t * st);
            FT tau_ray = (w11 * __ldg(t + ma) + w21 * __ldg(t + ma + n_gpt) + w12 * __ldg(t + mb) + w22 * __ldg(t + mb + n_gpt)) *
                         colp[4 * k + 3] * col_dry;
            tau = rmax(tau + tau_ray, FT(0));
            ssa = tau > FT(0) ? hdiv(tau_ray, tau) : FT(0);
            g = FT(0); pfrac = FT(0);
        }
        const FT* rc = r + 4 + L.nminor_max;
        if (P.use_cloud != 0 && mask_bit(k)) {
            if (NOSCAT) tau += rc[0];
            else increment_2stream(tau, ssa, g, rc[0], rc[1], rc[2]);
        }
        if (P.use_aero != 0 && ((cj >> 17) & 1)) {
            if (NOSCAT) tau += rc[3];
            else increment_2stream(tau, ssa, g, rc[3], rc[4], rc[5]);
        }
    }
};

// warp-wide sum, every lane gets the total
template <typename FT> __device__ __forceinline__ FT warp_sum(FT v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// kernel with the level store in shared memory (any precision, nlev <= 96)
// ---------------------------------------------------------------------------------------------
template <typename FT, int MODE>
__global__ void __launch_bounds__(256) solve_kernel(const SolveParams<FT> P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    // warp index through a shuffle: provably warp-uniform, so everything derived from it (column, shared
    // memory bases, table descriptors) lives in uniform registers instead of being re-broadcast with R2UR
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int wpc = blockDim.x >> 5;
    const long long col = (long long)blockIdx.x * wpc + warp;
    if (col >= P.ncol) return;   // warp-uniform; no block-level barriers are used below

    constexpr bool LW = MODE != MODE_SW_2STREAM;
    constexpr bool NOSCAT = MODE == MODE_LW_NOSCAT;
    constexpr int NV = MODE == MODE_LW_2STREAM ? 4 : 5;   // level-store values per level

    unsigned char* wbase = smem_raw + (size_t)warp * P.warp_bytes;
    Warp<FT, MODE, kMaxLevPerLane> W(P, wbase, lane, col);
    const GasLut<FT>& L = P.lut;
    const int nlay = P.nlay, nlev = nlay + 1, n_gpt = L.n_gpt;
    FT* store = reinterpret_cast<FT*>(wbase + P.off_store);   // [nlev][NV][32]
    const bool use_cloud = P.use_cloud != 0;

    W.phase0();

    // McICA column key; cloudy span (cloud_optics.jl:271-275,309-321)
    const uint64_t col_key = mcica_col_key(P.seed, (uint64_t)(P.col_offset + col));
    int cld_start = 0, cld_finish = 0;   // 1-based, 0 = no cloud
    if (use_cloud) {
        const FT* cf = P.io.cld_frac + (size_t)col * nlay;
        unsigned lo = 0xffffffffu, hi = 0;
        for (int k = lane; k < nlay; k += 32)
            if (__ldg(cf + k) > FT(0)) { lo = lo < (unsigned)(k + 1) ? lo : (unsigned)(k + 1); hi = hi > (unsigned)(k + 1) ? hi : (unsigned)(k + 1); }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (hi > 0) { cld_start = (int)lo; cld_finish = (int)hi; }
    }

    const FT mu0 = LW ? FT(1) : __ldg(P.io.cos_zenith + col);
    const bool day = LW || mu0 > FT(0);
    const FT toa = LW ? FT(0) : __ldg(P.io.toa_flux + col);

    FT acc_up[kMaxLevPerLane], acc_dn[kMaxLevPerLane], acc_dir[kMaxLevPerLane];
#pragma unroll
    for (int i = 0; i < kMaxLevPerLane; ++i) acc_up[i] = acc_dn[i] = acc_dir[i] = FT(0);
    int n_cloudy = 0;

    for (int g0 = 0; g0 < n_gpt; g0 += 32) {
        W.set_block(g0);
        __syncwarp();
        FT aod_e, aod_s;
        W.phase1(aod_e, aod_s);
        if (!LW && P.use_aero != 0 && P.io.aod_ext != nullptr && P.aero.iband_550nm >= W.b_first + 1 &&
            P.aero.iband_550nm <= W.b_first + W.nb) {
            aod_e = warp_sum(aod_e); aod_s = warp_sum(aod_s);
            if (lane == 0) { P.io.aod_ext[col] = aod_e; P.io.aod_sca[col] = aod_s; }
        }
        n_cloudy += W.mcica(col_key, cld_start, cld_finish);
        if (!day) continue;   // night: masks/AOD only (shortwave_2stream.jl:66-102)

        const int gpt = W.gpt, ibnd = W.ibnd, bl = W.bl;
        const bool lane_on = W.lane_on;
        auto S = [&](int lev, int v) -> FT& { return store[(lev * NV + v) * 32 + lane]; };

        if (MODE == MODE_LW_2STREAM) {
            // compute_optical_props.jl:157-195 sources + longwave_2stream.jl:243-334 adding
            const FT* pb = W.plk + bl * 2 * nlev;
            const FT emis = __ldg(P.io.sfc_emis + (size_t)col * L.n_bnd + ibnd);
            const FT inc = P.io.inc_flux_lw ? __ldg(P.io.inc_flux_lw + (size_t)gpt * P.ncol_total + col) : FT(0);
            FT tau, ssa, g, pf;
            W.optics(0, tau, ssa, g, pf);
            FT lev_bot = pb[0] * pf;
            FT albedo = FT(1) - emis;
            FT src = Num<FT>::pi() * emis * (pb[nlev + nlay] * pf);
            for (int k = 0; k < nlay; ++k) {
                FT tau_n = FT(0), ssa_n = FT(0), g_n = FT(0), pf_n = FT(0), lev_top;
                FT inc_k = pb[k + 1] * pf;                       // lev_src_inc of layer k
                if (k + 1 < nlay) {
                    W.optics(k + 1, tau_n, ssa_n, g_n, pf_n);
                    lev_top = hsqrt(inc_k * (pb[k + 1] * pf_n));   // sqrt(inc_prev * dec)
                } else {
                    lev_top = inc_k;
                }
                FT Rdif, Tdif, su, sd;
                lw_2stream_coeffs(tau, ssa, g, lev_bot, lev_top, Rdif, Tdif, su, sd);
                FT denom = hdiv(FT(1), FT(1) - Rdif * albedo);
                // level k (bottom of layer k): what the downward sweep needs
                S(k, 0) = Tdif * denom;                           // A_k
                S(k, 1) = (Rdif * src + sd) * denom;              // B_k
                S(k, 2) = albedo;
                S(k, 3) = src;
                FT albedo_n = Rdif + Tdif * Tdif * albedo * denom;
                src = su + Tdif * denom * (src + albedo * sd);
                albedo = albedo_n;
                lev_bot = lev_top; tau = tau_n; ssa = ssa_n; g = g_n; pf = pf_n;
            }
            // top of domain, then downward sweep: F_dn(k) = A_k F_dn(k+1) + B_k ; F_up(k) = alb_k F_dn(k) + src_k
            FT dn = inc;
            FT up = dn * albedo + src;
            S(nlay, 0) = lane_on ? up : FT(0);
            S(nlay, 1) = lane_on ? dn : FT(0);
            for (int k = nlay - 1; k >= 0; --k) {
                dn = S(k, 0) * dn + S(k, 1);
                up = dn * S(k, 2) + S(k, 3);
                S(k, 0) = lane_on ? up : FT(0);
                S(k, 1) = lane_on ? dn : FT(0);
            }
        } else if (MODE == MODE_SW_2STREAM) {
            // shortwave_2stream.jl:300-392
            const FT alb_dir = __ldg(P.io.sfc_alb_direct + (size_t)col * L.n_bnd + ibnd);
            const FT alb_dif = __ldg(P.io.sfc_alb_diffuse + (size_t)col * L.n_bnd + ibnd);
            const FT dir_top = toa * __ldg(L.solar_src_scaled + gpt) * mu0;
            const FT inv_mu0 = hdiv(FT(1), rmax(mu0, Num<FT>::eps()));
            FT tau_cum = FT(0), dir_above = dir_top;
            S(nlay, 4) = dir_top;
            for (int k = nlay - 1; k >= 0; --k) {   // direct beam + layer coefficients, top down
                FT tau, ssa, g, pf;
                W.optics(k, tau, ssa, g, pf);
                FT Rdir, Tdir, Rdif, Tdif;
                sw_2stream_coeffs(tau, ssa, g, mu0, inv_mu0, Rdir, Tdir, Rdif, Tdif);
                tau_cum += tau;
                S(k, 0) = Rdif; S(k, 1) = Tdif;
                S(k, 2) = Rdir * dir_above;           // src_up of layer k
                S(k, 3) = Tdir * dir_above;           // src_dn of layer k
                dir_above = dir_top * hexp(-tau_cum * inv_mu0);
                S(k, 4) = dir_above;                  // direct flux at level k
            }
            FT albedo = alb_dif, src = dir_above * alb_dir;
            for (int k = 0; k < nlay; ++k) {        // bottom up: albedo / source of everything below
                FT Rdif = S(k, 0), Tdif = S(k, 1), su = S(k, 2), sd = S(k, 3);
                FT denom = hdiv(FT(1), FT(1) - Rdif * albedo);
                S(k, 0) = Tdif * denom;
                S(k, 1) = (Rdif * src + sd) * denom;
                S(k, 2) = albedo;
                S(k, 3) = src;
                FT albedo_n = Rdif + Tdif * Tdif * albedo * denom;
                src = su + Tdif * denom * (src + albedo * sd);
                albedo = albedo_n;
            }
            FT dn = FT(0);                           // diffuse incident flux (shortwave_2stream.jl:331)
            FT up = dn * albedo + src;
            S(nlay, 0) = lane_on ? up : FT(0);
            S(nlay, 1) = lane_on ? dn + S(nlay, 4) : FT(0);
            if (!lane_on) S(nlay, 4) = FT(0);
            for (int k = nlay - 1; k >= 0; --k) {
                dn = S(k, 0) * dn + S(k, 1);
                up = dn * S(k, 2) + S(k, 3);
                S(k, 0) = lane_on ? up : FT(0);
                S(k, 1) = lane_on ? dn + S(k, 4) : FT(0);
                if (!lane_on) S(k, 4) = FT(0);
            }
        } else {
            // compute_optical_props.jl:43-82 sources + longwave_noscat.jl:224-301 per angle
            const FT* pb = W.plk + bl * 2 * nlev;
            const FT emis = __ldg(P.io.sfc_emis + (size_t)col * L.n_bnd + ibnd);
            const bool has_inc = P.io.inc_flux_lw != nullptr;
            const FT inc = has_inc ? __ldg(P.io.inc_flux_lw + (size_t)gpt * P.ncol_total + col) : FT(0);
            FT sfc_source = FT(0), inc_prev = FT(0);
            for (int k = 0; k < nlay; ++k) {
                FT tau, ssa, g, pf;
                W.optics(k, tau, ssa, g, pf);
                S(k, 0) = tau;
                S(k, 1) = pb[nlev + k] * pf;              // lay_source
                FT src_inc = pb[k + 1] * pf, src_dec = pb[k] * pf;
                if (k == 0) { sfc_source = pb[nlev + nlay] * pf; S(0, 2) = src_dec; }
                else S(k, 2) = hsqrt(inc_prev * src_dec);
                inc_prev = src_inc;
                S(k, 3) = FT(0); S(k, 4) = FT(0);
            }
            S(nlay, 2) = inc_prev; S(nlay, 3) = FT(0); S(nlay, 4) = FT(0);
            for (int imu = 0; imu < P.n_mu; ++imu) {
                const FT Ds = P.Ds[imu], i2f = Num<FT>::pi() * P.wts[imu];
                FT I = has_inc ? inc / Num<FT>::pi() : FT(0);
                S(nlay, 4) += I * i2f;
                for (int k = nlay - 1; k >= 0; --k) {
                    FT tau_loc = S(k, 0) * Ds;
                    FT trans = hexp(-tau_loc);
                    I = trans * I + lw_noscat_source(S(k, 2), S(k, 1), tau_loc, trans);
                    S(k, 4) += I * i2f;
                }
                I = I * (FT(1) - emis) + emis * sfc_source;
                S(0, 3) += I * i2f;
                for (int k = 1; k <= nlay; ++k) {
                    FT tau_loc = S(k - 1, 0) * Ds;
                    FT trans = hexp(-tau_loc);
                    I = trans * I + lw_noscat_source(S(k, 2), S(k - 1, 1), tau_loc, trans);
                    S(k, 3) += I * i2f;
                }
            }
            if (!lane_on)
                for (int k = 0; k <= nlay; ++k) { S(k, 3) = FT(0); S(k, 4) = FT(0); }
        }
        __syncwarp();

        // ---------------- g-point reduction: lane <-> level (driver_utils.jl:37-84) ----------------
        {
            constexpr int VU = NOSCAT ? 3 : 0, VD = NOSCAT ? 4 : 1;
#pragma unroll
            for (int i = 0; i < kMaxLevPerLane; ++i) {
                const int lev = lane + 32 * i;
                if (lev < nlev) {
                    const FT* su = store + (lev * NV + VU) * 32;
                    const FT* sd = store + (lev * NV + VD) * 32;
                    FT u = FT(0), d = FT(0), dr = FT(0);
#pragma unroll 8
                    for (int j = 0; j < 32; ++j) {
                        const int l2 = (lane + j) & 31;
                        u += su[l2]; d += sd[l2];
                        if (MODE == MODE_SW_2STREAM) dr += store[(lev * NV + 4) * 32 + l2];
                    }
                    acc_up[i] += u; acc_dn[i] += d; acc_dir[i] += dr;
                    if (P.io.band_up != nullptr) {   // Fluxes.jl:199-215, bands of this block
                        for (int b = 0; b < W.nb; ++b) {
                            FT bu = FT(0), bd = FT(0);
                            for (int j = 0; j < 32; ++j) {
                                const int l2 = (lane + j) & 31;
                                if (g0 + l2 < n_gpt && __ldg(L.gpt2bnd + g0 + l2) == W.b_first + b) { bu += su[l2]; bd += sd[l2]; }
                            }
                            size_t o = ((size_t)(W.b_first + b) * P.ncol_total + col) * nlev + lev;
                            P.io.band_up[o] += bu; P.io.band_dn[o] += bd;
                        }
                    }
                }
            }
        }
    }

    // ---------------- epilogue: (nlev, ncol) presentation, net, scaling, diagnostics ----------------
#pragma unroll
    for (int i = 0; i < kMaxLevPerLane; ++i) {
        const int lev = lane + 32 * i;
        if (lev < nlev) {
            const size_t o = (size_t)col * nlev + lev;
            FT up = day ? acc_up[i] : FT(0), dn = day ? acc_dn[i] : FT(0), dr = day ? acc_dir[i] : FT(0);
            FT net = up - dn;                                    // Fluxes.jl:225-233
            if (P.io.metric_scaling != nullptr) {                // Fluxes.jl:295-304, after net
                FT sc = __ldg(P.io.metric_scaling + o);
                up *= sc; dn *= sc; net *= sc; dr *= sc;
            }
            P.io.out_up[o] = up; P.io.out_dn[o] = dn; P.io.out_net[o] = net;
            if (!LW) P.io.out_dir[o] = dr;
            if (P.io.out_total_net != nullptr) P.io.out_total_net[o] = P.io.add_net[o] + net;   // Fluxes.jl:423-435
            if (P.io.band_up != nullptr) {
                for (int b = 0; b < L.n_bnd; ++b) {
                    size_t ob = ((size_t)b * P.ncol_total + col) * nlev + lev;
                    FT bu = P.io.band_up[ob], bd = P.io.band_dn[ob];
                    if (!day) { bu = FT(0); bd = FT(0); }
                    if (P.io.metric_scaling != nullptr) { FT sc = __ldg(P.io.metric_scaling + o); bu *= sc; bd *= sc; }
                    P.io.band_up[ob] = bu; P.io.band_dn[ob] = bd; P.io.band_net[ob] = bu - bd;
                }
            }
        }
    }
    if (lane == 0 && P.io.cld_cover != nullptr && use_cloud)
        P.io.cld_cover[col] = FT(n_cloudy) / FT(n_gpt);         // ext/cuda/rte_longwave_2stream.jl:131-138
}

// Launches solve_kernel<FT, MODE>; returns cudaError_t as int.  Defined in solver.cu.
template <typename FT> int launch_solve(int mode, SolveParams<FT>& P, int max_smem_optin, void* stream);
// Fills the shared-memory layout fields of P; returns bytes per warp.
template <typename FT> int plan_smem(int mode, SolveParams<FT>& P);

}  // namespace rb

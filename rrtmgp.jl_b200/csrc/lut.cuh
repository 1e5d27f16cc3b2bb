// Device-resident lookup tables, re-laid out for the "lane = g-point" kernels.
//
// The reference keeps every table eta-fastest (`kmajor (n_eta, n_p, n_t, n_gpt)`,
// src/optics/LookUpTables.jl:130-143) and has one thread per column, so each trilinear
// corner is a scattered 4-byte load.  Here the g-point axis is the
// fastest one: the 32 lanes of a warp are 32 consecutive g-points, the 16 lanes of a
// band share (j_eta, j_p, j_T), and one interpolation corner is one 64-byte segment per band.
#pragma once
#include <cstdint>

namespace rb {

template <typename FT>
struct GasLut {
    int n_gpt, n_bnd, n_eta, n_p, n_p_ref, n_t, ngas1, n_t_plnk;
    int idx_h2o;
    int nminor_max;   // max minor absorbers of any (band, lower/upper)
    int maxb;         // max number of bands that touch one 32-g-point block
    int is_sw;
    int bands_of_16;  // every band is 16 consecutive g-points (the real g256 / g224 tables)
    FT p_ref_tropo, p_ref_min, t_ref_min, t_ref_max, solar_src_tot;
    const FT* t_ref;        // [n_t]
    const FT* ln_p_ref;     // [n_p_ref]
    const FT* vmr_ref;      // reference order (2, ngas1, n_t) -> [n_t][ngas1][2]
    const int* key_species; // reference order (2, 2, n_bnd) -> [n_bnd][tropo][2], (0,0)->(2,2) applied
    const int* gpt2bnd;     // [n_gpt], 0-based band
    const FT* kmajor;       // [n_p][n_t][n_eta][n_gpt]
    const FT* pfrac;        // [n_p][n_t][n_eta][n_gpt]          (LW)
    const FT* kmaj_pf;      // [n_p][n_t][n_eta][n_gpt][2] {kmajor, pfrac} pairs (LW; fast kernels: one 64-bit load per corner)
    const FT* t_planck;     // [n_t_plnk]                         (LW)
    const FT* tot_planck;   // [n_bnd][n_t_plnk]                  (LW)
    const FT* rayl;         // [2 (lower, upper)][n_t][n_eta][n_gpt]  (SW)
    const FT* solar_src_scaled; // [n_gpt]                        (SW)
    // minor absorbers, index 0 = lower atmosphere, 1 = upper (LookUpTables.jl:36-53)
    const int* minor_bnd_st[2];  // [n_bnd+1], 0-based rows of gasdata
    const int* minor_gasdata[2]; // [n_abs][4] (idx_gas, idx_scaling_gas, scales_with_density, scale_by_complement)
    const FT* kminor[2];         // [slot < nminor_max][n_t][n_eta][n_gpt]; contributor (gpt, slot)
    const FT* kminor4[2];        // fast kernels: [group][n_t][n_eta][n_gpt][4], four slots per 128-bit load;
                                 // SW: slot 0 = Rayleigh of that atmosphere, minors follow; zero padded
    int n_minor_groups;          // groups of four slots in kminor4
    // Every SMALL table of this sweep (the gas tables above that are not g-point sized, the cloud tables and the
    // aerosol tables except sea salt) sits in one contiguous block of the arena, so the fast kernels stage the
    // block into shared memory with one TMA bulk copy and address a table at (its pointer - blob) in the copy.
    // When the whole block does not fit beside the per-warp state, a prefix ending at one of `blob_cut` (table
    // boundaries, ascending; tables are ordered by how often a column reads them) is staged and the rest is
    // read from global memory.
    const unsigned char* blob;
    int blob_bytes;              // multiple of 16
    int n_blob_cut;
    int blob_cut[32];
};

template <typename FT>
struct CldLut {   // LookUpTables.jl:239-284
    int nband, nrghice, nsize_liq, nsize_ice;
    FT radliq_lwr, radliq_upr, radice_lwr, radice_upr;
    const FT* liqdata;  // [nband][3*nsize_liq]  (ext | ssa | asy)
    const FT* icedata;  // [nrghice][nband][3*nsize_ice]
};

template <typename FT>
struct AeroLut {  // LookUpTables.jl:312-325, reference (first-index-fastest) layouts kept
    int nband, nbin, nrh, iband_550nm;   // iband_550nm 1-based, 0 = none
    const FT* size_bin_limits;  // [nbin][2]
    const FT* rh_levels;        // [nrh]
    const FT* dust;             // [nband][nbin][3]
    const FT* sea_salt;         // [nband][nbin][nrh][3]
    const FT* sulfate;          // [nband][nrh][3]
    const FT* black_carbon_rh;  // [nband][nrh][3]
    const FT* black_carbon;     // [nband][3]
    const FT* organic_carbon_rh;
    const FT* organic_carbon;
};

template <typename FT>
struct Luts {
    GasLut<FT> lw, sw;
    CldLut<FT> cld_lw, cld_sw;
    AeroLut<FT> aero_lw, aero_sw;
};

// Host-side owner of the device arena.
struct LutStore {
    void* arena = nullptr;
    size_t arena_bytes = 0;
    int n_gpt_lw = 0, n_bnd_lw = 0, n_gpt_sw = 0, n_bnd_sw = 0, ngas = 0, iband_550nm = 0;
    double p_ref_min = 0, t_ref_min = 0, t_ref_max = 0, solar_src_tot = 0;
    Luts<float> f32;
    Luts<double> f64;
    bool loaded = false;
};

// Parses a LUT pack (host memory), converts to float or double, re-lays out and uploads.
// Returns an rrtmgp_b200_status.  `err` receives a CUDA error string on failure.
int load_lut_pack(LutStore& store, const void* pack, size_t nbytes, bool f64, const char** err);
void free_lut_store(LutStore& store);

}  // namespace rb

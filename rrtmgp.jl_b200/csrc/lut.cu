// LUT pack -> device tables.  Replaces the reference's NetCDF constructors
// (ext/lookup_constructors.jl) downstream of the file read: the pack already holds the
// post-load arrays; this file converts precision, applies the `(0,0) -> (2,2)` key-species
// rewrite (lookup_constructors.jl:175-182) and re-lays the spectral tables out
// g-point-fastest (lut.cuh).
#include "lut.cuh"

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/rrtmgp_b200.h"

namespace rb {
namespace {

struct Entry { int dtype, ndim; uint32_t dims[6]; const unsigned char* data; size_t nbytes; };
typedef std::map<std::string, Entry> Pack;

bool parse(const unsigned char* buf, size_t n, Pack& out) {
    static const char magic[16] = {'R','R','T','M','G','P','B','2','0','0','L','U','T',0,0,0};
    if (!buf || n < 32 || std::memcmp(buf, magic, 16) != 0) return false;
    uint32_t version, nent; uint64_t total;
    std::memcpy(&version, buf + 16, 4); std::memcpy(&nent, buf + 20, 4); std::memcpy(&total, buf + 24, 8);
    if (version != 1 || total != n || 32 + (size_t)nent * 80 > n) return false;
    for (uint32_t i = 0; i < nent; ++i) {
        const unsigned char* e = buf + 32 + 80 * (size_t)i;
        char name[33]; std::memcpy(name, e, 32); name[32] = 0;
        Entry en; uint32_t dt, nd; uint64_t off, nb;
        std::memcpy(&dt, e + 32, 4); std::memcpy(&nd, e + 36, 4); std::memcpy(en.dims, e + 40, 24);
        std::memcpy(&off, e + 64, 8); std::memcpy(&nb, e + 72, 8);
        if (off > n || nb > n - off || nd > 6) return false;
        en.dtype = (int)dt; en.ndim = (int)nd; en.data = buf + off; en.nbytes = nb;
        out[name] = en;
    }
    return true;
}

struct Missing { std::string name; };

const Entry& get(const Pack& p, const std::string& k, int dtype) {
    auto it = p.find(k);
    if (it == p.end() || it->second.dtype != dtype) throw Missing{k};
    return it->second;
}
std::vector<double> getd(const Pack& p, const std::string& k) {
    const Entry& e = get(p, k, 0);
    std::vector<double> v(e.nbytes / 8);
    std::memcpy(v.data(), e.data, v.size() * 8);
    return v;
}
std::vector<int> geti(const Pack& p, const std::string& k) {
    const Entry& e = get(p, k, 1);
    std::vector<int> v(e.nbytes / 4);
    std::memcpy(v.data(), e.data, v.size() * 4);
    return v;
}

// host staging arena; pointers are patched after the single upload
struct Arena {
    std::vector<unsigned char> buf;
    struct Fix { void** where; size_t off; };
    std::vector<Fix> fixes;
    template <class T, class P> void add(const std::vector<T>& v, P*& field) {
        size_t off = (buf.size() + 255) / 256 * 256;
        buf.resize(off + std::max<size_t>(v.size() * sizeof(T), 16));
        if (!v.empty()) std::memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
        fixes.push_back({(void**)&field, off});
    }
    // Small tables of one sweep (group 0 = LW, 1 = SW) are collected and laid out contiguously by finalize().
    // `glued`: the table is addressed relative to the previous one (aerosol tables relative to `dust`), so the
    // staged prefix must not end between them.
    struct Small { std::vector<unsigned char> buf; std::vector<Fix> fixes; std::vector<int> cuts; };
    Small small[2];
    template <class T, class P> void add_small(int grp, const std::vector<T>& v, P*& field, bool glued = false) {
        Small& S = small[grp];
        size_t off = (S.buf.size() + 15) / 16 * 16;
        if (!glued && off > 0) S.cuts.push_back((int)off);
        S.buf.resize(off + std::max<size_t>(v.size() * sizeof(T), 16));
        if (!v.empty()) std::memcpy(S.buf.data() + off, v.data(), v.size() * sizeof(T));
        S.fixes.push_back({(void**)&field, off});
    }
    template <class P> void finalize_small(int grp, P*& blob, int& blob_bytes, int& n_cut, int* cut) {
        Small& S = small[grp];
        S.buf.resize((S.buf.size() + 15) / 16 * 16);
        n_cut = 0;
        for (size_t i = 0; i < S.cuts.size() && n_cut < 31; ++i) cut[n_cut++] = S.cuts[i];   // table starts
        cut[n_cut++] = (int)S.buf.size();
        size_t off = (buf.size() + 255) / 256 * 256;
        buf.resize(off + S.buf.size());
        std::memcpy(buf.data() + off, S.buf.data(), S.buf.size());
        for (auto& f : S.fixes) fixes.push_back({f.where, off + f.off});
        fixes.push_back({(void**)&blob, off});
        blob_bytes = (int)S.buf.size();
    }
    void patch(unsigned char* base) {
        for (auto& f : fixes) *f.where = base + f.off;
    }
};

template <class FT> std::vector<FT> cast(const std::vector<double>& v) {
    std::vector<FT> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = (FT)v[i];
    return o;
}

template <class FT>
void build_gas(const Pack& p, const std::string& pre, bool sw, GasLut<FT>& L, Arena& A) {
    const int grp = sw ? 1 : 0;
    const Entry& km = get(p, pre + "/kmajor", 0);
    if (km.ndim != 4) throw Missing{pre + "/kmajor (ndim)"};
    const int n_eta = km.dims[0], n_p = km.dims[1], n_t = km.dims[2], n_gpt = km.dims[3];
    L.n_eta = n_eta; L.n_p = n_p; L.n_t = n_t; L.n_gpt = n_gpt; L.is_sw = sw;
    L.n_bnd = get(p, pre + "/key_species", 1).dims[2];
    L.ngas1 = get(p, pre + "/vmr_ref", 0).dims[1];
    std::vector<double> prm = getd(p, pre + "/params");
    if (prm.size() < 5) throw Missing{pre + "/params (size)"};
    L.p_ref_tropo = (FT)prm[0]; L.p_ref_min = (FT)prm[1]; L.t_ref_min = (FT)prm[2]; L.t_ref_max = (FT)prm[3];
    L.solar_src_tot = (FT)prm[4];
    L.idx_h2o = geti(p, pre + "/idx_h2o")[0];

    std::vector<int> ks = geti(p, pre + "/key_species");
    for (size_t i = 0; i + 1 < ks.size(); i += 2)
        if (ks[i] == 0 && ks[i + 1] == 0) ks[i] = ks[i + 1] = 2;
    A.add_small(grp, ks, L.key_species);

    std::vector<int> g2b = geti(p, pre + "/major_gpt2bnd");
    for (auto& b : g2b) b -= 1;
    int maxb = 1;
    for (int g0 = 0; g0 < n_gpt; g0 += 32)
        maxb = std::max(maxb, g2b[std::min(g0 + 31, n_gpt - 1)] - g2b[g0] + 1);
    L.maxb = maxb;
    L.bands_of_16 = (n_gpt == 16 * L.n_bnd) ? 1 : 0;
    for (int g = 0; g < n_gpt && L.bands_of_16; ++g)
        if (g2b[g] != g / 16) L.bands_of_16 = 0;
    A.add_small(grp, g2b, L.gpt2bnd);

    // `p_ref` as the file has it, or -- from a host that dumps the loaded struct, which only keeps the logarithm
    // (ReferencePoints, LookUpTables.jl:70-74) -- `ln_p_ref` as is
    std::vector<FT> lnp;
    if (p.count(pre + "/ln_p_ref")) {
        lnp = cast<FT>(getd(p, pre + "/ln_p_ref"));
    } else {
        std::vector<FT> p_ref = cast<FT>(getd(p, pre + "/p_ref"));
        lnp.resize(p_ref.size());
        for (size_t i = 0; i < p_ref.size(); ++i) lnp[i] = std::log(p_ref[i]);  // lookup_constructors.jl:336
    }
    L.n_p_ref = (int)lnp.size();
    A.add_small(grp, lnp, L.ln_p_ref);
    A.add_small(grp, cast<FT>(getd(p, pre + "/t_ref")), L.t_ref);
    A.add_small(grp, cast<FT>(getd(p, pre + "/vmr_ref")), L.vmr_ref);

    auto relayout4 = [&](const std::string& name) {
        std::vector<double> src = getd(p, name);
        std::vector<FT> dst(src.size());
        for (int g = 0; g < n_gpt; ++g)
            for (int t = 0; t < n_t; ++t)
                for (int pp = 0; pp < n_p; ++pp)
                    for (int e = 0; e < n_eta; ++e)
                        dst[(((size_t)pp * n_t + t) * n_eta + e) * n_gpt + g] =
                            (FT)src[e + (size_t)n_eta * (pp + (size_t)n_p * (t + (size_t)n_t * g))];
        return dst;
    };
    A.add(relayout4(pre + "/kmajor"), L.kmajor);

    // minor absorbers: dense [slot][n_t][n_eta][n_gpt]
    int nmax = 0;
    std::vector<int> bst[2], gst[2];
    const char* tags[2] = {"/minor_lower", "/minor_upper"};
    for (int tr = 0; tr < 2; ++tr) {
        bst[tr] = geti(p, pre + tags[tr] + "/bnd_st");
        gst[tr] = geti(p, pre + tags[tr] + "/gpt_st");
        for (int b = 0; b < L.n_bnd; ++b) nmax = std::max(nmax, bst[tr][b + 1] - bst[tr][b]);
    }
    L.nminor_max = nmax;
    for (int tr = 0; tr < 2; ++tr) {
        std::vector<double> kmin = getd(p, pre + tags[tr] + "/kminor");
        std::vector<FT> dst((size_t)std::max(nmax, 1) * n_t * n_eta * n_gpt, FT(0));
        for (int g = 0; g < n_gpt; ++g) {
            int n = gst[tr][g + 1] - gst[tr][g];
            int b = g2b[g];
            if (n != bst[tr][b + 1] - bst[tr][b]) throw Missing{pre + tags[tr] + " (gpt_st/bnd_st mismatch)"};
            for (int i = 0; i < n; ++i) {
                size_t c = (size_t)(gst[tr][g] - 1 + i);
                for (int t = 0; t < n_t; ++t)
                    for (int e = 0; e < n_eta; ++e)
                        dst[(((size_t)i * n_t + t) * n_eta + e) * n_gpt + g] =
                            (FT)kmin[e + (size_t)n_eta * (t + (size_t)n_t * c)];
            }
        }
        A.add(dst, L.kminor[tr]);
        {   // 128-bit packed copy for the fast kernels (SW: Rayleigh in slot 0)
            const int nslots = nmax + (sw ? 1 : 0), ngrp = std::max((nslots + 3) / 4, 1);
            L.n_minor_groups = ngrp;
            std::vector<FT> d4((size_t)ngrp * n_t * n_eta * n_gpt * 4, FT(0));
            std::vector<double> ray;
            if (sw) ray = getd(p, pre + (tr == 0 ? "/rayl_lower" : "/rayl_upper"));
            for (int sl = 0; sl < nslots; ++sl)
                for (int t = 0; t < n_t; ++t)
                    for (int e = 0; e < n_eta; ++e)
                        for (int g = 0; g < n_gpt; ++g) {
                            FT v;
                            if (sw && sl == 0) v = (FT)ray[e + (size_t)n_eta * (t + (size_t)n_t * g)];
                            else v = dst[(((size_t)(sl - (sw ? 1 : 0)) * n_t + t) * n_eta + e) * n_gpt + g];
                            d4[(((((size_t)(sl / 4) * n_t + t) * n_eta + e) * n_gpt + g) * 4) + (sl % 4)] = v;
                        }
            A.add(d4, L.kminor4[tr]);
        }
        std::vector<int> b0 = bst[tr];
        for (auto& v : b0) v -= 1;
        A.add_small(grp, b0, L.minor_bnd_st[tr]);
        A.add_small(grp, geti(p, pre + tags[tr] + "/gasdata"), L.minor_gasdata[tr]);
    }

    L.pfrac = nullptr; L.t_planck = nullptr; L.tot_planck = nullptr; L.rayl = nullptr; L.solar_src_scaled = nullptr;
    L.n_t_plnk = 0;
    L.kmaj_pf = nullptr;
    if (!sw) {
        {
            std::vector<double> km = getd(p, pre + "/kmajor"), pfr = getd(p, pre + "/planck_fraction");
            std::vector<FT> dst(km.size() * 2);
            for (int g = 0; g < n_gpt; ++g)
                for (int t = 0; t < n_t; ++t)
                    for (int pp = 0; pp < n_p; ++pp)
                        for (int e = 0; e < n_eta; ++e) {
                            size_t src = e + (size_t)n_eta * (pp + (size_t)n_p * (t + (size_t)n_t * g));
                            size_t d0 = ((((size_t)pp * n_t + t) * n_eta + e) * n_gpt + g) * 2;
                            dst[d0] = (FT)km[src]; dst[d0 + 1] = (FT)pfr[src];
                        }
            A.add(dst, L.kmaj_pf);
        }
        A.add(relayout4(pre + "/planck_fraction"), L.pfrac);
        std::vector<FT> tp = cast<FT>(getd(p, pre + "/t_planck"));
        L.n_t_plnk = (int)tp.size();
        A.add_small(grp, tp, L.t_planck);
        A.add_small(grp, cast<FT>(getd(p, pre + "/tot_planck")), L.tot_planck);
    } else {
        std::vector<double> lo = getd(p, pre + "/rayl_lower"), up = getd(p, pre + "/rayl_upper");
        std::vector<FT> dst((size_t)2 * n_t * n_eta * n_gpt);
        for (int tr = 0; tr < 2; ++tr) {
            const std::vector<double>& src = tr == 0 ? lo : up;
            for (int g = 0; g < n_gpt; ++g)
                for (int t = 0; t < n_t; ++t)
                    for (int e = 0; e < n_eta; ++e)
                        dst[(((size_t)tr * n_t + t) * n_eta + e) * n_gpt + g] =
                            (FT)src[e + (size_t)n_eta * (t + (size_t)n_t * g)];
        }
        A.add(dst, L.rayl);
        A.add(cast<FT>(getd(p, pre + "/solar_src_scaled")), L.solar_src_scaled);
    }
}

template <class FT> void build_cld(const Pack& p, const std::string& pre, CldLut<FT>& C, Arena& A, int grp) {
    std::vector<int> d = geti(p, pre + "/dims");
    C.nband = d[0]; C.nrghice = d[1]; C.nsize_liq = d[2]; C.nsize_ice = d[3];
    std::vector<double> b = getd(p, pre + "/bounds");
    C.radliq_lwr = (FT)b[0]; C.radliq_upr = (FT)b[1]; C.radice_lwr = (FT)b[2]; C.radice_upr = (FT)b[3];
    A.add_small(grp, cast<FT>(getd(p, pre + "/liqdata")), C.liqdata);
    A.add_small(grp, cast<FT>(getd(p, pre + "/icedata")), C.icedata);
}

template <class FT> void build_aero(const Pack& p, const std::string& pre, AeroLut<FT>& L, Arena& A, int grp) {
    std::vector<int> d = geti(p, pre + "/dims");
    L.nband = d[0]; L.nbin = d[2]; L.nrh = d[3];
    L.iband_550nm = geti(p, pre + "/iband_550nm")[0];
    A.add_small(grp, cast<FT>(getd(p, pre + "/size_bin_limits")), L.size_bin_limits);
    A.add_small(grp, cast<FT>(getd(p, pre + "/rh_levels")), L.rh_levels);
    A.add_small(grp, cast<FT>(getd(p, pre + "/dust")), L.dust);
    A.add(cast<FT>(getd(p, pre + "/sea_salt")), L.sea_salt);
    A.add_small(grp, cast<FT>(getd(p, pre + "/sulfate")), L.sulfate, true);
    A.add_small(grp, cast<FT>(getd(p, pre + "/black_carbon_rh")), L.black_carbon_rh, true);
    A.add_small(grp, cast<FT>(getd(p, pre + "/black_carbon")), L.black_carbon, true);
    A.add_small(grp, cast<FT>(getd(p, pre + "/organic_carbon_rh")), L.organic_carbon_rh, true);
    A.add_small(grp, cast<FT>(getd(p, pre + "/organic_carbon")), L.organic_carbon, true);
}

template <class FT> void build_all(const Pack& p, Luts<FT>& L, Arena& A) {
    build_gas(p, "lw", false, L.lw, A);
    build_gas(p, "sw", true, L.sw, A);
    // block order = staging priority: gas tables, aerosol tables (sea salt, the one big one, stays outside), cloud
    // cloud / aerosol sections are optional (a clear-sky host loads the two gas files only,
    // ext/RRTMGPNCDatasetsExt.jl:26-92); rrtmgp_b200_load_luts rejects a pack that lacks what the method needs
    L.aero_lw = AeroLut<FT>{}; L.aero_sw = AeroLut<FT>{}; L.cld_lw = CldLut<FT>{}; L.cld_sw = CldLut<FT>{};
    if (p.count("aero_lw/dims") && p.count("aero_sw/dims")) {
        build_aero(p, "aero_lw", L.aero_lw, A, 0);
        build_aero(p, "aero_sw", L.aero_sw, A, 1);
    }
    if (p.count("cld_lw/dims") && p.count("cld_sw/dims")) {
        build_cld(p, "cld_lw", L.cld_lw, A, 0);
        build_cld(p, "cld_sw", L.cld_sw, A, 1);
    }
    A.finalize_small(0, L.lw.blob, L.lw.blob_bytes, L.lw.n_blob_cut, L.lw.blob_cut);
    A.finalize_small(1, L.sw.blob, L.sw.blob_bytes, L.sw.n_blob_cut, L.sw.blob_cut);
}

}  // namespace

void free_lut_store(LutStore& s) {
    if (s.arena) cudaFree(s.arena);
    s.arena = nullptr; s.arena_bytes = 0; s.loaded = false;
}

int load_lut_pack(LutStore& s, const void* pack, size_t nbytes, bool f64, const char** err) {
    Pack p;
    if (!parse((const unsigned char*)pack, nbytes, p)) return RRTMGP_B200_ERR_BAD_LUT_PACK;
    Arena A;
    try {
        if (f64) build_all(p, s.f64, A); else build_all(p, s.f32, A);
    } catch (const Missing&) {
        return RRTMGP_B200_ERR_BAD_LUT_PACK;
    } catch (...) {
        return RRTMGP_B200_ERR_BAD_LUT_PACK;
    }
    free_lut_store(s);
    cudaError_t e = cudaMalloc(&s.arena, A.buf.size());
    if (e == cudaSuccess) e = cudaMemcpy(s.arena, A.buf.data(), A.buf.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (err) *err = cudaGetErrorString(e);
        free_lut_store(s);
        return RRTMGP_B200_ERR_CUDA;
    }
    s.arena_bytes = A.buf.size();
    A.patch((unsigned char*)s.arena);
    if (f64) {
        s.n_gpt_lw = s.f64.lw.n_gpt; s.n_bnd_lw = s.f64.lw.n_bnd; s.n_gpt_sw = s.f64.sw.n_gpt; s.n_bnd_sw = s.f64.sw.n_bnd;
        s.ngas = s.f64.lw.ngas1 - 1; s.iband_550nm = s.f64.aero_sw.iband_550nm;
        s.p_ref_min = s.f64.lw.p_ref_min; s.t_ref_min = s.f64.lw.t_ref_min; s.t_ref_max = s.f64.lw.t_ref_max;
        s.solar_src_tot = s.f64.sw.solar_src_tot;
    } else {
        s.n_gpt_lw = s.f32.lw.n_gpt; s.n_bnd_lw = s.f32.lw.n_bnd; s.n_gpt_sw = s.f32.sw.n_gpt; s.n_bnd_sw = s.f32.sw.n_bnd;
        s.ngas = s.f32.lw.ngas1 - 1; s.iband_550nm = s.f32.aero_sw.iband_550nm;
        s.p_ref_min = s.f32.lw.p_ref_min; s.t_ref_min = s.f32.lw.t_ref_min; s.t_ref_max = s.f32.lw.t_ref_max;
        s.solar_src_tot = s.f32.sw.solar_src_tot;
    }
    s.loaded = true;
    return RRTMGP_B200_OK;
}

}  // namespace rb

// Scalar building blocks of the hot path, shared by every kernel.  Each function names the
// reference routine it computes (file:line in CliMA/RRTMGP.jl v1.0.0).
//
// Precision policy.  Float64 kernels use IEEE division / sqrt and the full-precision libdevice
// exp / expm1 / log throughout.  Float32 kernels use IEEE division (`pdiv`) for everything in
// phase 0 -- in particular every quotient that is truncated to a table index (jt, jpress, cloud
// size interval, Planck interval), so the indices equal the reference's Float32 path -- and
// full-precision exp / log there; the SFU approximations (`hdiv` = div.approx.ftz, ex2.approx,
// sqrt.approx: <= 2 ulp) serve the per-(layer, band) fractions of phase 1, which only enter
// continuous piecewise-linear interpolations, and the per-(layer, g-point) loop, where a
// full-precision IEEE divide costs ~15 instructions with a divergent slow path.  The guard
// constants of src/Numerics.jl are kept exactly; 1 - exp(-x) keeps its small-x accuracy (the
// reason the reference uses expm1, longwave_2stream.jl:168-170) through a series below x = 0.25.
// Parity against the Float64 oracle is asserted in tests/test_gpu_parity.py with the
// reference's own Float32 thresholds.  Compile with -DRB_EXACT_MATH=1 to use IEEE everywhere.
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#ifndef RB_EXACT_MATH
#define RB_EXACT_MATH 0
#endif

namespace rb {

// ---- precision traits: src/Numerics.jl:24,37,49,63 ----
template <typename FT> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float eps() { return FLT_EPSILON; }
    static __device__ __forceinline__ float k_min() { return 3.4526698300e-04f; }       // sqrt(eps)
    static __device__ __forceinline__ float tau_thresh() { return 1.8581361171e-02f; }  // eps^(1/4)
    static __device__ __forceinline__ float pi() { return 3.14159274101257324f; }
};
template <> struct Num<double> {
    static __device__ __forceinline__ double eps() { return DBL_EPSILON; }
    static __device__ __forceinline__ double k_min() { return 1.4901161193847656e-08; }
    static __device__ __forceinline__ double tau_thresh() { return 1.220703125e-04; }
    static __device__ __forceinline__ double pi() { return 3.141592653589793; }
};

// full-precision forms (phase 0 / phase 1, and everything in Float64)
__device__ __forceinline__ float rexp(float x) { return expf(x); }
__device__ __forceinline__ double rexp(double x) { return exp(x); }
__device__ __forceinline__ float rexpm1(float x) { return expm1f(x); }
__device__ __forceinline__ double rexpm1(double x) { return expm1(x); }
__device__ __forceinline__ float rlog(float x) { return logf(x); }
__device__ __forceinline__ double rlog(double x) { return log(x); }
__device__ __forceinline__ float rsqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ double rsqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float rcos(float x) { return cosf(x); }
__device__ __forceinline__ double rcos(double x) { return cos(x); }
template <typename FT> __device__ __forceinline__ FT rmax(FT a, FT b) { return a > b ? a : b; }
template <typename FT> __device__ __forceinline__ FT rmin(FT a, FT b) { return a < b ? a : b; }
__device__ __forceinline__ float rabs(float a) { return fabsf(a); }
__device__ __forceinline__ double rabs(double a) { return fabs(a); }
// max(x, 0) as ONE instruction (FMNMX / DMNMX).  `rmax(0, x)` keeps the sign of a zero result and compiles to a compare
// and a select; where only the value matters (clamps of the two-stream coefficients) this form is used
__device__ __forceinline__ float rmax0(float x) { return fmaxf(x, 0.f); }
__device__ __forceinline__ double rmax0(double x) { return fmax(x, 0.0); }

// IEEE division (phase 0: table indices and everything else computed once per (column, layer))
__device__ __forceinline__ float pdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double pdiv(double a, double b) { return a / b; }
// hot-loop forms
__device__ __forceinline__ double hdiv(double a, double b) { return a / b; }
__device__ __forceinline__ double hrcp(double x) { return 1.0 / x; }
__device__ __forceinline__ double hsqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double hexp(double x) { return exp(x); }
__device__ __forceinline__ double h_one_minus_exp_neg(double x, double) { return -expm1(-x); }
#if RB_EXACT_MATH
__device__ __forceinline__ float hrcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float hdiv(float a, float b) { return a / b; }
__device__ __forceinline__ float hsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ float hexp(float x) { return expf(x); }
__device__ __forceinline__ float h_one_minus_exp_neg(float x, float) { return -expm1f(-x); }
#else
// reciprocal as ONE MUFU.RCP (div.approx of 1 by x costs an extra canonicalising FADD.FTZ)
__device__ __forceinline__ float hrcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float hdiv(float a, float b) {
    float r;
    asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float hsqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float hexp(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
// 1 - exp(-x), x >= 0, given e = exp(-x): Taylor series below 0.25 (truncation < 2e-9 relative)
__device__ __forceinline__ float h_one_minus_exp_neg(float x, float e) {
    float p = fmaf(x, -1.0f / 5040.0f, 1.0f / 720.0f);
    p = fmaf(x, -p, 1.0f / 120.0f);
    p = fmaf(x, -p, 1.0f / 24.0f);
    p = fmaf(x, -p, 1.0f / 6.0f);
    p = fmaf(x, -p, 0.5f);
    p = fmaf(x, -p, 1.0f);
    return x < 0.25f ? x * p : 1.0f - e;
}
#endif

// ---- optics_utils.jl:189-202 ----
template <typename FT>
__device__ __forceinline__ void increment_2stream(FT& t1, FT& s1, FT& g1, FT t2, FT s2, FT g2) {
    FT tau = t1 + t2;
    FT ssa = t1 * s1 + t2 * s2;
    FT ssag = hdiv(t1 * s1 * g1 + t2 * s2 * g2, rmax(Num<FT>::eps(), ssa));
    ssa = hdiv(ssa, rmax(Num<FT>::eps(), tau));
    t1 = tau; s1 = ssa; g1 = ssag;
}
// ---- optics_utils.jl:208-223 (phase 1: IEEE) ----
template <typename FT> __device__ __forceinline__ void delta_scale(FT& tau, FT& ssa, FT& g) {
    FT ssa_one_minus_g2 = ssa * (FT(1) - g) * (FT(1) + g);
    FT one_minus_wf = (FT(1) - ssa) + ssa_one_minus_g2;
    FT tau_s = one_minus_wf * tau;
    FT ssa_s = hdiv(ssa_one_minus_g2, rmax(Num<FT>::eps(), one_minus_wf));
    FT g_s = hdiv(g, rmax(Num<FT>::eps(), FT(1) + g));
    tau = tau_s; ssa = ssa_s; g = g_s;
}
// Small-table load: S = false reads global memory through the read-only path; S = true dereferences the pointer as
// it is (the fast kernels pass pointers into their shared-memory copy of the small tables).
template <bool S, typename T> __device__ __forceinline__ T ldt(const T* p) {
    if (S) return *p;
    return __ldg(p);
}
// ---- optics_utils.jl:7-14 (equispaced) ----
template <bool S = false, typename FT> __device__ __forceinline__ int loc_lower_eq(FT xi, FT dx, int n, const FT* __restrict__ x) {
    if (xi <= ldt<S>(x)) return 1;
    if (xi >= ldt<S>(x + n - 1)) return n - 1;
    int j = (int)pdiv(xi - ldt<S>(x), dx) + 1;
    return j < n - 1 ? j : n - 1;
}
// ---- optics_utils.jl:34-44 split into "locate" (per level) and "evaluate" (per band) ----
// loc = 0 encodes "below range -> y[0]", loc = n encodes "above range -> y[n-1]"
template <bool S = false, typename FT>
__device__ __forceinline__ void interp1d_eq_locate(FT xi, const FT* __restrict__ x, int n, int& loc, FT& factor) {
    if (xi < ldt<S>(x)) { loc = 0; factor = FT(0); return; }
    if (xi > ldt<S>(x + n - 1)) { loc = n; factor = FT(0); return; }
    FT dx = ldt<S>(x + 1) - ldt<S>(x);
    loc = loc_lower_eq<S>(xi, dx, n, x);
    factor = pdiv(xi - ldt<S>(x + loc - 1), dx);
}
template <bool S = false, typename FT>
__device__ __forceinline__ FT interp1d_eq_eval(int loc, FT factor, const FT* __restrict__ y, int n) {
    // branch-free: the two clamped cases read the end value twice with factor 0 (exactly y[0] / y[n-1])
    const int i0 = loc > 0 ? loc - 1 : 0, i1 = loc < n ? loc : n - 1;
    return ldt<S>(y + i0) * (FT(1) - factor) + ldt<S>(y + i1) * factor;
}
// ---- optics_utils.jl:51-62 + :21-27 (non-uniform x) ----
template <bool S = false, typename FT>
__device__ __forceinline__ void interp1d_loc_factor(FT xi, const FT* __restrict__ x, int n, int& loc, FT& factor) {
    if (xi < ldt<S>(x)) { loc = 1; factor = FT(0); return; }
    if (xi > ldt<S>(x + n - 1)) { loc = n - 1; factor = FT(1); return; }
    loc = n - 1;
    if (xi <= ldt<S>(x)) loc = 1;
    else
        for (int i = 1; i <= n; ++i)
            if (xi < ldt<S>(x + i - 1)) { loc = i - 1; break; }
    factor = pdiv(xi - ldt<S>(x + loc - 1), ldt<S>(x + loc) - ldt<S>(x + loc - 1));
}

// ---- longwave_2stream.jl:149-222 ----
template <typename FT>
__device__ __forceinline__ void lw_2stream_coeffs(FT tau, FT ssa, FT g, FT lev_src_bot, FT lev_src_top, FT& Rdif,
                                                  FT& Tdif, FT& src_up, FT& src_dn) {
    const FT lw_diff_sec = FT(1.66);
    FT g1 = lw_diff_sec * (FT(1) - FT(0.5) * ssa * (FT(1) + g));
    FT g2 = lw_diff_sec * FT(0.5) * ssa * (FT(1) - g);
    FT k = hsqrt(rmax(lw_diff_sec * (FT(1) - ssa) * (g1 + g2), Num<FT>::k_min()));
    FT tk = tau * k;
    FT e1 = hexp(-tk);
    FT om1 = h_one_minus_exp_neg(tk, e1);
    FT coeff = e1 * e1;
    FT one_minus_e2kt = om1 * (FT(1) + e1);
    FT RT_term = hdiv(FT(1), k * (FT(1) + coeff) + g1 * one_minus_e2kt);
    Rdif = RT_term * g2 * one_minus_e2kt;
    Tdif = RT_term * FT(2) * k * e1;
    // branch-free: a genuinely empty layer (tau = 0) emits nothing (longwave_2stream.jl:202-220)
    FT dB = lev_src_bot - lev_src_top;
    FT g_sum = g1 + g2;
    FT one_p_e1 = FT(1) + e1;
    FT emis_fac = om1 * (k * om1 + lw_diff_sec * (FT(1) - ssa) * one_p_e1) * RT_term;
    FT dBz = hdiv(dB * hdiv(om1, tau) * (k * om1 + g_sum * one_p_e1) * RT_term, rmax(g_sum, Num<FT>::eps()));
    FT su = Num<FT>::pi() * (lev_src_top * emis_fac - Tdif * dB + dBz);
    FT sd = Num<FT>::pi() * (lev_src_bot * emis_fac + Tdif * dB - dBz);
    src_up = tau > FT(0) ? su : FT(0);
    src_dn = tau > FT(0) ? sd : FT(0);
}

// longwave_2stream.jl:149-222 split for the fast kernels: everything that does not depend on the level sources
// (so it overlaps the table gathers of the next layer), then the two sources from
//   src_up = pi (B_top emis_fac - q dB),  src_dn = pi (B_bot emis_fac + q dB),  dB = B_bot - B_top,
// with q = Tdif - [(1 - e1)/tau (k (1 - e1) + (g1 + g2)(1 + e1)) RT / max(g1 + g2, eps)]; both vanish for tau = 0.
struct LwCoef { float Rdif, Tdif, emis_fac, q; };
__device__ __forceinline__ LwCoef lw_2stream_coeffs_nosrc(float tau, float ssa, float g) {
    const float lw_diff_sec = 1.66f;
    const float g1 = lw_diff_sec * (1.f - 0.5f * ssa * (1.f + g));
    const float g2 = lw_diff_sec * 0.5f * ssa * (1.f - g);
    const float g_sum = g1 + g2;
    const float k = hsqrt(rmax(lw_diff_sec * (1.f - ssa) * g_sum, Num<float>::k_min()));
    const float tk = tau * k;
    const float e1 = hexp(-tk);
    const float om1 = h_one_minus_exp_neg(tk, e1);
    const float one_p_e1 = 1.f + e1;
    const float one_minus_e2kt = om1 * one_p_e1;
    const float RT_term = hrcp(k * fmaf(e1, e1, 1.f) + g1 * one_minus_e2kt);
    LwCoef c;
    c.Rdif = RT_term * g2 * one_minus_e2kt;
    c.Tdif = RT_term * 2.f * k * e1;
    const float kom1 = k * om1;
    const float ef = om1 * (kom1 + lw_diff_sec * (1.f - ssa) * one_p_e1) * RT_term;
    const float dBfac = hdiv(hdiv(om1, tau) * (kom1 + g_sum * one_p_e1) * RT_term, rmax(g_sum, Num<float>::eps()));
    const bool has = tau > 0.f;
    c.emis_fac = has ? ef : 0.f;
    c.q = has ? c.Tdif - dBfac : 0.f;
    return c;
}

// ---- shortwave_2stream.jl:189-279 ----
template <typename FT>
__device__ __forceinline__ void sw_2stream_coeffs(FT tau, FT ssa, FT g, FT mu0, FT inv_mu0, FT& Rdir, FT& Tdir,
                                                  FT& Rdif, FT& Tdif) {
    // PIFM coefficients (shortwave_2stream.jl:193-196) with the constant factors folded in: (8 - ssa (5 + 3 g)) / 4 =
    // 2 - ssa (1.25 + 0.75 g), 3 ssa (1 - g) / 4 = u - u g with u = 0.75 ssa, (2 - 3 mu0 g) / 4 = 0.5 - (0.75 mu0) g
    const FT u = FT(0.75) * ssa, c3 = FT(0.75) * mu0;
    FT g1 = FT(2) - ssa * (FT(1.25) + FT(0.75) * g);
    FT g2 = u - u * g;
    FT g3 = FT(0.5) - c3 * g;
    FT g4 = FT(0.5) + c3 * g;
    FT a1 = g1 * g4 + g2 * g3;
    FT a2 = g1 * g3 + g2 * g4;
    FT k = hsqrt(rmax((FT(2) - FT(2) * ssa) * (g1 + g2), Num<FT>::k_min()));
    FT tk = tau * k;
    FT e = hexp(-tk);
    FT e2 = e * e;
    FT om1 = h_one_minus_exp_neg(tk, e);
    FT one_minus_e2kt = om1 * (FT(1) + e);
    FT one_plus_e2kt = FT(1) + e2;
    FT RT_term = hrcp(k * one_plus_e2kt + g1 * one_minus_e2kt);
    Rdif = RT_term * g2 * one_minus_e2kt;
    Tdif = RT_term * FT(2) * k * e;
    FT T0 = hexp(-tau * inv_mu0);                       // inv_mu0 = 1 / max(mu0, eps) (Numerics.jl:63)
    FT k_mu = k * mu0;
    FT k_mu2 = k_mu * k_mu;
    FT diff = FT(1) - k_mu2;
    const FT win = Num<FT>::k_min();                    // resonance_window = sqrt(eps) (Numerics.jl:49)
    {   // nudge k mu0 off the removable singularity (branch-free select; the two possible roots are constants):
        // k_mu2 = 1 -+ window, so the denominator 1 - k_mu2 below is +- window itself
        const bool res = rabs(diff) < win;
        const bool below = diff >= FT(0);
        const FT root_lo = rsqrt_(FT(1) - win), root_hi = rsqrt_(FT(1) + win);
        diff = res ? (below ? FT(1) - (FT(1) - win) : FT(1) - (FT(1) + win)) : diff;
        k_mu = res ? (below ? root_lo : root_hi) : k_mu;
    }
    FT k_g3 = k * g3, k_g4 = k * g4;
    RT_term = hdiv(ssa * RT_term, diff);
    // shortwave_2stream.jl:239-256 regrouped: with U = a2 - k_mu k_g3, S = k_g3 - a2 k_mu one has
    // (1 - k_mu)(a2 + k_g3) = U + S and (1 + k_mu)(a2 - k_g3) = U - S, hence
    //   Rdir = RT [U (1 - e^2) + S (1 + e^2 - 2 e T0)],
    // and with V = a1 + k_mu k_g4, F = k_g4 + a1 k_mu: (1 + k_mu)(a1 + k_g4) = V + F, (1 - k_mu)(a1 - k_g4) = V - F,
    //   Tdir = -RT [V T0 (1 - e^2) + F (T0 (1 + e^2) - 2 e)]
    // -- the same polynomial in the same variables, 9 operations fewer, and 1 - e^2 comes from the series-accurate
    // 1 - e (one_minus_e2kt) instead of a difference of O(1) terms
    const FT U = a2 - k_mu * k_g3, S = k_g3 - a2 * k_mu;
    const FT V = a1 + k_mu * k_g4, F = k_g4 + a1 * k_mu;
    FT Rdir_u = RT_term * (U * one_minus_e2kt + S * (one_plus_e2kt - FT(2) * (e * T0)));
    FT Tdir_u = -RT_term * (V * (T0 * one_minus_e2kt) + F * (T0 * one_plus_e2kt - FT(2) * e));
    Rdir = rmax0(Rdir_u);
    Tdir = rmax0(Tdir_u);
    FT av_energy = rmax0(FT(1) - T0);
    FT tot_dir = Rdir + Tdir;
    {
        FT scale = hdiv(av_energy, rmax(Num<FT>::eps(), tot_dir));
        scale = tot_dir > av_energy ? scale : FT(1);
        Rdir *= scale; Tdir *= scale;
    }
}

// ---- longwave_noscat.jl:171-205 ----
template <typename FT> __device__ __forceinline__ FT lw_noscat_source(FT lev_source, FT lay_source, FT tau_loc, FT trans) {
    FT fact = (tau_loc > Num<FT>::tau_thresh())
                  ? (hdiv(FT(1) - trans, tau_loc) - trans)
                  : tau_loc * (FT(0.5) + tau_loc * (-FT(1.0 / 3.0) + tau_loc * FT(0.125)));
    return (FT(1) - trans) * lev_source + FT(2) * fact * (lay_source - lev_source);
}

// ---- counter-based McICA uniforms shared bit-for-bit with the CPU oracle ----
// (the reference draws Random.rand(), cloud_optics.jl:253-262,283,293; see DESIGN.md "McICA")
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t mcica_col_key(uint64_t seed, uint64_t gcol0) {
    return splitmix64(splitmix64(seed) ^ (gcol0 * 0xD1B54A32D192ED03ULL));
}
__host__ __device__ __forceinline__ double mcica_rand(uint64_t col_key, int sw, int igpt, int ilay) {
    uint64_t h = splitmix64(col_key ^ ((uint64_t)sw << 40) ^ ((uint64_t)igpt << 16) ^ (uint64_t)ilay);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

}  // namespace rb

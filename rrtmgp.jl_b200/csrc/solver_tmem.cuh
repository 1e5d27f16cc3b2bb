// Blackwell-native variant of the two-stream kernels (Float32, nlay <= 64): the per-level
// recurrence state of the 32 (column, g-point) problems a warp owns lives in TENSOR MEMORY
// instead of shared memory.
//
// Why.  The adding method needs, for every level, the values produced by the upward sweep
// when the downward sweep passes the same level: 3 values x 64 levels x 4 B = 768 B per lane,
// 24.6 KB per warp.  In shared memory that caps residency at 4-6 warps per SM and the kernel
// is latency bound (profiles/r1a_*).  B200's 256 KB of TMEM per SM are otherwise idle here
// (no MMA on this path), and tcgen05.ld/st give every thread a private 512-word column
// array of its own TMEM lane -- exactly the access pattern of this store.  With two of the
// three values in TMEM (128 columns per 4-warp CTA) and one in shared memory, 12 warps are
// resident per SM.
//
// The g-point reduction cannot use TMEM (no cross-lane access); partial sums go through warp
// shuffles and land in per-lane broadband accumulators (lane = level), as in solver.cuh.
//
// SW uses the adding method marched from the TOP (reflectance/source of everything ABOVE a
// level), algebraically identical to shortwave_2stream.jl:300-392 but needing only one
// optics sweep before the recurrence can start, because the direct beam is also marched
// from the top; see DESIGN.md "SW adding from the top".
#pragma once
#include "solver.cuh"

namespace rb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(a)) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
    uint32_t x, y;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(taddr) : "memory");
    a = __uint_as_float(x); b = __uint_as_float(y);
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, float& a) {
    uint32_t x;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(taddr) : "memory");
    a = __uint_as_float(x);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// add `v` (already summed over lanes) to the accumulator of level `lev` on the lane that owns it
__device__ __forceinline__ void acc_add(float (&acc)[2], int lane, int lev, float v) {
    if ((lev & 31) == lane) {
        if (lev < 32) acc[0] += v; else acc[1] += v;
    }
}

constexpr int kTmemWarpsPerCta = 4;

// ALPHA_TMEM = true : all three values in TMEM (256 columns per CTA, 8 warps / SM)
// ALPHA_TMEM = false: (A, B) in TMEM (128 columns per CTA), albedo in shared memory (12 warps / SM)
template <int MODE, bool ALPHA_TMEM>
__global__ void __launch_bounds__(kTmemWarpsPerCta * 32, ALPHA_TMEM ? 2 : 3) solve_kernel_tmem(const SolveParams<float> P) {
    using FT = float;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_smem;
    constexpr bool LW = MODE != MODE_SW_2STREAM;
    constexpr uint32_t NCOLS = ALPHA_TMEM ? 256u : 128u;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long col = (long long)blockIdx.x * kTmemWarpsPerCta + warp;
    const bool active = col < P.ncol;   // warp-uniform

    if (warp == 0) tmem_alloc(&tmem_base_smem, NCOLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_smem + ((uint32_t)(warp & 3) << 21);   // lane field = 32 * (warp % 4), bits 31:16

    if (active) {
        unsigned char* wbase = smem_raw + (size_t)warp * P.warp_bytes;
        Warp<FT, MODE, 2> W(P, wbase, lane, col);
        const GasLut<FT>& L = P.lut;
        const int nlay = P.nlay, nlev = nlay + 1, n_gpt = L.n_gpt;
        FT* alpha_s = reinterpret_cast<FT*>(wbase + P.off_store);   // [nlay][32] (ALPHA_TMEM == false)
        const bool use_cloud = P.use_cloud != 0;

        W.phase0();

        const uint64_t col_key = mcica_col_key(P.seed, (uint64_t)(P.col_offset + col));
        int cld_start = 0, cld_finish = 0;
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

        FT acc_up[2] = {0.f, 0.f}, acc_dn[2] = {0.f, 0.f}, acc_dir[2] = {0.f, 0.f};   // levels lane, lane+32; level 64 below
        FT top_up = 0.f, top_dn = 0.f, top_dir = 0.f;                                  // level nlay when nlay == 64 (lane 0)
        int n_cloudy = 0;

        auto add_level = [&](FT (&acc)[2], FT& top, int lev, FT v) {
            if (lev < 64) acc_add(acc, lane, lev, v);
            else if (lane == 0) top += v;
        };

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
            if (!day) continue;

            const int gpt = W.gpt, ibnd = W.ibnd, bl = W.bl;
            const FT on = W.lane_on ? 1.f : 0.f;   // lanes past the last g-point contribute nothing

            if (LW) {
                // compute_optical_props.jl:157-195 sources + longwave_2stream.jl:243-334 adding (from the bottom)
                const FT* pb = W.plk + bl * 2 * nlev;
                const FT emis = __ldg(P.io.sfc_emis + (size_t)col * L.n_bnd + ibnd);
                const FT inc = P.io.inc_flux_lw ? __ldg(P.io.inc_flux_lw + (size_t)gpt * P.ncol + col) : FT(0);
                FT tau, ssa, g, pf;
                W.optics(0, tau, ssa, g, pf);
                FT lev_bot = pb[0] * pf;
                FT albedo = FT(1) - emis;
                FT src = Num<FT>::pi() * emis * (pb[nlev + nlay] * pf);
                for (int k = 0; k < nlay; ++k) {
                    FT tau_n = FT(0), ssa_n = FT(0), g_n = FT(0), pf_n = FT(0), lev_top;
                    FT inc_k = pb[k + 1] * pf;
                    if (k + 1 < nlay) {
                        W.optics(k + 1, tau_n, ssa_n, g_n, pf_n);
                        lev_top = hsqrt(inc_k * (pb[k + 1] * pf_n));
                    } else {
                        lev_top = inc_k;
                    }
                    FT Rdif, Tdif, su, sd;
                    lw_2stream_coeffs(tau, ssa, g, lev_bot, lev_top, Rdif, Tdif, su, sd);
                    FT denom = hdiv(FT(1), FT(1) - Rdif * albedo);
                    // level k: F_dn(k) = A_k F_dn(k+1) + B_k ; F_up(k) = albedo_k F_dn(k) + src_k
                    tmem_st2(tbase + 2 * k, Tdif * denom, (Rdif * src + sd) * denom);
                    if (ALPHA_TMEM) tmem_st1(tbase + 128 + k, albedo); else alpha_s[k * 32 + lane] = albedo;
                    add_level(acc_up, top_up, k, warp_sum(src * on));          // sum_g src_k
                    FT albedo_n = Rdif + Tdif * Tdif * albedo * denom;
                    src = su + Tdif * denom * (src + albedo * sd);
                    albedo = albedo_n;
                    lev_bot = lev_top; tau = tau_n; ssa = ssa_n; g = g_n; pf = pf_n;
                }
                FT dn = inc;
                add_level(acc_up, top_up, nlay, warp_sum((dn * albedo + src) * on));
                add_level(acc_dn, top_dn, nlay, warp_sum(dn * on));
                tmem_wait_st();
                FT A, B, al = 0.f;
                tmem_ld2(tbase + 2 * (nlay - 1), A, B);
                if (ALPHA_TMEM) tmem_ld1(tbase + 128 + nlay - 1, al);
                tmem_wait_ld();
                for (int k = nlay - 1; k >= 0; --k) {
                    const FT Ak = A, Bk = B;
                    const FT alk = ALPHA_TMEM ? al : alpha_s[k * 32 + lane];
                    if (k > 0) {   // prefetch the next level while this one is reduced
                        tmem_ld2(tbase + 2 * (k - 1), A, B);
                        if (ALPHA_TMEM) tmem_ld1(tbase + 128 + k - 1, al);
                    }
                    dn = Ak * dn + Bk;
                    add_level(acc_dn, top_dn, k, warp_sum(dn * on));
                    add_level(acc_up, top_up, k, warp_sum(alk * dn * on));
                    tmem_wait_ld();
                }
            } else {
                // shortwave_2stream.jl:300-392 with the adding marched from the top
                const FT alb_dir = __ldg(P.io.sfc_alb_direct + (size_t)col * L.n_bnd + ibnd);
                const FT alb_dif = __ldg(P.io.sfc_alb_diffuse + (size_t)col * L.n_bnd + ibnd);
                const FT dir_top = toa * __ldg(L.solar_src_scaled + gpt) * mu0;
                const FT inv_mu0 = FT(1) / rmax(mu0, Num<FT>::eps());
                FT tau_cum = FT(0), dir = dir_top;
                FT beta = FT(0), d = FT(0);   // reflectance / downward diffuse source of everything above the level
                {
                    FT s = warp_sum(dir_top * on);
                    add_level(acc_dir, top_dir, nlay, s);
                    add_level(acc_dn, top_dn, nlay, s);     // diffuse incident flux is zero (shortwave_2stream.jl:331)
                }
                for (int k = nlay - 1; k >= 0; --k) {
                    FT tau, ssa, g, pf;
                    W.optics(k, tau, ssa, g, pf);
                    FT Rdir, Tdir, Rdif, Tdif;
                    sw_2stream_coeffs(tau, ssa, g, mu0, inv_mu0, Rdir, Tdir, Rdif, Tdif);
                    const FT su = Rdir * dir, sd = Tdir * dir;       // direct-beam sources (dir = direct flux at level k+1)
                    const FT denom = hdiv(FT(1), FT(1) - Rdif * beta);
                    // F_up(k+1) = A'_k F_up(k) + B'_k ; F_dn_dif(k+1) = beta_{k+1} F_up(k+1) + d_{k+1}
                    tmem_st2(tbase + 2 * k, Tdif * denom, (Rdif * d + su) * denom);
                    if (ALPHA_TMEM) tmem_st1(tbase + 128 + k, beta); else alpha_s[k * 32 + lane] = beta;
                    if (k < nlay - 1) add_level(acc_dn, top_dn, k + 1, warp_sum(d * on));   // d_{nlay} = 0
                    d = sd + Tdif * denom * (d + beta * su);
                    beta = Rdif + Tdif * Tdif * beta * denom;
                    tau_cum += tau;
                    dir = dir_top * hexp(-tau_cum * inv_mu0);         // direct flux at level k
                    FT s = warp_sum(dir * on);
                    add_level(acc_dir, top_dir, k, s);
                    add_level(acc_dn, top_dn, k, s);
                }
                // surface: F_up(0) = alb_dif F_dn_dif(0) + alb_dir dir(0) ; F_dn_dif(0) = d_0 + beta_0 F_up(0)
                FT up = hdiv(alb_dif * d + alb_dir * dir, FT(1) - alb_dif * beta);
                add_level(acc_up, top_up, 0, warp_sum(up * on));
                add_level(acc_dn, top_dn, 0, warp_sum((d + beta * up) * on));
                tmem_wait_st();
                FT A, B, be = 0.f;
                tmem_ld2(tbase, A, B);
                if (ALPHA_TMEM) tmem_ld1(tbase + 128, be);
                tmem_wait_ld();
                for (int k = 0; k < nlay; ++k) {
                    const FT Ak = A, Bk = B;
                    const FT bek = ALPHA_TMEM ? be : alpha_s[k * 32 + lane];
                    if (k + 1 < nlay) {
                        tmem_ld2(tbase + 2 * (k + 1), A, B);
                        if (ALPHA_TMEM) tmem_ld1(tbase + 128 + k + 1, be);
                    }
                    up = Ak * up + Bk;                                  // F_up(k+1)
                    add_level(acc_up, top_up, k + 1, warp_sum(up * on));
                    add_level(acc_dn, top_dn, k + 1, warp_sum(bek * up * on));
                    tmem_wait_ld();
                }
            }
        }

        // ---------------- epilogue ----------------
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int lev = lane + 32 * i;
            if (lev < nlev && (i < 2 || lane == 0)) {
                const size_t o = (size_t)col * nlev + lev;
                FT up = i < 2 ? acc_up[i < 2 ? i : 0] : top_up;
                FT dn = i < 2 ? acc_dn[i < 2 ? i : 0] : top_dn;
                FT dr = i < 2 ? acc_dir[i < 2 ? i : 0] : top_dir;
                if (!day) { up = dn = dr = FT(0); }
                FT net = up - dn;
                if (P.io.metric_scaling != nullptr) {
                    FT sc = __ldg(P.io.metric_scaling + o);
                    up *= sc; dn *= sc; net *= sc; dr *= sc;
                }
                P.io.out_up[o] = up; P.io.out_dn[o] = dn; P.io.out_net[o] = net;
                if (!LW) P.io.out_dir[o] = dr;
                if (P.io.out_total_net != nullptr) P.io.out_total_net[o] = P.io.add_net[o] + net;
            }
        }
        if (lane == 0 && P.io.cld_cover != nullptr && use_cloud) P.io.cld_cover[col] = FT(n_cloudy) / FT(n_gpt);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_smem, NCOLS);
}

}  // namespace rb

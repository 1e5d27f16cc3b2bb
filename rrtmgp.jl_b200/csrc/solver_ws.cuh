// Warp-specialised two-stream kernels (round 2): the fused column solve of solver_fast.cuh split into two ROLES that
// run as a pipeline inside one persistent CTA of 16 warps (4 per scheduler instead of 3):
//
//   gas warps (8..15)   everything that ends in (tau, ssa, g[, Planck source]) of a (layer, g-point) cell: phase 0 /
//                       phase 1 / McICA of solver.cuh, the k-distribution corner gathers and the trilinear
//                       interpolation (gas_optics.jl:176-320, optics_utils.jl:85-181), the cloud / aerosol increment
//                       and -- longwave -- the level sources (compute_optical_props.jl:157-195).  Layers are
//                       independent here, so the gathers of layer k+1 are in flight while layer k is interpolated.
//   RT warps (0..7)     the two-stream coefficients and the adding recurrences (longwave_2stream.jl:149-334,
//                       shortwave_2stream.jl:189-392), the g-point reductions and the (nlev, ncol) epilogue.  Their
//                       level store is tensor memory only: 2 warps per TMEM lane quadrant, 256 columns each, so all
//                       three values of all 64 levels fit and the shared-memory albedo spill of solver_fast.cuh is gone.
//
// Gas warp 8 + i feeds RT warp i through a shared-memory ring: stage = the (up to) four layers 4s..4s+3 of the block
// in hand, one float4 per (layer, lane) + one header row, `kWsStages` stages deep, full / empty mbarriers (all 32
// lanes arrive).  The column index travels through a two-deep mailbox with its own mbarrier pair; gas warps own the
// atomic column queue.  Registers are re-balanced with setmaxnreg (RT 104, gas 152 = the whole register file).
//
// Why not TMA for the corner gathers (north star): measured on B200
// (profiles/r2a_tma_gather_micro.txt) one SM retires one cp.async.bulk / cp.async.bulk.tensor request per ~13 clocks
// whatever its size up to 512 B, so the 8-12 requests a (layer, 32 g-points) step needs take 105-135 clocks against
// 69 clocks for the same 4 KB through per-lane LDG -- and the whole fused step took 113.  TMA stays where it fits: the
// one-off bulk staging of the small tables (solver_fast.cuh).
#pragma once
#include "solver_fast.cuh"

namespace rb {

constexpr int kWsPairs = 8;          // (gas warp, RT warp) pairs per CTA
constexpr int kWsHL = 4;             // layers per hand-off stage
constexpr int kWsStages = 3;         // ring depth in stages
constexpr int kWsMaxLay = 64;
constexpr int kWsAccStride = kWsMaxLay + 4;
constexpr int kWsStageF4 = (kWsHL + 1) * 32;   // float4 per stage: four layer rows + one header row
constexpr int kWsRegsRT = 104, kWsRegsGas = 152;

struct WsSmem {
    int off_ring, off_stage, off_acc, off_bacc;   // byte offsets from the pair's base (after the gas warp's Warp<> layout)
    int off_blob, off_vmr, staged_bytes;          // CTA-shared tail
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// table corners + the two record groups that carry the table offsets of one (layer, g-point) cell
template <bool LWG, int NG> struct GasLoads {
    float2 c2[LWG ? 8 : 1];
    float c1[LWG ? 1 : 8];
    float4 m[4 * NG];
    float4 s, x;
};

template <int MODE, int NGPT, int NG, bool HAS_CLD, bool HAS_AER, bool SPECTRAL>
__global__ void __launch_bounds__(kWsPairs * 64, 1) solve_kernel_ws(const SolveParams<float> P, const WsSmem F) {
    static_assert(MODE == MODE_LW_2STREAM || MODE == MODE_SW_2STREAM, "two-stream modes only");
    using FT = float;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(8) uint64_t blob_bar;
    // per pair: full[kWsStages], empty[kWsStages], column mailbox full[2], empty[2]
    __shared__ __align__(8) uint64_t bars[kWsPairs][2 * kWsStages + 4];
    __shared__ long long colslot[kWsPairs][2];
    constexpr bool LW = MODE == MODE_LW_2STREAM;
    constexpr bool INCR = HAS_CLD || HAS_AER;
    constexpr int NETA = 9, NT = 14;
    constexpr int KE = NGPT, KT = NETA * KE, KP = NT * KT;           // major-table strides (LW: in float2): eta, T, p
    constexpr int ME = NGPT, MT = NETA * NGPT, MS = NT * MT;         // minor-table strides in float4: eta, T, group
    constexpr int UP = 0, DN = 1, DIR = 2;
    constexpr int kAcc = kWsAccStride;
    (void)KT; (void)MT;

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (solver_fast.cuh)
    const int pair = warp & (kWsPairs - 1);
    const bool is_rt = warp < kWsPairs;

    unsigned char* sblob = smem_raw + F.off_blob;
    float* svmr = P.vmr_kind == 0 ? reinterpret_cast<float*>(smem_raw + F.off_vmr) : nullptr;
    if (threadIdx.x == 0) {
        mbar_init(&blob_bar, 1);
        mbar_expect_tx(&blob_bar, (uint32_t)F.staged_bytes);
        if (F.staged_bytes > 0) tma_bulk_g2s(sblob, P.lut.blob, (uint32_t)F.staged_bytes, &blob_bar);
    }
    if (threadIdx.x < kWsPairs * (2 * kWsStages + 4)) {
        const int b = threadIdx.x % (2 * kWsStages + 4);
        mbar_init(&bars[threadIdx.x / (2 * kWsStages + 4)][b], b < 2 * kWsStages ? 32u : 1u);
    }
    if (svmr != nullptr)
        for (int i = threadIdx.x; i < P.ngas; i += blockDim.x) svmr[i] = __ldg(P.io.vmr + i);
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(&blob_bar, 0);

    unsigned char* pbase = smem_raw + (size_t)pair * P.warp_bytes;
    float4* ring = reinterpret_cast<float4*>(pbase + F.off_ring);    // [kWsStages][kWsHL + 1][32]
    uint64_t* bar_full = &bars[pair][0];
    uint64_t* bar_empty = &bars[pair][kWsStages];
    uint64_t* col_full = &bars[pair][2 * kWsStages];
    uint64_t* col_empty = &bars[pair][2 * kWsStages + 2];
    const GasLut<FT>& L = P.lut;
    const int nlay = P.nlay, nlev = nlay + 1;
    // a small table: its copy inside the staged prefix of the block, or the global array
    auto tb = [&](const int* g) -> const int* {
        const long long off = reinterpret_cast<const unsigned char*>(g) - L.blob;
        return (off >= 0 && off < F.staged_bytes) ? reinterpret_cast<const int*>(sblob + off) : g;
    };
    int ring_idx = 0;            // stage of the ring in hand and its phase parity; both roles step them identically
    uint32_t ring_phase = 0;
    auto ring_next = [&]() { if (++ring_idx == kWsStages) { ring_idx = 0; ring_phase ^= 1u; } };

    if (is_rt) {
        // =====================================================================================================
        // RT role: coefficients + adding + reductions + epilogue
        // =====================================================================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsRegsRT));
        const uint32_t tA = tmem_base_smem + ((uint32_t)(warp & 3) << 21) + (uint32_t)((warp >> 2) * 256);
        const uint32_t tAl = tA + 2u * kWsMaxLay;
        FT* stage = reinterpret_cast<FT*>(pbase + F.off_stage);      // [16][kStageStride]
        FT* accs = reinterpret_cast<FT*>(pbase + F.off_acc);         // [3][kAcc]
        constexpr bool spectral = SPECTRAL;
        FT* bacc = reinterpret_cast<FT*>(pbase + (F.off_bacc >= 0 ? F.off_bacc : 0));   // [2][2][kAcc]
        const int hb = lane >> 4;
        const int* gpt2bnd = tb(L.gpt2bnd);
        auto row_sum = [&](FT& half) -> FT {
            const float4* row = reinterpret_cast<const float4*>(stage + (lane & 15) * kStageStride + (lane >> 4) * 16);
            const float4 a = row[0], b = row[1], c = row[2], d = row[3];
            half = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)) + (((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)));
            return half + __shfl_xor_sync(0xffffffffu, half, 16);
        };
        auto warp_sum2 = [&](FT v, FT& half) -> FT {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            half = v;
            return v + __shfl_xor_sync(0xffffffffu, v, 16);
        };
        auto band_add = [&](int q, int lev, FT half) { bacc[(hb * 2 + q) * kAcc + lev] += half; };

        for (unsigned ncolumn = 0;; ++ncolumn) {
            mbar_wait(&col_full[ncolumn & 1], (ncolumn >> 1) & 1);
            const long long col = colslot[pair][ncolumn & 1];
            __syncwarp();
            if (lane == 0) mbar_arrive(&col_empty[ncolumn & 1]);
            if (col < 0) break;
            for (int i = lane; i < 3 * kAcc; i += 32) accs[i] = FT(0);
            if (spectral)
                for (int i = lane; i < 4 * kAcc; i += 32) bacc[i] = FT(0);
            const FT mu0 = LW ? FT(1) : __ldg(P.io.cos_zenith + col);
            const bool day = LW || mu0 > FT(0);
            const FT toa = LW ? FT(0) : __ldg(P.io.toa_flux + col);
            __syncwarp();

            for (int g0 = 0; g0 < NGPT; g0 += 32) {
                const int gpt = g0 + lane;
                const int ibnd = gpt2bnd[gpt], b_first = gpt2bnd[g0], nb = gpt2bnd[g0 + 31] - b_first + 1;
                auto flush_bands = [&](bool zero) {
                    __syncwarp();
                    for (int b = 0; b < nb; ++b) {
                        const size_t ob = ((size_t)(b_first + b) * P.ncol_total + col) * nlev;
                        for (int lev = lane; lev < nlev; lev += 32) {
                            FT bu = zero ? 0.f : bacc[(b * 2 + UP) * kAcc + lev];
                            FT bd = zero ? 0.f : bacc[(b * 2 + DN) * kAcc + lev];
                            bacc[(b * 2 + UP) * kAcc + lev] = 0.f; bacc[(b * 2 + DN) * kAcc + lev] = 0.f;
                            if (P.io.metric_scaling != nullptr) { const FT sc = __ldg(P.io.metric_scaling + (size_t)col * nlev + lev); bu *= sc; bd *= sc; }
                            P.io.band_up[ob + lev] = bu; P.io.band_dn[ob + lev] = bd; P.io.band_net[ob + lev] = bu - bd;
                        }
                    }
                    __syncwarp();
                };
                if (!day) {   // night: exactly zero (shortwave_2stream.jl:169-175); the gas warp sends nothing
                    if (spectral) flush_bands(true);
                    continue;
                }
                if (LW) {
                    // longwave_2stream.jl:243-334, adding from the bottom; the hand-off row of layer k is
                    // (tau, ssa, g, Planck source at the layer's top level), the header (source at level 0, surface Planck)
                    const FT emis = __ldg(P.io.sfc_emis + (size_t)col * L.n_bnd + ibnd);
                    const FT inc = P.io.inc_flux_lw ? __ldg(P.io.inc_flux_lw + (size_t)gpt * P.ncol_total + col) : 0.f;
                    FT lev_bot = 0.f, albedo = 1.f - emis, src = 0.f;
                    for (int t0 = 0; t0 < nlay; t0 += 16) {                  // tiles of <= 16 layers = <= 4 stages
                        const int tend = t0 + 16 < nlay ? t0 + 16 : nlay;
                        for (int s0 = t0; s0 < tend; s0 += kWsHL) {
                            mbar_wait(&bar_full[ring_idx], ring_phase);
                            const float4* slot = ring + ring_idx * kWsStageF4 + lane;
                            if (s0 == 0) {
                                const float4 h = slot[kWsHL * 32];
                                lev_bot = h.x;
                                src = Num<FT>::pi() * emis * h.y;
                            }
                            const int send = s0 + kWsHL < nlay ? s0 + kWsHL : nlay;
#pragma unroll 1
                            for (int k = s0; k < send; ++k) {
                                const float4 v = slot[(k - s0) * 32];
                                const LwCoef C = lw_2stream_coeffs_nosrc(v.x, v.y, v.z);
                                const FT denom = rcp_approx(1.f - C.Rdif * albedo);
                                const FT lev_top = v.w;
                                const FT dB = lev_bot - lev_top;
                                const FT su = Num<FT>::pi() * (lev_top * C.emis_fac - C.q * dB);
                                const FT sd = Num<FT>::pi() * (lev_bot * C.emis_fac + C.q * dB);
                                // level k: F_dn(k) = A F_dn(k+1) + B ; F_up(k) = albedo F_dn(k) + src
                                tmem_st2(tA + 2 * k, C.Tdif * denom, (C.Rdif * src + sd) * denom);
                                tmem_st1(tAl + k, albedo);
                                stage[(k - t0) * kStageStride + lane] = src;
                                src = su + C.Tdif * denom * (src + albedo * sd);
                                albedo = C.Rdif + C.Tdif * C.Tdif * albedo * denom;
                                lev_bot = lev_top;
                            }
                            mbar_arrive(&bar_empty[ring_idx]);
                            ring_next();
                        }
                        __syncwarp();
                        {                                                     // sum_g src of levels t0 .. tend-1
                            FT hs;
                            const FT sum = row_sum(hs);
                            if (lane < 16 && lane < tend - t0) accs[UP * kAcc + t0 + lane] += sum;
                            if (spectral && (lane & 15) < tend - t0) band_add(UP, t0 + (lane & 15), hs);
                        }
                        __syncwarp();
                    }
                    FT dn = inc;
                    {
                        FT hu, hd;
                        FT u = warp_sum2(dn * albedo + src, hu), d = warp_sum2(dn, hd);
                        if (lane == 0) { accs[UP * kAcc + nlay] += u; accs[DN * kAcc + nlay] += d; }
                        if (spectral && (lane & 15) == 0) { band_add(UP, nlay, hu); band_add(DN, nlay, hd); }
                    }
                    tmem_wait_st();
                    for (int kc = (nlay - 1) & ~7; kc >= 0; kc -= 8) {     // 8 levels x (dn, albedo * dn) per tile
                        const int ktop = kc + 7 < nlay - 1 ? kc + 7 : nlay - 1;
                        if (ktop == kc + 7) {
                            float ab[16], al8[8];
                            tmem_ld16(tA + 2 * kc, ab);
                            tmem_ld8(tAl + kc, al8);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 7; i >= 0; --i) {
                                dn = ab[2 * i] * dn + ab[2 * i + 1];
                                stage[(i * 2 + 0) * kStageStride + lane] = dn;
                                stage[(i * 2 + 1) * kStageStride + lane] = al8[i] * dn;
                            }
                        } else {
                            for (int k = ktop; k >= kc; --k) {
                                FT A, B, al;
                                tmem_ld2(tA + 2 * k, A, B);
                                tmem_ld1(tAl + k, al);
                                tmem_wait_ld();
                                dn = A * dn + B;
                                stage[((k - kc) * 2 + 0) * kStageStride + lane] = dn;
                                stage[((k - kc) * 2 + 1) * kStageStride + lane] = al * dn;
                            }
                        }
                        __syncwarp();
                        {
                            const int lev = kc + ((lane & 15) >> 1);
                            FT hs;
                            const FT sum = row_sum(hs);
                            if (lane < 16 && lev <= ktop) accs[((lane & 1) ? UP : DN) * kAcc + lev] += sum;
                            if (spectral && lev <= ktop) band_add((lane & 1) ? UP : DN, lev, hs);
                        }
                        __syncwarp();
                    }
                } else {
                    // shortwave_2stream.jl:300-392 with the adding marched from the top (solver_fast.cuh / DESIGN.md);
                    // the hand-off row of layer k is (tau, ssa, g, -)
                    const FT alb_dir = __ldg(P.io.sfc_alb_direct + (size_t)col * L.n_bnd + ibnd);
                    const FT alb_dif = __ldg(P.io.sfc_alb_diffuse + (size_t)col * L.n_bnd + ibnd);
                    const FT dir_top = toa * __ldg(L.solar_src_scaled + gpt) * mu0;
                    const FT inv_mu0 = hdiv(1.f, rmax(mu0, FLT_EPSILON));
                    const FT neg_inv_mu0_l2e = -inv_mu0 * 1.4426950408889634f;
                    FT tau_cum = 0.f, dir = dir_top;
                    FT beta = 0.f, d = 0.f;   // reflectance / downward diffuse source of everything above the level
                    {
                        FT hs;
                        FT sum = warp_sum2(dir_top, hs);   // TOA: diffuse incident flux is zero (shortwave_2stream.jl:331)
                        if (lane == 0) { accs[DIR * kAcc + nlay] += sum; accs[DN * kAcc + nlay] += sum; }
                        if (spectral && (lane & 15) == 0) band_add(DN, nlay, hs);
                    }
                    for (int kc = (nlay - 1) & ~7; kc >= 0; kc -= 8) {     // tiles of <= 8 layers = <= 2 stages, top down
                        const int ktop = kc + 7 < nlay - 1 ? kc + 7 : nlay - 1;
                        for (int shi = ktop; shi >= kc; shi = (shi & ~3) - 1) {
                            const int slo = shi & ~3;
                            mbar_wait(&bar_full[ring_idx], ring_phase);
                            const float4* slot = ring + ring_idx * kWsStageF4 + lane;
#pragma unroll 1
                            for (int k = shi; k >= slo; --k) {
                                const float4 v = slot[(k & 3) * 32];
                                FT Rdir, Tdir, Rdif, Tdif;
                                sw_2stream_coeffs(v.x, v.y, v.z, mu0, inv_mu0, Rdir, Tdir, Rdif, Tdif);
                                const FT su = Rdir * dir, sd = Tdir * dir;       // dir = direct flux at level k+1
                                const FT denom = rcp_approx(1.f - Rdif * beta);
                                // F_up(k+1) = A'_k F_up(k) + B'_k ; F_dn_dif(k+1) = beta_{k+1} F_up(k+1) + d_{k+1}
                                tmem_st2(tA + 2 * k, Tdif * denom, (Rdif * d + su) * denom);
                                tmem_st1(tAl + k, beta);
                                stage[((k - kc) * 2 + 0) * kStageStride + lane] = d;        // d_{k+1}
                                d = sd + Tdif * denom * (d + beta * su);
                                beta = Rdif + Tdif * Tdif * beta * denom;
                                tau_cum += v.x;
                                float ex;
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(tau_cum * neg_inv_mu0_l2e));
                                dir = dir_top * ex;                               // direct flux at level k
                                stage[((k - kc) * 2 + 1) * kStageStride + lane] = dir;
                            }
                            mbar_arrive(&bar_empty[ring_idx]);
                            ring_next();
                        }
                        __syncwarp();
                        {
                            const int kk = kc + ((lane & 15) >> 1);
                            FT hs;
                            const FT sum = row_sum(hs);
                            const bool okb = kk <= ktop, ok = lane < 16 && okb;
                            // d_{kk+1} (even lanes) and dir_kk (odd lanes) both feed F_dn: two ordered steps
                            if (ok && (lane & 1)) { accs[DN * kAcc + kk] += sum; accs[DIR * kAcc + kk] += sum; }
                            if (spectral && okb && (lane & 1)) band_add(DN, kk, hs);
                            __syncwarp();
                            if (ok && !(lane & 1)) accs[DN * kAcc + kk + 1] += sum;
                            if (spectral && okb && !(lane & 1)) band_add(DN, kk + 1, hs);
                        }
                        __syncwarp();
                    }
                    // surface: F_up(0) = alb_dif F_dn_dif(0) + alb_dir dir(0) ; F_dn_dif(0) = d_0 + beta_0 F_up(0)
                    FT up = hdiv(alb_dif * d + alb_dir * dir, 1.f - alb_dif * beta);
                    {
                        FT hu, hdd;
                        FT u = warp_sum2(up, hu), dd = warp_sum2(d + beta * up, hdd);
                        if (lane == 0) { accs[UP * kAcc] += u; accs[DN * kAcc] += dd; }
                        if (spectral && (lane & 15) == 0) { band_add(UP, 0, hu); band_add(DN, 0, hdd); }
                    }
                    tmem_wait_st();
                    for (int kc = 0; kc < nlay; kc += 8) {                // 8 levels x (F_up, beta * F_up) per tile
                        const int kend = kc + 8 < nlay ? kc + 8 : nlay;
                        if (kend == kc + 8) {
                            float ab[16], be8[8];
                            tmem_ld16(tA + 2 * kc, ab);
                            tmem_ld8(tAl + kc, be8);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                up = ab[2 * i] * up + ab[2 * i + 1];              // F_up(k+1)
                                stage[(i * 2 + 0) * kStageStride + lane] = up;
                                stage[(i * 2 + 1) * kStageStride + lane] = be8[i] * up;
                            }
                        } else {
                            for (int k = kc; k < kend; ++k) {
                                FT A, B, be;
                                tmem_ld2(tA + 2 * k, A, B);
                                tmem_ld1(tAl + k, be);
                                tmem_wait_ld();
                                up = A * up + B;                                    // F_up(k+1)
                                stage[((k - kc) * 2 + 0) * kStageStride + lane] = up;
                                stage[((k - kc) * 2 + 1) * kStageStride + lane] = be * up;
                            }
                        }
                        __syncwarp();
                        {
                            const int kk = kc + ((lane & 15) >> 1);
                            FT hs;
                            const FT sum = row_sum(hs);
                            if (lane < 16 && kk < kend) accs[((lane & 1) ? DN : UP) * kAcc + kk + 1] += sum;
                            if (spectral && kk < kend) band_add((lane & 1) ? DN : UP, kk + 1, hs);
                        }
                        __syncwarp();
                    }
                }
                if (spectral) flush_bands(false);
            }
            __syncwarp();

            // ---------------- epilogue: (nlev, ncol) presentation, net, scaling ----------------
#pragma unroll
            for (int i = 0; i < (kWsMaxLay + 32) / 32; ++i) {
                const int lev = lane + 32 * i;
                if (lev < nlev) {
                    const size_t o = (size_t)col * nlev + lev;
                    FT up = accs[UP * kAcc + lev], dn = accs[DN * kAcc + lev], dr = accs[DIR * kAcc + lev];
                    if (!day) { up = dn = dr = 0.f; }
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
            __syncwarp();
        }
    } else {
        // =====================================================================================================
        // gas role: phase 0 / phase 1 / McICA, corner gathers, interpolation, increments, level sources
        // =====================================================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsRegsGas));
        const FT* major = LW ? L.kmaj_pf : L.kmajor;
        const float4* minor4 = reinterpret_cast<const float4*>(L.kminor4[0]);
        const int RW = P.rec_words;
        auto next_column = [&]() -> long long {
            unsigned int v = 0;
            if (lane == 0) v = atomicAdd(P.work_counter, 1u);
            return (long long)__shfl_sync(0xffffffffu, v, 0);
        };
        long long col_next = next_column();
        for (unsigned ncolumn = 0;; ++ncolumn) {
            const long long col = col_next;
            if (lane == 0) {   // column mailbox, two deep
                mbar_wait(&col_empty[ncolumn & 1], ((ncolumn >> 1) & 1) ^ 1u);
                colslot[pair][ncolumn & 1] = col < P.ncol ? col : -1;
                mbar_arrive(&col_full[ncolumn & 1]);
            }
            __syncwarp();
            if (col >= P.ncol) break;
            col_next = next_column();
            Warp<FT, MODE, 2, true> W(P, pbase, lane, col, sblob, F.staged_bytes, svmr);
            {
                const long long nc = col_next;
                if (nc < P.ncol) {   // the next column's inputs (read once, cold in DRAM) into L2
                    auto prefetch_row = [&](const FT* base, int n) {
                        if (base == nullptr) return;
                        const char* b = reinterpret_cast<const char*>(base + (size_t)nc * n);
                        const int bytes = n * (int)sizeof(FT);
                        for (int o = lane * 128; o < bytes + 127; o += 32 * 128)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(b + (o < bytes ? o : bytes - 1)));
                    };
                    prefetch_row(P.io.layerdata, 4 * nlay);
                    prefetch_row(P.io.t_lev, nlev);
                    if (P.vmr_kind == 0) { prefetch_row(P.io.vmr_h2o, nlay); prefetch_row(P.io.vmr_o3, nlay); }
                    else prefetch_row(P.io.vmr, nlay * P.ngas);
                    if (HAS_CLD) {
                        prefetch_row(P.io.cld_frac, nlay); prefetch_row(P.io.cld_path_liq, nlay); prefetch_row(P.io.cld_path_ice, nlay);
                        prefetch_row(P.io.cld_r_eff_liq, nlay); prefetch_row(P.io.cld_r_eff_ice, nlay);
                    }
                    if (HAS_AER) { prefetch_row(P.io.aero_mass, 15 * nlay); prefetch_row(P.io.aero_size, 15 * nlay); }
                }
            }
            W.phase0();
            const uint64_t col_key = mcica_col_key(P.seed, (uint64_t)(P.col_offset + col));
            int cld_start = 0, cld_finish = 0;
            if (HAS_CLD) {
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
            int n_cloudy = 0;
            __syncwarp();

            for (int g0 = 0; g0 < NGPT; g0 += 32) {
                W.set_block(g0);
                __syncwarp();
                if (HAS_CLD) n_cloudy += W.mcica(col_key, cld_start, cld_finish);
                FT aod_e = 0.f, aod_s = 0.f;
                const bool aod_here = !LW && HAS_AER && P.io.aod_ext != nullptr && P.aero.iband_550nm >= W.b_first + 1 &&
                                      P.aero.iband_550nm <= W.b_first + W.nb;
                auto build_records = [&](int part) {
                    __syncwarp();
                    FT e, sc;
                    W.phase1(e, sc, part);
                    aod_e += e; aod_s += sc;
                };
                if (!day) {   // night: AOD and masks only (shortwave_2stream.jl:66-102)
                    if (aod_here) {
                        for (int part = 0; part * 32 < nlay; ++part) build_records(part);
                        aod_e = warp_sum(aod_e); aod_s = warp_sum(aod_s);
                        if (lane == 0) { P.io.aod_ext[col] = aod_e; P.io.aod_sca[col] = aod_s; }
                    }
                    continue;
                }
                const int gpt = W.gpt, bl = W.bl;
                const FT* rec_lane = W.rec + bl * RW;
                const FT* major_lane = major + (LW ? 2 : 1) * gpt;
                const float4* minor_lane = minor4 + gpt;
                const unsigned mask0 = W.mask[0], mask1 = W.mask[1];

                // ---- issue the table gathers of cell (layer k, this g-point) ----
                auto issue = [&](int k, GasLoads<LW, NG>& G) {
                    const FT* r = rec_lane + (k & 31) * P.rec_row;
                    bool cb = false;
                    if (HAS_CLD) cb = ((k < 32 ? mask0 : mask1) >> (k & 31)) & 1u;
                    G.s = *reinterpret_cast<const float4*>(r + 8);
                    G.x = *reinterpret_cast<const float4*>(r + 12 + 4 * NG + (cb ? 4 : 0));
                    const int ia = __float_as_int(G.s.z), ib = __float_as_int(G.s.w);   // (jp-1, jt, je1), (jp-1, jt+1, je2)
                    const int ma = __float_as_int(G.x.w), mb = ma + (ib - ia);          // (jt, je1), (jt+1, je2): MT == KT
                    if (LW) {
                        const float2* pa = reinterpret_cast<const float2*>(major_lane) + ia;
                        const float2* pb = reinterpret_cast<const float2*>(major_lane) + ib;
                        G.c2[0] = __ldg(pa); G.c2[1] = __ldg(pa + KE); G.c2[2] = __ldg(pa + KP); G.c2[3] = __ldg(pa + KP + KE);
                        G.c2[4] = __ldg(pb); G.c2[5] = __ldg(pb + KE); G.c2[6] = __ldg(pb + KP); G.c2[7] = __ldg(pb + KP + KE);
                    } else {
                        const FT* pa = major_lane + ia;
                        const FT* pb = major_lane + ib;
                        G.c1[0] = __ldg(pa); G.c1[1] = __ldg(pa + KE); G.c1[2] = __ldg(pa + KP); G.c1[3] = __ldg(pa + KP + KE);
                        G.c1[4] = __ldg(pb); G.c1[5] = __ldg(pb + KE); G.c1[6] = __ldg(pb + KP); G.c1[7] = __ldg(pb + KP + KE);
                    }
#pragma unroll
                    for (int gi = 0; gi < NG; ++gi) {
                        G.m[4 * gi + 0] = __ldg(minor_lane + ma + gi * MS); G.m[4 * gi + 1] = __ldg(minor_lane + ma + gi * MS + ME);
                        G.m[4 * gi + 2] = __ldg(minor_lane + mb + gi * MS); G.m[4 * gi + 3] = __ldg(minor_lane + mb + gi * MS + ME);
                    }
                };
                // ---- gas + cloud + aerosol optics of the gathered cell (gas_optics.jl:176-320, optics_utils.jl:85-202) ----
                auto optics = [&](int k, const GasLoads<LW, NG>& G, FT& tau, FT& ssa, FT& g, FT& pfrac) {
                    const FT* r = rec_lane + (k & 31) * P.rec_row;
                    const float4 v0 = *reinterpret_cast<const float4*>(r), v1 = *reinterpret_cast<const float4*>(r + 4);
                    if (LW) {
                        const float2* c = G.c2;
                        tau = G.s.x * (v0.x * c[0].x + v0.y * c[1].x + v0.z * c[2].x + v0.w * c[3].x) +
                              G.s.y * (v1.x * c[4].x + v1.y * c[5].x + v1.z * c[6].x + v1.w * c[7].x);
                        pfrac = (v0.x * c[0].y + v0.y * c[1].y + v0.z * c[2].y + v0.w * c[3].y) +
                                (v1.x * c[4].y + v1.y * c[5].y + v1.z * c[6].y + v1.w * c[7].y);
                    } else {
                        const FT* c = G.c1;
                        tau = G.s.x * (v0.x * c[0] + v0.y * c[1] + v0.z * c[2] + v0.w * c[3]) +
                              G.s.y * (v1.x * c[4] + v1.y * c[5] + v1.z * c[6] + v1.w * c[7]);
                        pfrac = 0.f;
                    }
                    const FT w11 = v0.x + v0.z, w21 = v0.y + v0.w, w12 = v1.x + v1.z, w22 = v1.y + v1.w;
                    FT tau_ray = 0.f;
#pragma unroll
                    for (int gi = 0; gi < NG; ++gi) {
                        const float4 m11 = G.m[4 * gi], m21 = G.m[4 * gi + 1], m12 = G.m[4 * gi + 2], m22 = G.m[4 * gi + 3];
                        const float4 sc = *reinterpret_cast<const float4*>(r + 12 + 4 * gi);
                        const FT x0 = w11 * m11.x + w21 * m21.x + w12 * m12.x + w22 * m22.x;
                        const FT x1 = w11 * m11.y + w21 * m21.y + w12 * m12.y + w22 * m22.y;
                        const FT x2 = w11 * m11.z + w21 * m21.z + w12 * m12.z + w22 * m22.z;
                        const FT x3 = w11 * m11.w + w21 * m21.w + w12 * m12.w + w22 * m22.w;
                        if (!LW && gi == 0) {
                            tau_ray = x0 * sc.x;
                            tau += x1 * sc.y + x2 * sc.z + x3 * sc.w;
                        } else {
                            tau += x0 * sc.x + x1 * sc.y + x2 * sc.z + x3 * sc.w;
                        }
                    }
                    if (LW) {
                        tau = rmax(tau, 0.f);
                        ssa = 0.f; g = 0.f;
                    } else {
                        tau = rmax(tau + tau_ray, 0.f);
                        ssa = tau > 0.f ? hdiv(tau_ray, tau) : 0.f;
                        g = 0.f;
                    }
                    if (INCR) {   // one fused, unconditional increment (optics_utils.jl:189-202, additive form)
                        const FT tn = tau + G.x.x;
                        const FT w = LW ? G.x.y : tau * ssa + G.x.y;
                        const FT h = G.x.z;
                        g = hdiv(h, rmax(FLT_EPSILON, w));
                        ssa = hdiv(w, rmax(FLT_EPSILON, tn));
                        tau = tn;
                    }
                };
                // ---- hand layer k to the RT warp: slot (k & 3) of stage k >> 2 ----
                FT hdr0 = 0.f, hdr1 = 0.f;
                auto emit = [&](int k, FT a, FT b, FT c, FT d) {
                    const bool first = LW ? (k & 3) == 0 : ((k & 3) == 3 || k == nlay - 1);
                    const bool last = LW ? ((k & 3) == 3 || k == nlay - 1) : (k & 3) == 0;
                    float4* slot = ring + ring_idx * kWsStageF4 + lane;
                    if (first) mbar_wait(&bar_empty[ring_idx], ring_phase ^ 1u);
                    slot[(k & 3) * 32] = make_float4(a, b, c, d);
                    if (LW && k == 0) slot[kWsHL * 32] = make_float4(hdr0, hdr1, 0.f, 0.f);
                    if (last) { mbar_arrive(&bar_full[ring_idx]); ring_next(); }
                };
                GasLoads<LW, NG> GA, GB;
                if (LW) {
                    // bottom -> top; layer k leaves once pfrac of layer k + 1 is known (its top-level source is the
                    // geometric mean across the interface, compute_optical_props.jl:187-195)
                    const FT* pbk = W.plk + bl * (nlev + 1);
                    FT tau_p = 0.f, ssa_p = 0.f, g_p = 0.f, pf_p = 0.f;
                    auto lw_step = [&](int k, FT tau, FT ssa, FT g, FT pf) {
                        if (k == 0) {
                            hdr0 = pbk[0] * pf;          // source at level 0
                            hdr1 = pbk[nlev] * pf;       // surface Planck (compute_optical_props.jl:184-186)
                        } else {
                            const FT bk = pbk[k];
                            emit(k - 1, tau_p, ssa_p, g_p, hsqrt((bk * pf_p) * (bk * pf)));
                        }
                        tau_p = tau; ssa_p = ssa; g_p = g; pf_p = pf;
                    };
                    for (int part = 0; part * 32 < nlay; ++part) {
                        build_records(part);
                        const int lo = part * 32, hi = lo + 32 < nlay ? lo + 32 : nlay;
                        issue(lo, GA);
                        for (int k = lo; k < hi; k += 2) {
                            FT tau, ssa, g, pf;
                            if (k + 1 < hi) issue(k + 1, GB);
                            optics(k, GA, tau, ssa, g, pf);
                            lw_step(k, tau, ssa, g, pf);
                            if (k + 1 < hi) {
                                if (k + 2 < hi) issue(k + 2, GA);
                                optics(k + 1, GB, tau, ssa, g, pf);
                                lw_step(k + 1, tau, ssa, g, pf);
                            }
                        }
                    }
                    emit(nlay - 1, tau_p, ssa_p, g_p, pbk[nlay] * pf_p);   // top layer: its own increment (:193-195)
                } else {
                    for (int part = (nlay - 1) >> 5; part >= 0; --part) {   // top -> bottom
                        build_records(part);
                        const int lo = part * 32, hi = lo + 31 < nlay - 1 ? lo + 31 : nlay - 1;
                        issue(hi, GA);
                        for (int k = hi; k >= lo; k -= 2) {
                            FT tau, ssa, g, pf;
                            if (k - 1 >= lo) issue(k - 1, GB);
                            optics(k, GA, tau, ssa, g, pf);
                            emit(k, tau, ssa, g, 0.f);
                            if (k - 1 >= lo) {
                                if (k - 2 >= lo) issue(k - 2, GA);
                                optics(k - 1, GB, tau, ssa, g, pf);
                                emit(k - 1, tau, ssa, g, 0.f);
                            }
                        }
                    }
                    if (aod_here) {
                        aod_e = warp_sum(aod_e); aod_s = warp_sum(aod_s);
                        if (lane == 0) { P.io.aod_ext[col] = aod_e; P.io.aod_sca[col] = aod_s; }
                    }
                }
            }
            if (lane == 0 && P.io.cld_cover != nullptr && HAS_CLD) P.io.cld_cover[col] = __fdiv_rn(FT(n_cloudy), FT(NGPT));
            __syncwarp();
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_smem, 512u);
}

}  // namespace rb

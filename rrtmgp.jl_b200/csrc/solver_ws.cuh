// Warp-specialised two-stream kernels (round 2): the fused column solve of solver_fast.cuh split into two ROLES that
// run as a pipeline inside one persistent CTA of 16 warps (4 per scheduler instead of 3):
//
//   gas warps (8..15)   everything that ends in (tau, ssa, g[, Planck source]) of a (layer, g-point) cell: phase 0 /
//                       phase 1 / McICA of solver.cuh, the k-distribution corner gathers and the trilinear
//                       interpolation (gas_optics.jl:176-320, optics_utils.jl:85-181) and the cloud / aerosol
//                       increment.  Layers are independent here, so the gathers of layer k+1 are in flight while
//                       layer k is interpolated.
//   RT warps (0..7)     the longwave level sources (compute_optical_props.jl:157-195), the two-stream coefficients
//                       and the adding recurrences (longwave_2stream.jl:149-334,
//                       shortwave_2stream.jl:189-392), the g-point reductions and the (nlev, ncol) epilogue.  Their
//                       level store is tensor memory only: 2 warps per TMEM lane quadrant, 256 columns each, so all
//                       three values of all 64 levels fit and the shared-memory albedo spill of solver_fast.cuh is gone.
//
// Gas warp 8 + i feeds RT warp i through a shared-memory ring: stage = the (up to) four layers 4s..4s+3 of the block
// in hand, one float4 per (layer, lane) + two header rows (longwave: Planck function of the band at the stage's
// levels), `kWsStages` stages deep, full / empty mbarriers (all 32 lanes arrive).  The column index travels through a two-deep mailbox with its own mbarrier pair; gas warps own the
// atomic column queue.  Registers are re-balanced with setmaxnreg (RT 104, gas 152 = the whole register file).
//
// Why not TMA for the corner gathers (north star): measured on B200
// (profiles/r2a_tma_gather_micro.txt) one SM retires one cp.async.bulk / cp.async.bulk.tensor request per ~13 clocks
// whatever its size up to 512 B, so the 8-12 requests a (layer, 32 g-points) step needs take 105-135 clocks against
// 69 clocks for the same 4 KB through per-lane LDG -- and the whole fused step took 113.  TMA stays where it fits: the
// one-off bulk staging of the small tables (solver_fast.cuh).
//
// STATUS (measured, profiles/r2b_*, r2c_*): parity-green on the whole GPU suite, but SLOWER than the single-role
// kernels of solver_fast.cuh on B200 -- 27.3 + 22.8 ms against 20.2 + 18.6 ms per 1e5 columns.  The gas role is the
// bottleneck (141 instructions per (layer, 32 g-points) step against 113 for the RT role, at 8.8 clocks per
// instruction: its gathers have one layer of cover and two warps per scheduler to hide behind), so the RT warps sleep
// on the ring 60 % of the time.  Selected with RRTMGP_B200_KERNEL=ws; the default stays solver_fast.cuh.
#pragma once
#include "solver_fast.cuh"

namespace rb {

constexpr int kWsPairs = 8;          // (gas warp, RT warp) pairs per CTA
constexpr int kWsHL = 4;             // layers per hand-off stage
constexpr int kWsStages = 3;         // ring depth in stages
constexpr int kWsMaxLay = 64;
constexpr int kWsAccStride = kWsMaxLay + 4;
constexpr int kWsStageF4 = (kWsHL + 2) * 32;   // float4 per stage: four layer rows + two header rows (longwave Planck values)
constexpr int kWsRegsRT = 104, kWsRegsGas = 152;

struct WsSmem {
    int off_ring, off_stage, off_acc, off_bacc;   // byte offsets from the pair's base (after the gas warp's Warp<> layout)
    int off_blob, off_vmr, staged_bytes;          // CTA-shared tail
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the ring's barriers are addressed by their 32-bit shared-memory address, computed once (the generic -> shared
// conversion costs an S2R + address arithmetic each time)
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {   // with a suspend-time hint: a starved warp sleeps
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
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
    __shared__ const void* small_tables[TB_COUNT];   // where each small table lives (solver.cuh)
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
    float* svmr = P.vmr_kind == 0 ? reinterpret_cast<float*>(smem_raw + F.off_vmr) + 1 : nullptr;
    if (threadIdx.x == 0) {
        mbar_init(&blob_bar, 1);
        mbar_expect_tx(&blob_bar, (uint32_t)F.staged_bytes);
        if (F.staged_bytes > 0) tma_bulk_g2s(sblob, P.lut.blob, (uint32_t)F.staged_bytes, &blob_bar);
    }
    if (threadIdx.x < kWsPairs * (2 * kWsStages + 4)) {
        const int b = threadIdx.x % (2 * kWsStages + 4);
        mbar_init(&bars[threadIdx.x / (2 * kWsStages + 4)][b], b < 2 * kWsStages ? 32u : 1u);
    }
    if (svmr != nullptr)   // svmr[-1] = 1 (dry air), svmr[ig - 1] = global-mean vmr of gas ig
        for (int i = threadIdx.x; i <= P.ngas; i += blockDim.x) svmr[i - 1] = i == 0 ? 1.f : __ldg(P.io.vmr + i - 1);
    fill_small_table_pointers(small_tables, P, sblob, F.staged_bytes, (int)threadIdx.x);
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(&blob_bar, 0);

    unsigned char* pbase = smem_raw + (size_t)pair * P.warp_bytes;
    float4* ring_lane = reinterpret_cast<float4*>(pbase + F.off_ring) + lane;    // [kWsStages][kWsHL + 2][32], this lane's column
    const uint32_t bar0 = smem_u32(&bars[pair][0]);                   // full[s] at bar0 + 8 s, empty[s] at bar0 + 8 (kWsStages + s)
    uint64_t* col_full = &bars[pair][2 * kWsStages];
    uint64_t* col_empty = &bars[pair][2 * kWsStages + 2];
    const GasLut<FT>& L = P.lut;
    const int nlay = P.nlay, nlev = nlay + 1;
    // a small table: its copy inside the staged prefix of the block, or the global array
    auto tb = [&](const int* g) -> const int* {
        const long long off = reinterpret_cast<const unsigned char*>(g) - L.blob;
        return (off >= 0 && off < F.staged_bytes) ? reinterpret_cast<const int*>(sblob + off) : g;
    };
    // The ring stage in hand: its rows, its `full` barrier and the phase parity of the current pass.  Both roles step
    // through the stages in the same order (stage = layers 4s..4s+3 of a block), so they keep these in lock step.
    float4* sp = ring_lane;
    uint32_t bfull = bar0, ring_phase = 0;
    int ring_idx = 0;
    auto ring_next = [&]() {
        if (++ring_idx == kWsStages) { ring_idx = 0; ring_phase ^= 1u; sp = ring_lane; bfull = bar0; }
        else { sp += kWsStageF4; bfull += 8u; }
    };

    if (is_rt) {
        // =====================================================================================================
        // RT role: level sources (longwave), coefficients + adding + reductions + epilogue
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
                    // longwave_2stream.jl:243-334, adding from the bottom.  Row of layer k: (tau, ssa, g, Planck fraction);
                    // header rows of a stage: Planck function of the band at levels 4s+1..4s+4, and (stage 0) at level 0 and
                    // of the surface.  The source at the top level of layer k is the geometric mean across the interface
                    // (compute_optical_props.jl:187-195), so layer k is closed once row k + 1 has arrived.
                    const FT emis = __ldg(P.io.sfc_emis + (size_t)col * L.n_bnd + ibnd);
                    const FT inc = P.io.inc_flux_lw ? __ldg(P.io.inc_flux_lw + (size_t)gpt * P.ncol_total + col) : 0.f;
                    mbar_wait_a(bfull, ring_phase);
                    float4 v = sp[0];
                    FT lev_bot, albedo = 1.f - emis, src;
                    {
                        const float4 h = sp[(kWsHL + 1) * 32];
                        lev_bot = h.x * v.w;                               // source at level 0
                        src = Num<FT>::pi() * emis * (h.y * v.w);          // surface (compute_optical_props.jl:184-186)
                    }
                    const FT* bkp = reinterpret_cast<const FT*>(sp + kWsHL * 32);   // Planck at level k + 1, this lane's band
                    const float4* vp = sp + 32;                                     // row k + 1
                    for (int t0 = 0; t0 < nlay; t0 += 16) {                  // tiles of <= 16 layers
                        const int tend = t0 + 16 < nlay ? t0 + 16 : nlay;
#pragma unroll 1
                        for (int k = t0; k < tend; ++k) {
                            const bool stage_end = (k & 3) == 3 || k + 1 == nlay;
                            const FT bk = *bkp;
                            float4 vn = v;
                            uint32_t done_bar = 0;
                            if (stage_end) {                                 // warp-uniform
                                done_bar = bfull + 8u * kWsStages;           // `empty` of the stage in hand
                                if (k + 1 < nlay) {
                                    ring_next();
                                    mbar_wait_a(bfull, ring_phase);
                                    vn = sp[0];
                                    bkp = reinterpret_cast<const FT*>(sp + kWsHL * 32);
                                    vp = sp + 32;
                                }
                            } else {
                                vn = *vp;
                                vp += 32; ++bkp;
                            }
                            const FT inc_k = bk * v.w;
                            const FT lev_top = k + 1 < nlay ? hsqrt(inc_k * (bk * vn.w)) : inc_k;
                            const LwCoef C = lw_2stream_coeffs_nosrc(v.x, v.y, v.z);
                            const FT denom = rcp_approx(1.f - C.Rdif * albedo);
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
                            v = vn;
                            if (stage_end) mbar_arrive_a(done_bar);
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
                    ring_next();                                             // past the block's last stage
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
                            mbar_wait_a(bfull, ring_phase);
                            const float4* vp = sp + (shi & 3) * 32;
                            FT* stp = stage + ((shi - kc) * 2) * kStageStride + lane;
#pragma unroll 1
                            for (int k = shi; k >= slo; --k) {
                                const float4 v = *vp;
                                vp -= 32;
                                FT Rdir, Tdir, Rdif, Tdif;
                                sw_2stream_coeffs(v.x, v.y, v.z, mu0, inv_mu0, Rdir, Tdir, Rdif, Tdif);
                                const FT su = Rdir * dir, sd = Tdir * dir;       // dir = direct flux at level k+1
                                const FT denom = rcp_approx(1.f - Rdif * beta);
                                // F_up(k+1) = A'_k F_up(k) + B'_k ; F_dn_dif(k+1) = beta_{k+1} F_up(k+1) + d_{k+1}
                                tmem_st2(tA + 2 * k, Tdif * denom, (Rdif * d + su) * denom);
                                tmem_st1(tAl + k, beta);
                                stp[0] = d;                                       // d_{k+1}
                                d = sd + Tdif * denom * (d + beta * su);
                                beta = Rdif + Tdif * Tdif * beta * denom;
                                tau_cum += v.x;
                                float ex;
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(tau_cum * neg_inv_mu0_l2e));
                                dir = dir_top * ex;                               // direct flux at level k
                                stp[kStageStride] = dir;
                                stp -= 2 * kStageStride;
                            }
                            mbar_arrive_a(bfull + 8u * kWsStages);
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
        // gas role: phase 0 / phase 1 / McICA, corner gathers, interpolation, increments
        // =====================================================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsRegsGas));
        constexpr int RW = 20 + 4 * NG;                                   // words per band record (plan_smem_ws)
        constexpr int RR = (((2 * RW) >> 2) & 1) ? 2 * RW : 2 * RW + 4;   // words per record row
        // per-lane table bases as opaque 64-bit values: one IMAD.WIDE per gather base instead of a uniform base +
        // lane offset re-added (and sign-extended) for every address
        auto opaque = [](const void* p) { unsigned long long v = (unsigned long long)p; asm("" : "+l"(v)); return v; };
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
            W.tptr = small_tables;
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
                const unsigned long long major_lane = opaque((LW ? L.kmaj_pf : L.kmajor) + (LW ? 2 : 1) * gpt);
                const unsigned long long minor_lane = opaque(reinterpret_cast<const float4*>(L.kminor4[0]) + gpt);

                // ---- issue the table gathers of the cell whose band-record row is `rk`; `cb` = cloudy (McICA) ----
                auto issue = [&](const FT* rk, bool cb, GasLoads<LW, NG>& G) {
                    G.s = *reinterpret_cast<const float4*>(rk + 8);
                    G.x = *reinterpret_cast<const float4*>(rk + 12 + 4 * NG + (cb ? 4 : 0));
                    const int ia = __float_as_int(G.s.z), ib = __float_as_int(G.s.w);   // (jp-1, jt, je1), (jp-1, jt+1, je2)
                    const int ma = __float_as_int(G.x.w), mb = ma + (ib - ia);          // (jt, je1), (jt+1, je2): MT == KT
                    if (LW) {
                        const float2* pa = reinterpret_cast<const float2*>(major_lane) + ia;
                        const float2* pb = reinterpret_cast<const float2*>(major_lane) + ib;
                        G.c2[0] = __ldg(pa); G.c2[1] = __ldg(pa + KE); G.c2[2] = __ldg(pa + KP); G.c2[3] = __ldg(pa + KP + KE);
                        G.c2[4] = __ldg(pb); G.c2[5] = __ldg(pb + KE); G.c2[6] = __ldg(pb + KP); G.c2[7] = __ldg(pb + KP + KE);
                    } else {
                        const FT* pa = reinterpret_cast<const FT*>(major_lane) + ia;
                        const FT* pb = reinterpret_cast<const FT*>(major_lane) + ib;
                        G.c1[0] = __ldg(pa); G.c1[1] = __ldg(pa + KE); G.c1[2] = __ldg(pa + KP); G.c1[3] = __ldg(pa + KP + KE);
                        G.c1[4] = __ldg(pb); G.c1[5] = __ldg(pb + KE); G.c1[6] = __ldg(pb + KP); G.c1[7] = __ldg(pb + KP + KE);
                    }
                    const float4* qa = reinterpret_cast<const float4*>(minor_lane) + ma;
                    const float4* qb = reinterpret_cast<const float4*>(minor_lane) + mb;
#pragma unroll
                    for (int gi = 0; gi < NG; ++gi) {
                        G.m[4 * gi + 0] = __ldg(qa + gi * MS); G.m[4 * gi + 1] = __ldg(qa + gi * MS + ME);
                        G.m[4 * gi + 2] = __ldg(qb + gi * MS); G.m[4 * gi + 3] = __ldg(qb + gi * MS + ME);
                    }
                };
                // ---- gas + cloud + aerosol optics of the gathered cell (gas_optics.jl:176-320, optics_utils.jl:85-202);
                //      returns the hand-off row: LW (tau, ssa, g, Planck fraction), SW (tau, ssa, g, 0) ----
                auto optics = [&](const FT* rk, const GasLoads<LW, NG>& G) -> float4 {
                    const float4 v0 = *reinterpret_cast<const float4*>(rk), v1 = *reinterpret_cast<const float4*>(rk + 4);
                    FT tau, ssa, g, pfrac;
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
                        const float4 sc = *reinterpret_cast<const float4*>(rk + 12 + 4 * gi);
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
                        g = G.x.z * rcp_approx(rmax(FLT_EPSILON, w));
                        ssa = w * rcp_approx(rmax(FLT_EPSILON, tn));
                        tau = tn;
                    }
                    return make_float4(tau, ssa, g, pfrac);
                };
                const uint32_t bempty_off = 8u * kWsStages;
                GasLoads<LW, NG> GA, GB;
                if (LW) {
                    // bottom -> top, stage by stage (layers 4s..4s+3); the gathers of the next layer are in flight while
                    // the current one is interpolated
                    const FT* pbk = W.plk + bl * (nlev + 1);
                    for (int part = 0; part * 32 < nlay; ++part) {
                        build_records(part);
                        const int lo = part * 32, hi = lo + 32 < nlay ? lo + 32 : nlay;
                        const FT* rk = rec_lane;
                        unsigned mw = HAS_CLD ? (part == 0 ? W.mask[0] : W.mask[1]) : 0u;   // bit j = layer lo + j
                        int k = lo;
                        if (k + 4 <= hi) issue(rk, mw & 1u, GA);
                        for (; k + 4 <= hi; k += 4) {
                            mbar_wait_a(bfull + bempty_off, ring_phase ^ 1u);
                            issue(rk + RR, (mw >> 1) & 1u, GB);
                            sp[0] = optics(rk, GA);
                            issue(rk + 2 * RR, (mw >> 2) & 1u, GA);
                            sp[32] = optics(rk + RR, GB);
                            issue(rk + 3 * RR, (mw >> 3) & 1u, GB);
                            sp[64] = optics(rk + 2 * RR, GA);
                            if (k + 8 <= hi) issue(rk + 4 * RR, (mw >> 4) & 1u, GA);
                            sp[96] = optics(rk + 3 * RR, GB);
                            sp[kWsHL * 32] = make_float4(pbk[k + 1], pbk[k + 2], pbk[k + 3], pbk[k + 4]);
                            if (k == 0) sp[(kWsHL + 1) * 32] = make_float4(pbk[0], pbk[nlev], 0.f, 0.f);
                            mbar_arrive_a(bfull);
                            ring_next();
                            rk += 4 * RR; mw >>= 4;
                        }
                        if (k < hi) {   // the column's top stage when nlay is not a multiple of 4
                            mbar_wait_a(bfull + bempty_off, ring_phase ^ 1u);
                            const int k0 = k;
                            for (; k < hi; ++k) {
                                issue(rk, mw & 1u, GA);
                                sp[(k & 3) * 32] = optics(rk, GA);
                                rk += RR; mw >>= 1;
                            }
                            const int n1 = nlev;   // pbk has nlev + 1 entries
                            sp[kWsHL * 32] = make_float4(pbk[k0 + 1], pbk[k0 + 2 < n1 ? k0 + 2 : n1], pbk[k0 + 3 < n1 ? k0 + 3 : n1], pbk[k0 + 4 < n1 ? k0 + 4 : n1]);
                            if (k0 == 0) sp[(kWsHL + 1) * 32] = make_float4(pbk[0], pbk[nlev], 0.f, 0.f);
                            mbar_arrive_a(bfull);
                            ring_next();
                        }
                    }
                } else {
                    for (int part = (nlay - 1) >> 5; part >= 0; --part) {   // top -> bottom, stage by stage (layers 4s+3..4s)
                        build_records(part);
                        const int lo = part * 32;
                        int k = lo + 31 < nlay - 1 ? lo + 31 : nlay - 1;
                        const FT* rk = rec_lane + (k & 31) * RR;
                        unsigned mw = HAS_CLD ? (part == 0 ? W.mask[0] : W.mask[1]) << (31 - (k & 31)) : 0u;   // bit 31 = layer k
                        if ((k & 3) != 3) {   // the column's top stage when nlay is not a multiple of 4
                            mbar_wait_a(bfull + bempty_off, ring_phase ^ 1u);
                            for (;; --k) {
                                issue(rk, (mw >> 31) & 1u, GA);
                                sp[(k & 3) * 32] = optics(rk, GA);
                                rk -= RR; mw <<= 1;
                                if ((k & 3) == 0) break;
                            }
                            --k;
                            mbar_arrive_a(bfull);
                            ring_next();
                        }
                        if (k >= lo) issue(rk, (mw >> 31) & 1u, GA);
                        for (; k >= lo; k -= 4) {
                            mbar_wait_a(bfull + bempty_off, ring_phase ^ 1u);
                            issue(rk - RR, (mw >> 30) & 1u, GB);
                            sp[96] = optics(rk, GA);
                            issue(rk - 2 * RR, (mw >> 29) & 1u, GA);
                            sp[64] = optics(rk - RR, GB);
                            issue(rk - 3 * RR, (mw >> 28) & 1u, GB);
                            sp[32] = optics(rk - 2 * RR, GA);
                            if (k - 4 >= lo) issue(rk - 4 * RR, (mw >> 27) & 1u, GA);
                            sp[0] = optics(rk - 3 * RR, GB);
                            mbar_arrive_a(bfull);
                            ring_next();
                            rk -= 4 * RR; mw <<= 4;
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

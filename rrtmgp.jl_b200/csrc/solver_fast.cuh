// Blackwell-native fast path of the two-stream kernels and of the no-scattering longwave kernel: Float32, nlay <= 95,
// real-table shape (n_eta = 9, n_T = 14, 16-g-point bands, n_gpt and the number of minor-slot groups template constants).
//
//  * PERSISTENT: one CTA of 12 warps per SM (8 for columns taller than 64 layers, FastGeom); warp = column,
//    lane = g-point (as in solver.cuh).  Columns come
//    from an atomic work queue, the next one is known one column ahead and its inputs are prefetched into L2.
//  * TENSOR MEMORY as the level store.  The adding method needs, for every level, three values
//    per (column, g-point) from the first sweep when the second sweep passes the same level:
//    768 B per lane, 24.6 KB per warp.  In shared memory that caps residency at 4-6 warps per
//    SM and the kernel is latency bound (profiles/r1a_*).  B200's 256 KB of TMEM per SM are
//    otherwise idle here (no MMA on this path) and tcgen05.ld/st give every thread a private
//    column array in its own TMEM lane -- exactly the access pattern of this store.  The CTA
//    allocates all 512 columns; the three warps that share a lane quadrant get 170 columns
//    each: (A, B) of 64 levels + the albedo of the lowest 41 levels; the other 23 albedos sit
//    in shared memory.
//  * SMALL TABLES STAGED BY TMA: key species, reference vmr, minor-absorber lists, Planck table, cloud and
//    aerosol tables of the sweep are one block of the arena (GasLut::blob), copied to shared memory once per
//    CTA with cp.async.bulk + mbarrier; phase 1 never waits on global memory for them.
//  * BAND RECORDS: everything a (layer, band) contributes to its 16 g-points -- 8 trilinear corner weights,
//    column amounts, table offsets, slot scalings, cloud/aerosol increment products -- is computed once per
//    (layer, band) by phase 1 (lane = layer) and read by the cells with five 128-bit loads.
//  * COMPILE-TIME TABLE STRIDES and PACKED TABLES: a cell gathers its 8 {kmajor, Planck fraction}
//    corners with 64-bit loads and its minor absorbers (+ Rayleigh) four slots at a time with
//    128-bit loads: 12 loads instead of 28-32 scalar loads.
//  * LOADS FIRST, ONE BASIC BLOCK PER ITERATION: an iteration issues the gathers of layer k, then does the
//    source-independent two-stream coefficients of layer k-1, then interpolates layer k and closes layer k-1;
//    record rebuilds and g-point reductions sit between tiles of <= 16 iterations, nothing branches inside.
//  * G-POINT REDUCTION through a shared staging tile read transposed (lane = level) with 128-bit loads.
//  * SECOND SWEEPS read the level store 8 levels at a time (one tcgen05.ld.x16 for the (A, B) pairs, one .x8 or eight
//    LDS for the albedos) instead of a TMEM round trip per level.
//  * BAND-RECORD ROWS are 4 * odd words (solver_launch.cuh): phase 1 writes them with lane = layer, 128-bit stores.
//  * NO-SCATTERING LW (MODE_LW_NOSCAT, 1-4 Gauss angles): marched from the top, the down sweep of every angle in the
//    pass that does the gas optics; (tau, B_lay pfrac, top-level source) per layer in the level store for the up sweep.
//  * SW adding marched from the TOP (reflectance/source of everything ABOVE a level),
//    algebraically identical to shortwave_2stream.jl:300-392, so the direct beam, the layer
//    coefficients and the first recurrence share one sweep; see DESIGN.md.
#pragma once
#include "solver.cuh"

#include <type_traits>

// ABLATION BUILDS (tools/ablation.sh, DESIGN.md section 7 "where the time goes"): -DRB_WHATIF=n removes one part of the kernels
// so that its cost in elapsed time can be measured (the results of such a build are wrong by construction; the
// shipped library is built with RB_WHATIF = 0 and contains none of this).  1: band records (phase 1) only for the
// first g-point block of a column; 2: no McICA sampling; 3: no second sweeps; 8: 1 + 2 + 3; 4: no table gathers;
// 9: table gathers always from the same few rows (all L1 hits); 5: trivial two-stream coefficients; 6: no level-store writes.
#ifndef RB_WHATIF
#define RB_WHATIF 0
#endif

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
// 8 levels at once: 16 consecutive columns ((A, B) pairs) / 8 consecutive columns (albedos) of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Geometry of the persistent CTA: 12 warps (3 per TMEM lane quadrant, 170 columns each) for nlay <= 64, or 8 warps
// (2 per quadrant, 256 columns each) for nlay <= 96.  Per warp: (A, B) of every level, then as many albedos as fit
// (+ one dummy column); the other albedos go to shared memory.
template <int WARPS> struct FastGeom {
    static_assert(WARPS == 12 || WARPS == 8, "3 or 2 warps per TMEM lane quadrant");
    static constexpr int warps = WARPS;
    static constexpr int max_lay = WARPS == 12 ? 64 : 96;
    static constexpr int cols_per_warp = 512 / (WARPS / 4);
    static constexpr int alpha_tmem_levels = cols_per_warp - 2 * max_lay - 1;   // 41 / 63
    static constexpr int acc_stride = max_lay + 4;    // per-quantity stride of the shared broadband accumulators
    static constexpr int n_own = max_lay / 32;        // layers per lane in phases 0 / 1; record parts per column
};
constexpr int kStageStride = 36;      // staging-tile row stride: 16-byte aligned rows, conflict-free 128-bit row reads

struct FastSmem {
    int off_alpha, off_stage, off_acc;   // byte offsets from the warp's base, extending SolveParams' layout
    int off_blob, off_vmr;               // CTA-shared: staged small-table block, global-mean vmr array (from the smem base)
    int staged_bytes;                    // staged prefix of GasLut::blob
    int off_bacc;                        // per warp: per-band accumulators [2 bands][up, dn][kAccStride] (spectral fluxes), or -1
    int active_warps;                    // warps per CTA that take work items (fewer than all when there are few columns)
};

// ---- TMA bulk copy global -> shared with mbarrier completion (the small-table block, once per CTA) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Raw table corners and band-record words of one (layer, g-point) cell: the loads are issued together
// (fast_gather) and consumed later (fast_finish), with independent arithmetic of the previous layer in between.
template <bool LW, int NG> struct FastCell {
    float2 c2[LW ? 8 : 1];   // LW: {kmajor, planck_fraction} corners
    float c1[LW ? 1 : 8];    // SW: kmajor corners
    float4 m[4 * NG];        // per group: four minor-absorber slots (SW slot 0 = Rayleigh) at the four (T, eta) corners
    float4 v0, v1;           // band record: trilinear corner weights
    float4 s;                // s1, s2 (column amounts of the two T nodes), major-table offsets
    float4 sc[NG];           // slot scalings
    float4 x;                // increment products (aerosol only or cloud + aerosol), minor-table offset
};

// NMU (no-scattering LW only): 1 = one Gauss angle, 4 = up to four, P.n_mu of them active
template <int MODE, int NGPT, int NG, bool HAS_CLD, bool HAS_AER, bool SPECTRAL, int WARPS, int NMU = 1>
__global__ void __launch_bounds__(WARPS * 32, 1) solve_kernel_fast(const SolveParams<float> P, const FastSmem F) {
    using FT = float;
    using Geom = FastGeom<WARPS>;
    constexpr int kFastWarps = Geom::warps, kFastColsPerWarp = Geom::cols_per_warp, kAlphaTmemLevels = Geom::alpha_tmem_levels;
    constexpr int kAccStride = Geom::acc_stride, NOWN = Geom::n_own, kMaxLay = Geom::max_lay;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(8) uint64_t blob_bar;
    __shared__ const void* small_tables[TB_COUNT];   // where each small table lives (solver.cuh)
    constexpr bool LW = MODE == MODE_LW_2STREAM;
    constexpr bool NOSCAT = MODE == MODE_LW_NOSCAT;
    constexpr bool LWG = LW || NOSCAT;      // longwave gas optics: {kmajor, Planck fraction} pairs, no Rayleigh, always "day"
    constexpr bool INCR = HAS_CLD || HAS_AER;
    constexpr int NETA = 9, NT = 14;
    constexpr int KE = NGPT, KT = NETA * KE, KP = NT * KT;           // major-table strides (LW: in float2): eta, T, p
    constexpr int ME = NGPT, MT = NETA * NGPT, MS = NT * MT;         // minor-table strides in float4: eta, T, group
    constexpr int UP = 0, DN = 1, DIR = 2;

    const int lane = threadIdx.x & 31;
    // warp index through a shuffle: provably warp-uniform, so everything derived from it (column, shared
    // memory bases, table descriptors) lives in uniform registers instead of being re-broadcast with R2UR
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);

    // Stage this sweep's small tables (GasLut::blob: key species, reference vmr, minor-absorber lists, Planck
    // table, cloud and aerosol tables) into shared memory with one TMA bulk copy: the persistent CTA reads them
    // ~10^5 times per column and they would otherwise fight the k-distribution gathers for L1.
    unsigned char* sblob = smem_raw + F.off_blob;
    float* svmr = P.vmr_kind == 0 ? reinterpret_cast<float*>(smem_raw + F.off_vmr) + 1 : nullptr;
    if (threadIdx.x == 0) {
        mbar_init(&blob_bar, 1);
        mbar_expect_tx(&blob_bar, (uint32_t)F.staged_bytes);
        if (F.staged_bytes > 0) tma_bulk_g2s(sblob, P.lut.blob, (uint32_t)F.staged_bytes, &blob_bar);
    }
    if (svmr != nullptr)   // svmr[-1] = 1 (dry air), svmr[ig - 1] = global-mean vmr of gas ig
        for (int i = threadIdx.x; i <= P.ngas; i += blockDim.x) svmr[i - 1] = i == 0 ? 1.f : __ldg(P.io.vmr + i - 1);
    fill_small_table_pointers(small_tables, P, sblob, F.staged_bytes, (int)threadIdx.x);
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(&blob_bar, 0);
    // lane field (bits 31:16) = 32 * (warp % 4); column = 170 * (warp / 4)
    const uint32_t tA = tmem_base_smem + ((uint32_t)(warp & 3) << 21) + (uint32_t)((warp >> 2) * kFastColsPerWarp);
    const uint32_t tAl = tA + 2u * kMaxLay;

    unsigned char* wbase = smem_raw + (size_t)warp * P.warp_bytes;
    FT* alpha_hi = reinterpret_cast<FT*>(wbase + F.off_alpha);   // [nlay - 41 + 1][32]
    FT* stage = reinterpret_cast<FT*>(wbase + F.off_stage);      // [16][kStageStride]
    FT* accs = reinterpret_cast<FT*>(wbase + F.off_acc);         // [3][kAccStride]
    // per-band (spectral) fluxes, Fluxes.jl:170-215: lanes 0-15 / 16-31 of a block are its two 16-g-point bands, so
    // the half-row sums of the staging tile are the band sums
    constexpr bool spectral = SPECTRAL;   // a template parameter: the broadband-only kernels carry none of this
    FT* bacc = reinterpret_cast<FT*>(wbase + (F.off_bacc >= 0 ? F.off_bacc : 0));   // [2][2][kAccStride]
    const int hb = lane >> 4;                                    // this lane's half = band within the block
    const GasLut<FT>& L = P.lut;
    const int nlay = P.nlay, nlev = nlay + 1;
    const FT* major = LWG ? L.kmaj_pf : L.kmajor;
    const float4* minor4 = reinterpret_cast<const float4*>(L.kminor4[0]);
    constexpr int RW = 20 + 4 * NG;                                   // words per band record (plan_smem_fast)
    constexpr int RR = (((2 * RW) >> 2) & 1) ? 2 * RW : 2 * RW + 4;   // words per record row: 4 * odd

    // Columns are handed out by an atomic counter, in order, one at a time: night columns (SW) and cloud-free
    // columns cost a fraction of the others, and a static assignment leaves the time of a launch to the unluckiest
    // of the 1776 warps.  Every warp holds its next assignment one column ahead, so that column's inputs (read
    // exactly once, cold in DRAM) are pulled into L2 while the current one is computed.
    auto next_column = [&]() -> long long {
        unsigned int v = 0;
        if (lane == 0) v = atomicAdd(P.work_counter, 1u);
        return (long long)__shfl_sync(0xffffffffu, v, 0);
    };
    // SMALL SHARDS (P.split = 2 or 4, chosen by the host when a warp would get fewer than four columns, api.cu): a work item is
    // (column, contiguous share of its g-point blocks), so the last item of a warp costs a half / a quarter of a column.
    // Every item leaves its broadband partial sums in global scratch; the LAST arriver of a column (atomic counter) adds
    // the shares in the fixed order 0, 1, .. and writes the column -- results do not depend on who arrives last.
    const int split_log2 = P.split == 4 ? 2 : (P.split == 2 ? 1 : 0);
    const long long nitems = (long long)P.ncol << split_log2;
    constexpr int kScr = 3 * kAccStride + 4;      // floats per (column, share) of the scratch: accumulators + cloudy count
    // few work items: the launcher spreads them over the SMs (one CTA each) and lets only `active_warps` warps per CTA
    // take any, so e.g. 128 columns run one warp per SM instead of twelve warps on eleven SMs
    long long col_next = warp < F.active_warps ? next_column() : nitems;
    while (col_next < nitems) {
        const long long item = col_next;
        const long long col = item >> split_log2;
        const int share = (int)(item & ((1 << split_log2) - 1));
        col_next = next_column();
        Warp<FT, MODE, NOWN, true> W(P, wbase, lane, col, sblob, F.staged_bytes, svmr);
        W.tptr = small_tables;
        {
            const long long nc = col_next >> split_log2;
            if (nc < P.ncol && nc != col) {
                // one rolled loop over (row pointer, bytes) pairs: inlined per array this was 13 KB of code run once per
                // column, which evicted the hot loops from the instruction cache
                const char* rows[12];
                int bytes[12];
                int nrow = 0;
                auto add_row = [&](const FT* base, int n) {
                    if (base == nullptr) return;
                    rows[nrow] = reinterpret_cast<const char*>(base + (size_t)nc * n);
                    bytes[nrow++] = n * (int)sizeof(FT);
                };
                add_row(P.io.layerdata, 4 * nlay);
                add_row(P.io.t_lev, nlev);
                if (P.vmr_kind == 0) { add_row(P.io.vmr_h2o, nlay); add_row(P.io.vmr_o3, nlay); }
                else add_row(P.io.vmr, nlay * P.ngas);
                if (HAS_CLD) {
                    add_row(P.io.cld_frac, nlay); add_row(P.io.cld_path_liq, nlay); add_row(P.io.cld_path_ice, nlay);
                    add_row(P.io.cld_r_eff_liq, nlay); add_row(P.io.cld_r_eff_ice, nlay);
                }
                if (HAS_AER) { add_row(P.io.aero_mass, 15 * nlay); add_row(P.io.aero_size, 15 * nlay); }
#pragma unroll 1
                for (int a = 0; a < nrow; ++a) {
                    const char* b = rows[a];
                    const int nb = bytes[a];
#pragma unroll 1
                    for (int o = lane * 128; o < nb + 127; o += 32 * 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(b + (o < nb ? o : nb - 1)));
                }
            }
        }
        W.phase0();
        for (int i = lane; i < 3 * kAccStride; i += 32) accs[i] = FT(0);
        if (spectral)
            for (int i = lane; i < 4 * kAccStride; i += 32) bacc[i] = FT(0);

        const uint64_t col_key = mcica_col_key(P.seed, (uint64_t)(P.col_offset + col));
        int cld_start = 0, cld_finish = 0;
        if (HAS_CLD) {
            const FT* cf = P.io.cld_frac + (size_t)col * nlay;
            unsigned lo = 0xffffffffu, hi = 0;
            bool frac = false;   // a fraction strictly between 0 and 1 (or above 1): the sample needs its draws
#pragma unroll
            for (int j = 0; j < NOWN; ++j) {
                const int k = lane + 32 * j;
                const FT c = k < nlay ? __ldg(cf + k) : FT(0);
                if (c > FT(0)) { lo = lo < (unsigned)(k + 1) ? lo : (unsigned)(k + 1); hi = hi > (unsigned)(k + 1) ? hi : (unsigned)(k + 1); }
                frac = frac || (c > FT(0) && c != FT(1));
                W.cf_words[j] = __ballot_sync(0xffffffffu, c > FT(0));
            }
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (hi > 0) { cld_start = (int)lo; cld_finish = (int)hi; }
            W.cf_binary = !__any_sync(0xffffffffu, frac);
        }
        const FT mu0 = LWG ? FT(1) : __ldg(P.io.cos_zenith + col);
        const bool day = LWG || mu0 > FT(0);
        const FT toa = LWG ? FT(0) : __ldg(P.io.toa_flux + col);
        int n_cloudy = 0;
        __syncwarp();

        // this item's blocks of 32 g-points (NGPT is a multiple of 32: every lane owns a g-point)
        const int blk0 = (share * (NGPT / 32)) >> split_log2, blk1 = ((share + 1) * (NGPT / 32)) >> split_log2;
        for (int g0 = 32 * blk0; g0 < 32 * blk1; g0 += 32) {
            W.set_block(g0);
            __syncwarp();
#if RB_WHATIF == 2 || RB_WHATIF == 8
            W.mask[0] = W.mask[NOWN - 1] = 0xffffffffu;
#else
            if (HAS_CLD) n_cloudy += W.mcica(col_key, cld_start, cld_finish);
#endif
            // band records are built half a column (32 layers) at a time, just before the sweep needs them
            FT aod_e = 0.f, aod_s = 0.f;
            const bool aod_here = !LWG && HAS_AER && P.io.aod_ext != nullptr && P.aero.iband_550nm >= W.b_first + 1 &&
                                  P.aero.iband_550nm <= W.b_first + W.nb;
            auto build_records = [&](int half) {
#if RB_WHATIF == 1 || RB_WHATIF == 8
                if (g0 != 32 * blk0) return;
#endif
                __syncwarp();
                FT e, sc;
                W.phase1(e, sc, half);
                aod_e += e; aod_s += sc;
            };
            // writes (or, at night, zeroes) the per-band fluxes of this block's two bands and clears the accumulators
            auto flush_bands = [&](bool zero) {
                __syncwarp();
                for (int b = 0; b < W.nb; ++b) {
                    const size_t ob = ((size_t)(W.b_first + b) * P.ncol_total + col) * nlev;
                    for (int lev = lane; lev < nlev; lev += 32) {
                        FT bu = zero ? 0.f : bacc[(b * 2 + UP) * kAccStride + lev];
                        FT bd = zero ? 0.f : bacc[(b * 2 + DN) * kAccStride + lev];
                        bacc[(b * 2 + UP) * kAccStride + lev] = 0.f; bacc[(b * 2 + DN) * kAccStride + lev] = 0.f;
                        if (P.io.metric_scaling != nullptr) { const FT sc = __ldg(P.io.metric_scaling + (size_t)col * nlev + lev); bu *= sc; bd *= sc; }
                        P.io.band_up[ob + lev] = bu; P.io.band_dn[ob + lev] = bd; P.io.band_net[ob + lev] = bu - bd;
                    }
                }
                __syncwarp();
            };
            if (!day) {   // night: AOD and masks only (shortwave_2stream.jl:66-102); the band records of the 550 nm
                          // block are still built, by the ONE call site of the shortwave sweep below (phase 1 is ~15 KB of
                          // code per inlined copy, and the hot code of a block should fit the 32 KB instruction cache)
                if (spectral) flush_bands(true);
                if (!aod_here) continue;
            }

            const int gpt = W.gpt, ibnd = W.ibnd, bl = W.bl;
            const FT* rec_lane = W.rec + bl * RW;          // this lane's band within a record row pair
            // per-lane table bases (offset by this lane's g-point) as opaque 64-bit values: every gather base is then one
            // IMAD.WIDE instead of a uniform base + lane offset re-added and sign-extended per address (profiles/r1r: 9
            // integer instructions for the two minor-table bases)
            auto opaque = [](const void* p) { unsigned long long v = (unsigned long long)p; asm("" : "+l"(v)); return v; };
            const unsigned long long major_lane = opaque(major + (LWG ? 2 : 1) * gpt);
            const unsigned long long minor_lane = opaque(minor4 + gpt);
            const unsigned mask0 = W.mask[0], mask1 = W.mask[1], mask2 = W.mask[NOWN - 1];   // (mask2 used when NOWN = 3)
            auto mask_word = [&](int k) -> unsigned { return k < 32 ? mask0 : ((NOWN > 2 && k >= 64) ? mask2 : mask1); };

            // ---- issue every load of the cell whose band-record row is `r` (`cb`: cloudy in this g-point's McICA
            //      sample): compile-time strides, 64/128-bit gathers ----
            auto gather_r = [&](const FT* r, bool cb, FastCell<LWG, NG>& G) {
                G.s = *reinterpret_cast<const float4*>(r + 8);
                G.x = *reinterpret_cast<const float4*>(r + 12 + 4 * NG + (cb ? 4 : 0));
                G.v0 = *reinterpret_cast<const float4*>(r);
                G.v1 = *reinterpret_cast<const float4*>(r + 4);
#pragma unroll
                for (int gi = 0; gi < NG; ++gi) G.sc[gi] = *reinterpret_cast<const float4*>(r + 12 + 4 * gi);
#if RB_WHATIF == 9
                const int ia = __float_as_int(G.s.z) & 1, ib = (__float_as_int(G.s.w) & 1) + KT;
                const int ma = __float_as_int(G.x.w) & 1, mb = ma + (ib - ia);
#else
                const int ia = __float_as_int(G.s.z), ib = __float_as_int(G.s.w);   // (jp-1, jt, je1), (jp-1, jt+1, je2)
                const int ma = __float_as_int(G.x.w), mb = ma + (ib - ia);          // (jt, je1), (jt+1, je2): MT == KT
#endif
#if RB_WHATIF == 4
                {
                    const float f = __int_as_float((ia & 0xff) | 0x3f000000), h = __int_as_float((ib & 0xff) | 0x3e000000);
#pragma unroll
                    for (int i = 0; i < (LWG ? 8 : 1); ++i) G.c2[i] = make_float2(f * 1e-3f, h);
#pragma unroll
                    for (int i = 0; i < (LWG ? 1 : 8); ++i) G.c1[i] = f * 1e-3f;
#pragma unroll
                    for (int i = 0; i < 4 * NG; ++i) G.m[i] = make_float4(f * 1e-4f, h * 1e-4f, f * 1e-5f, h * 1e-5f);
                    (void)ma; (void)mb;
                    return;
                }
#endif
                if (LWG) {   // {kmajor, planck_fraction} pairs
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
            auto gather = [&](int k, FastCell<LWG, NG>& G) {
                bool cb = false;
                if (HAS_CLD) cb = (mask_word(k) >> (k & 31)) & 1u;
                gather_r(rec_lane + (k & 31) * RR, cb, G);
            };
            // ---- gas + cloud + aerosol optics of the gathered cell (gas_optics.jl:176-320, optics_utils.jl:85-181) ----
            auto finish = [&](const FastCell<LWG, NG>& G, FT& tau, FT& ssa, FT& g, FT& pfrac) {
                const float4 v0 = G.v0, v1 = G.v1;
                if (LWG) {
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
                // minor absorbers (+ Rayleigh in SW slot 0): four slots per 128-bit load (optics_utils.jl:85-98);
                // the (T, eta) weights are the corner weights summed over the two pressure nodes
                const FT w11 = v0.x + v0.z, w21 = v0.y + v0.w, w12 = v1.x + v1.z, w22 = v1.y + v1.w;
                FT tau_ray = 0.f;
#pragma unroll
                for (int gi = 0; gi < NG; ++gi) {   // real tables can have more than four (three in SW) minors per band
                    const float4 m11 = G.m[4 * gi], m21 = G.m[4 * gi + 1], m12 = G.m[4 * gi + 2], m22 = G.m[4 * gi + 3];
                    const float4 sc = G.sc[gi];
                    const FT v0 = w11 * m11.x + w21 * m21.x + w12 * m12.x + w22 * m22.x;
                    const FT v1 = w11 * m11.y + w21 * m21.y + w12 * m12.y + w22 * m22.y;
                    const FT v2 = w11 * m11.z + w21 * m21.z + w12 * m12.z + w22 * m22.z;
                    const FT v3 = w11 * m11.w + w21 * m21.w + w12 * m12.w + w22 * m22.w;
                    if (!LWG && gi == 0) {
                        tau_ray = v0 * sc.x;
                        tau += v1 * sc.y + v2 * sc.z + v3 * sc.w;
                    } else {
                        tau += v0 * sc.x + v1 * sc.y + v2 * sc.z + v3 * sc.w;
                    }
                }
                if (LWG) {
                    tau = rmax(tau, 0.f);
                    ssa = 0.f; g = 0.f;
                } else {
                    tau = rmax(tau + tau_ray, 0.f);
                    // gas ssa = tau_ray / tau (0 if tau <= 0, gas_optics.jl:313-318).  With an increment to follow only the
                    // product tau ssa is used, and that is tau_ray itself: no division
                    if (INCR) tau_ray = tau > 0.f ? tau_ray : 0.f;
                    else ssa = tau > 0.f ? hdiv(tau_ray, tau) : 0.f;
                    g = 0.f;
                }
                // one fused, unconditional increment (optics_utils.jl:189-202 is additive in tau, tau ssa, tau ssa g;
                // the record holds zeros where neither cloud nor aerosol is present)
                if (INCR && NOSCAT) {
                    tau += G.x.x;   // absorption optical depth of cloud + aerosol (cloud_optics.jl:1-50, aerosol_optics.jl:18-61)
                } else if (INCR) {
                    const FT tn = tau + G.x.x;
                    const FT w = LW ? G.x.y : tau_ray + G.x.y;         // LW gas: ssa = 0; SW gas: tau ssa = tau_ray
                    const FT h = G.x.z;                                  // gas: g = 0 in both
                    g = hdiv(h, rmax(FLT_EPSILON, w));
                    ssa = hdiv(w, rmax(FLT_EPSILON, tn));
                    tau = tn;
                }
            };
            // albedo of level k: TMEM for k < 41 (column 41 is a dummy), shared memory above (slot 0 is a dummy);
            // branch-free so the level loops stay single basic blocks
            auto st_alpha = [&](int k, FT v) {
                tmem_st1(tAl + (k < kAlphaTmemLevels ? k : kAlphaTmemLevels), v);
                alpha_hi[(k < kAlphaTmemLevels ? 0 : k - kAlphaTmemLevels + 1) * 32 + lane] = v;
            };
            // the same store when the caller knows on which side of the split level k lies (the level loops of the
            // two-stream sweeps run as two single-block loops, below and above the split: one store and its address
            // instead of two stores and four selects per cell)
            struct InTmem {}; struct InSmem {}; struct Either {};
            auto st_alpha_at = [&](auto where, int k, FT v) {
                using W_ = decltype(where);
                if constexpr (std::is_same<W_, InTmem>::value) tmem_st1(tAl + k, v);
                else if constexpr (std::is_same<W_, InSmem>::value) alpha_hi[(k - kAlphaTmemLevels + 1) * 32 + lane] = v;
                else st_alpha(k, v);
            };
            auto ld_alpha_t = [&](int k, FT& v) { tmem_ld1(tAl + (k < kAlphaTmemLevels ? k : kAlphaTmemLevels), v); };
            auto ld_alpha_s = [&](int k) -> FT { return alpha_hi[(k < kAlphaTmemLevels ? 0 : k - kAlphaTmemLevels + 1) * 32 + lane]; };
            // g-point sum of the staging tile, read transposed: lane r and lane r + 16 each add half of row r
            // (128-bit reads), one shuffle joins the halves; every lane ends with the total of row (lane & 15)
            auto row_sum = [&](FT& half) -> FT {
                const float4* row = reinterpret_cast<const float4*>(stage + (lane & 15) * kStageStride + (lane >> 4) * 16);
                const float4 a = row[0], b = row[1], c = row[2], d = row[3];
                half = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)) + (((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)));
                return half + __shfl_xor_sync(0xffffffffu, half, 16);
            };
            // sum over the warp of a per-lane value; `half` = the sum over this lane's 16-lane half (its band)
            auto warp_sum2 = [&](FT v, FT& half) -> FT {
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                half = v;
                return v + __shfl_xor_sync(0xffffffffu, v, 16);
            };
            // band accumulators: quantity q (UP / DN) of level lev gets this half's sum; callers pass lanes whose
            // (lane & 15) names a valid row, and one lane per half for warp_sum2 results
            auto band_add = [&](int q, int lev, FT half) { bacc[(hb * 2 + q) * kAccStride + lev] += half; };
            FastCell<LWG, NG> G;
            if (LW) {
                // compute_optical_props.jl:157-195 sources + longwave_2stream.jl:243-334 adding (from the bottom).
                // Iteration k gathers layer k and finishes layer k-1 (its top-level source needs pfrac of layer k).
                const FT* pbk = W.plk + bl * (nlev + 1);
                const FT emis = __ldg(P.io.sfc_emis + (size_t)col * L.n_bnd + ibnd);
                const FT inc = P.io.inc_flux_lw ? __ldg(P.io.inc_flux_lw + (size_t)gpt * P.ncol_total + col) : 0.f;
                FT tau = 0.f, ssa = 0.f, g = 0.f, pf = 0.f;
                FT lev_bot = 0.f, albedo = 1.f - emis, src = 0.f;
                // finishes layer kl = k - 1 given the Planck source at its top
                auto close_layer = [&](auto where, int kl, const LwCoef& C, FT denom, FT lev_top) {
                    const FT dB = lev_bot - lev_top;
                    const FT su = Num<FT>::pi() * (lev_top * C.emis_fac - C.q * dB);
                    const FT sd = Num<FT>::pi() * (lev_bot * C.emis_fac + C.q * dB);
                    // level kl: F_dn(kl) = A F_dn(kl+1) + B ; F_up(kl) = albedo F_dn(kl) + src
#if RB_WHATIF != 6
                    tmem_st2(tA + 2 * kl, C.Tdif * denom, (C.Rdif * src + sd) * denom);
                    st_alpha_at(where, kl, albedo);
#endif
                    const FT src_lev = src;
                    src = su + C.Tdif * denom * (src + albedo * sd);
                    albedo = C.Rdif + C.Tdif * C.Tdif * albedo * denom;
                    lev_bot = lev_top;
                    return src_lev;
                };
                for (int k0 = 0; k0 < nlay; k0 += 16) {                 // tiles of <= 16 interfaces k
                    const int ks = k0 > 0 ? k0 : 1, ke = k0 + 16 < nlay ? k0 + 16 : nlay;
                    if ((k0 & 31) == 0) build_records(k0 >> 5);             // this and the next 31 layers' records (one call site)
                    if (k0 == 0) {                                           // layer 0 and the surface
                        gather(0, G);
                        finish(G, tau, ssa, g, pf);
                        lev_bot = pbk[0] * pf;
                        src = Num<FT>::pi() * emis * (pbk[nlev] * pf);
                    }
                    // record row, McICA bit, Planck value and staging row of layer k advance by increments
                    const FT* rk = rec_lane + (ks & 31) * RR;
                    unsigned mw = HAS_CLD ? mask_word(ks) >> (ks & 31) : 0u;
                    const FT* pk = pbk + ks;
                    FT* sk = stage + lane;
                    auto level = [&](auto where, int k) {                 // single basic block
                        gather_r(rk, mw & 1u, G);
#if RB_WHATIF == 5
                        LwCoef C; C.Rdif = ssa * 0.5f; C.Tdif = 0.5f + tau * 1e-3f; C.emis_fac = 0.3f + g; C.q = tau * 1e-2f;
#else
                        const LwCoef C = lw_2stream_coeffs_nosrc(tau, ssa, g);
#endif
                        const FT denom = hrcp(1.f - C.Rdif * albedo);
                        const FT bk = *pk;
                        const FT inc_k = bk * pf;
                        finish(G, tau, ssa, g, pf);
                        const FT lev_top = hsqrt(inc_k * (bk * pf));
                        *sk = close_layer(where, k - 1, C, denom, lev_top);
                        rk += RR; mw >>= 1; ++pk; sk += kStageStride;
                    };
                    // iteration k closes level k - 1: below the split its albedo goes to tensor memory, above to shared memory
                    const int ksplit = ke < kAlphaTmemLevels + 1 ? ke : (ks > kAlphaTmemLevels + 1 ? ks : kAlphaTmemLevels + 1);
                    for (int k = ks; k < ksplit; ++k) level(InTmem{}, k);
                    for (int k = ksplit; k < ke; ++k) level(InSmem{}, k);
                    __syncwarp();
                    {                                                     // sum_g src of levels ks-1 .. ke-2
                        FT hs;
                        const FT sum = row_sum(hs);
                        if (lane < 16 && lane < ke - ks) accs[UP * kAccStride + ks - 1 + lane] += sum;
                        if (spectral && (lane & 15) < ke - ks) band_add(UP, ks - 1 + (lane & 15), hs);
                    }
                    __syncwarp();
                }
                {   // top layer: its upper source is its own increment (compute_optical_props.jl:193-195)
                    const LwCoef C = lw_2stream_coeffs_nosrc(tau, ssa, g);
                    const FT denom = hrcp(1.f - C.Rdif * albedo);
                    const FT s_top = close_layer(Either{}, nlay - 1, C, denom, pbk[nlay] * pf);
                    FT hs;
                    const FT ssum = warp_sum2(s_top, hs);
                    if (lane == 0) accs[UP * kAccStride + nlay - 1] += ssum;
                    if (spectral && (lane & 15) == 0) band_add(UP, nlay - 1, hs);
                }
                FT dn = inc;
                {
                    FT hu, hd;
                    FT u = warp_sum2(dn * albedo + src, hu), d = warp_sum2(dn, hd);
                    if (lane == 0) { accs[UP * kAccStride + nlay] += u; accs[DN * kAccStride + nlay] += d; }
                    if (spectral && (lane & 15) == 0) { band_add(UP, nlay, hu); band_add(DN, nlay, hd); }
                }
                tmem_wait_st();
                // Second sweep, 8 levels per tile.  A full tile whose albedos all sit in TMEM or all in shared memory
                // reads its (A, B) pairs with ONE tcgen05.ld.x16 (and the albedos with one .x8 or eight LDS) and has no
                // per-level TMEM round trip; the top tile of a column whose layer count is not a multiple of 8 and the
                // tile that straddles the TMEM / shared-memory albedo split take the level-by-level path.
                for (int kc = (nlay - 1) & ~7; kc >= ((RB_WHATIF == 3 || RB_WHATIF == 8) ? 1 << 20 : 0); kc -= 8) {     // 8 levels x (dn, albedo * dn) per tile
                    const int ktop = kc + 7 < nlay - 1 ? kc + 7 : nlay - 1;
                    const bool al_tmem = kc + 8 <= kAlphaTmemLevels, al_smem = kc >= kAlphaTmemLevels;
                    if (ktop == kc + 7 && (al_tmem || al_smem)) {
                        float ab[16], al8[8];
                        tmem_ld16(tA + 2 * kc, ab);
                        if (al_tmem) {
                            tmem_ld8(tAl + kc, al8);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) al8[i] = alpha_hi[(kc + i - kAlphaTmemLevels + 1) * 32 + lane];
                        }
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 7; i >= 0; --i) {
                            dn = ab[2 * i] * dn + ab[2 * i + 1];
                            stage[(i * 2 + 0) * kStageStride + lane] = dn;
                            stage[(i * 2 + 1) * kStageStride + lane] = al8[i] * dn;
                        }
                    } else {
                        FT A, B, al;
                        tmem_ld2(tA + 2 * ktop, A, B);
                        ld_alpha_t(ktop, al);
                        tmem_wait_ld();
                        for (int k = ktop; k >= kc; --k) {
                            const FT Ak = A, Bk = B;
                            const FT alk = k < kAlphaTmemLevels ? al : ld_alpha_s(k);
                            const int kn = k > 0 ? k - 1 : 0;            // prefetch the next level (harmless reload at k = 0)
                            tmem_ld2(tA + 2 * kn, A, B);
                            ld_alpha_t(kn, al);
                            dn = Ak * dn + Bk;
                            stage[((k - kc) * 2 + 0) * kStageStride + lane] = dn;
                            stage[((k - kc) * 2 + 1) * kStageStride + lane] = alk * dn;
                            tmem_wait_ld();
                        }
                    }
                    __syncwarp();
                    {
                        const int lev = kc + ((lane & 15) >> 1);
                        FT hs;
                        const FT sum = row_sum(hs);
                        if (lane < 16 && lev <= ktop) accs[((lane & 1) ? UP : DN) * kAccStride + lev] += sum;
                        if (spectral && lev <= ktop) band_add((lane & 1) ? UP : DN, lev, hs);
                    }
                    __syncwarp();
                }
            } else if (NOSCAT) {
                // compute_optical_props.jl:43-82 sources + longwave_noscat.jl:224-301, marched from the TOP so the
                // down sweep shares the pass with the gas optics: iteration j gathers layer j and steps the
                // intensities of every angle through layer j + 1 (its bottom-level source needs pfrac of layer j).
                // (tau, B_lay pfrac, source of the layer's top level) go to the level store for the up sweep.
                const FT* pbk = W.plk + bl * (2 * nlev);        // B(t_lev[0..nlay]), B(t_sfc)
                const FT* pby = pbk + nlev + 1;                 // B(t_lay[0..nlay-1])
                const FT emis = __ldg(P.io.sfc_emis + (size_t)col * L.n_bnd + ibnd);
                const FT inc = P.io.inc_flux_lw ? __ldg(P.io.inc_flux_lw + (size_t)gpt * P.ncol_total + col) : 0.f;
                FT Ia[NMU], Ds[NMU], i2f[NMU];
#pragma unroll
                for (int a = 0; a < NMU; ++a) {
                    const bool on = NMU == 1 || a < P.n_mu;
                    Ds[a] = on ? P.Ds[a] : 0.f;                 // an inactive angle: trans = 1, source = 0, weight 0
                    i2f[a] = on ? Num<FT>::pi() * P.wts[a] : 0.f;
                    Ia[a] = P.io.inc_flux_lw ? hdiv(inc, Num<FT>::pi()) : 0.f;
                }
                {
                    FT f = 0.f, hs;
#pragma unroll
                    for (int a = 0; a < NMU; ++a) f += Ia[a] * i2f[a];
                    const FT sum = warp_sum2(f, hs);
                    if (lane == 0) accs[DN * kAccStride + nlay] += sum;
                    if (spectral && (lane & 15) == 0) band_add(DN, nlay, hs);
                }
                int cur_half = (nlay - 1) >> 5;
                build_records(cur_half);
                FT tau, ssa, g, pf;
                gather(nlay - 1, G);
                finish(G, tau, ssa, g, pf);
                FT lev_top = pbk[nlay] * pf;                     // source of the top level of the layer in hand
                // down step of every angle through layer k = (tau_k, lay_k) given the source of its bottom level
                auto step_down = [&](int k, FT tau_k, FT lay_k, FT lev_bot) -> FT {
                    tmem_st2(tA + 2 * k, tau_k, lay_k);
                    st_alpha(k, lev_top);
                    FT f = 0.f;
#pragma unroll
                    for (int a = 0; a < NMU; ++a) {
                        const FT tl = tau_k * Ds[a];
                        const FT tr = hexp(-tl);
                        Ia[a] = tr * Ia[a] + lw_noscat_source(lev_bot, lay_k, tl, tr);
                        f += Ia[a] * i2f[a];
                    }
                    lev_top = lev_bot;
                    return f;
                };
                // The gathers of layer j are issued one iteration ahead, right after the registers of layer j + 1
                // are consumed, so they are in flight across the angle loop; at the first layer of a 32-layer record
                // part there is nothing to run ahead to (its records are not built yet): that iteration re-issues
                // its own loads (discarded) and the next part starts with a fresh gather.
                bool need_gather = true;
                for (int jc = (nlay - 2) & ~15; jc >= 0; jc -= 16) {   // tiles of <= 16 layers j; level j + 1 per row
                    const int jtop = jc + 15 < nlay - 2 ? jc + 15 : nlay - 2;
                    if ((jc >> 5) != cur_half) { cur_half = jc >> 5; build_records(cur_half); need_gather = true; }
                    if (need_gather) { gather(jtop, G); need_gather = false; }
                    const int part_lo = cur_half << 5;
                    for (int j = jtop; j >= jc; --j) {                  // single basic block
                        const FT tau_u = tau, lay_u = pby[j + 1] * pf;    // layer j + 1
                        const FT bk = pbk[j + 1];
                        const FT dec_u = bk * pf;
                        finish(G, tau, ssa, g, pf);                       // layer j
                        gather(j > part_lo ? j - 1 : j, G);
                        const FT lev_bot = hsqrt((bk * pf) * dec_u);      // compute_optical_props.jl:66-75
                        stage[(j - jc) * kStageStride + lane] = step_down(j + 1, tau_u, lay_u, lev_bot);
                    }
                    __syncwarp();
                    {
                        FT hs;
                        const FT sum = row_sum(hs);
                        if (lane < 16 && lane <= jtop - jc) accs[DN * kAccStride + jc + 1 + lane] += sum;
                        if (spectral && (lane & 15) <= jtop - jc) band_add(DN, jc + 1 + (lane & 15), hs);
                    }
                    __syncwarp();
                }
                {   // lowest layer, then the surface (longwave_noscat.jl:262-268)
                    FT hd, hu;
                    const FT d0 = warp_sum2(step_down(0, tau, pby[0] * pf, pbk[0] * pf), hd);
                    const FT sfc_source = pbk[nlev] * pf;
                    FT f = 0.f;
#pragma unroll
                    for (int a = 0; a < NMU; ++a) {
                        Ia[a] = Ia[a] * (1.f - emis) + emis * sfc_source;
                        f += Ia[a] * i2f[a];
                    }
                    const FT u0 = warp_sum2(f, hu);
                    if (lane == 0) { accs[DN * kAccStride] += d0; accs[UP * kAccStride] += u0; }
                    if (spectral && (lane & 15) == 0) { band_add(DN, 0, hd); band_add(UP, 0, hu); }
                }
                tmem_wait_st();
                // (no batched level-store reads here: the angle loop covers the TMEM round trip of the next level, and
                // the eight-fold unrolled angle loops cost more than they save -- measured 11.5 -> 14.2 ms with 3 angles)
                FT tk, yk, lt;
                tmem_ld2(tA, tk, yk);
                ld_alpha_t(0, lt);
                tmem_wait_ld();
                for (int kc = 0; kc < nlay; kc += 16) {                // up sweep: 16 levels k + 1 per tile
                    const int kend = kc + 16 < nlay ? kc + 16 : nlay;
                    for (int k = kc; k < kend; ++k) {
                        const FT tau_k = tk, lay_k = yk;
                        const FT lev_k1 = k < kAlphaTmemLevels ? lt : ld_alpha_s(k);
                        const int kn = k + 1 < nlay ? k + 1 : k;
                        tmem_ld2(tA + 2 * kn, tk, yk);
                        ld_alpha_t(kn, lt);
                        FT f = 0.f;
#pragma unroll
                        for (int a = 0; a < NMU; ++a) {
                            const FT tl = tau_k * Ds[a];
                            const FT tr = hexp(-tl);
                            Ia[a] = tr * Ia[a] + lw_noscat_source(lev_k1, lay_k, tl, tr);
                            f += Ia[a] * i2f[a];
                        }
                        stage[(k - kc) * kStageStride + lane] = f;
                        tmem_wait_ld();
                    }
                    __syncwarp();
                    {
                        FT hs;
                        const FT sum = row_sum(hs);
                        if (lane < 16 && kc + lane < kend) accs[UP * kAccStride + kc + 1 + lane] += sum;
                        if (spectral && kc + (lane & 15) < kend) band_add(UP, kc + 1 + (lane & 15), hs);
                    }
                    __syncwarp();
                }
            } else {
                // shortwave_2stream.jl:300-392 with the adding marched from the top.  Iteration j gathers layer j
                // and processes layer j + 1 with the optics finished one iteration earlier.
                const FT alb_dir = __ldg(P.io.sfc_alb_direct + (size_t)col * L.n_bnd + ibnd);
                const FT alb_dif = __ldg(P.io.sfc_alb_diffuse + (size_t)col * L.n_bnd + ibnd);
                const FT dir_top = toa * __ldg(L.solar_src_scaled + gpt) * mu0;
                const FT inv_mu0 = hdiv(1.f, rmax(mu0, FLT_EPSILON));
                const FT neg_inv_mu0_l2e = -inv_mu0 * 1.4426950408889634f;
                FT tau_cum = 0.f, dir = dir_top;
                FT beta = 0.f, d = 0.f;   // reflectance / downward diffuse source of everything above the level
                if (day) {
                    FT hs;
                    FT sum = warp_sum2(dir_top, hs);   // TOA: diffuse incident flux is zero (shortwave_2stream.jl:331)
                    if (lane == 0) { accs[DIR * kAccStride + nlay] += sum; accs[DN * kAccStride + nlay] += sum; }
                    if (spectral && (lane & 15) == 0) band_add(DN, nlay, hs);
                }
                FT tau = 0.f, ssa = 0.f, g = 0.f, pf = 0.f;
                // layer k: coefficients, TMEM store, marching update; returns d_{k+1} (before the update)
                auto march = [&](auto where, int k) -> FT {
                    FT Rdir, Tdir, Rdif, Tdif;
#if RB_WHATIF == 5
                    Rdir = ssa * 0.3f; Tdir = 0.2f + tau * 1e-3f; Rdif = g * 0.1f + 0.1f; Tdif = 0.5f + tau * 1e-3f;
#else
                    sw_2stream_coeffs(tau, ssa, g, mu0, inv_mu0, Rdir, Tdir, Rdif, Tdif);
#endif
                    const FT su = Rdir * dir, sd = Tdir * dir;       // dir = direct flux at level k+1
                    const FT denom = hrcp(1.f - Rdif * beta);
                    // F_up(k+1) = A'_k F_up(k) + B'_k ; F_dn_dif(k+1) = beta_{k+1} F_up(k+1) + d_{k+1}
#if RB_WHATIF != 6
                    tmem_st2(tA + 2 * k, Tdif * denom, (Rdif * d + su) * denom);
                    st_alpha_at(where, k, beta);
#endif
                    const FT d_above = d;
                    d = sd + Tdif * denom * (d + beta * su);
                    beta = Rdif + Tdif * Tdif * beta * denom;
                    tau_cum += tau;
                    float ex;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(tau_cum * neg_inv_mu0_l2e));
                    dir = dir_top * ex;                               // direct flux at level k
                    return d_above;
                };
                int part_built = -1;
                // 8 layers x (d_{k+1}, dir_k) per tile, k = j + 1; the first pass only gathers the top layer
                for (int jc = ((nlay - 2) & ~7) + 8; jc >= 0; jc -= 8) {
                    const bool top_only = jc > nlay - 2;
                    const int part = top_only ? (nlay - 1) >> 5 : jc >> 5;
                    if (part != part_built) { build_records(part); part_built = part; }   // the one call site
                    if (!day) continue;                                  // night: records (AOD) only
                    if (top_only) {
                        gather(nlay - 1, G);
                        finish(G, tau, ssa, g, pf);
                        continue;
                    }
                    const int jtop = jc + 7 < nlay - 2 ? jc + 7 : nlay - 2;
                    const FT* rk = rec_lane + (jtop & 31) * RR;
                    unsigned mw = HAS_CLD ? mask_word(jtop) << (31 - (jtop & 31)) : 0u;   // bit 31 = layer j
                    FT* sk = stage + ((jtop - jc) * 2) * kStageStride + lane;
                    auto level = [&](auto where, int j) {               // single basic block
                        gather_r(rk, (mw >> 31) & 1u, G);
                        sk[0] = march(where, j + 1);
                        sk[kStageStride] = dir;
                        finish(G, tau, ssa, g, pf);
                        rk -= RR; mw <<= 1; sk -= 2 * kStageStride;
                    };
                    // iteration j stores level j + 1: shared memory at and above the split, tensor memory below
                    const int jsplit = jc > kAlphaTmemLevels - 1 ? jc : (jtop + 1 < kAlphaTmemLevels - 1 ? jtop + 1 : kAlphaTmemLevels - 1);
                    for (int j = jtop; j >= jsplit; --j) level(InSmem{}, j);
                    for (int j = (jsplit - 1 < jtop ? jsplit - 1 : jtop); j >= jc; --j) level(InTmem{}, j);
                    __syncwarp();
                    {
                        const int kk = jc + 1 + ((lane & 15) >> 1);
                        FT hs;
                        const FT sum = row_sum(hs);
                        const bool okb = kk <= jtop + 1, ok = lane < 16 && okb;
                        // d_{kk+1} (even lanes) and dir_kk (odd lanes) both feed F_dn: two ordered steps,
                        // never two lanes read-modify-writing one accumulator in the same instruction
                        if (ok && (lane & 1)) { accs[DN * kAccStride + kk] += sum; accs[DIR * kAccStride + kk] += sum; }
                        if (spectral && okb && (lane & 1)) band_add(DN, kk, hs);
                        __syncwarp();
                        if (ok && !(lane & 1)) accs[DN * kAccStride + kk + 1] += sum;
                        if (spectral && okb && !(lane & 1)) band_add(DN, kk + 1, hs);
                    }
                    __syncwarp();
                }
                if (!day) {   // night: the 550 nm optical depths are all this block produces
                    aod_e = warp_sum(aod_e); aod_s = warp_sum(aod_s);
                    if (lane == 0) { P.io.aod_ext[col] = aod_e; P.io.aod_sca[col] = aod_s; }
                    continue;
                }
                {   // lowest layer
                    FT hd1, hdir0;
                    const FT d1 = warp_sum2(march(Either{}, 0), hd1), dir0 = warp_sum2(dir, hdir0);
                    if (lane == 0) { accs[DN * kAccStride + 1] += d1; accs[DN * kAccStride] += dir0; accs[DIR * kAccStride] += dir0; }
                    if (spectral && (lane & 15) == 0) { band_add(DN, 1, hd1); band_add(DN, 0, hdir0); }
                }
                if (aod_here) {
                    aod_e = warp_sum(aod_e); aod_s = warp_sum(aod_s);
                    if (lane == 0) { P.io.aod_ext[col] = aod_e; P.io.aod_sca[col] = aod_s; }
                }
                // surface: F_up(0) = alb_dif F_dn_dif(0) + alb_dir dir(0) ; F_dn_dif(0) = d_0 + beta_0 F_up(0)
                FT up = hdiv(alb_dif * d + alb_dir * dir, 1.f - alb_dif * beta);
                {
                    FT hu, hdd;
                    FT u = warp_sum2(up, hu), dd = warp_sum2(d + beta * up, hdd);
                    if (lane == 0) { accs[UP * kAccStride] += u; accs[DN * kAccStride] += dd; }
                    if (spectral && (lane & 15) == 0) { band_add(UP, 0, hu); band_add(DN, 0, hdd); }
                }
                tmem_wait_st();
                for (int kc = ((RB_WHATIF == 3 || RB_WHATIF == 8) ? 1 << 20 : 0); kc < nlay; kc += 8) {                // 8 levels x (F_up, beta * F_up) per tile
                    const int kend = kc + 8 < nlay ? kc + 8 : nlay;
                    const bool be_tmem = kc + 8 <= kAlphaTmemLevels, be_smem = kc >= kAlphaTmemLevels;
                    if (kend == kc + 8 && (be_tmem || be_smem)) {       // one tcgen05.ld.x16 (+ .x8) per tile, as in the LW sweep
                        float ab[16], be8[8];
                        tmem_ld16(tA + 2 * kc, ab);
                        if (be_tmem) {
                            tmem_ld8(tAl + kc, be8);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) be8[i] = alpha_hi[(kc + i - kAlphaTmemLevels + 1) * 32 + lane];
                        }
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            up = ab[2 * i] * up + ab[2 * i + 1];              // F_up(k+1)
                            stage[(i * 2 + 0) * kStageStride + lane] = up;
                            stage[(i * 2 + 1) * kStageStride + lane] = be8[i] * up;
                        }
                    } else {
                        FT A, B, be;
                        tmem_ld2(tA + 2 * kc, A, B);
                        ld_alpha_t(kc, be);
                        tmem_wait_ld();
                        for (int k = kc; k < kend; ++k) {
                            const FT Ak = A, Bk = B;
                            const FT bek = k < kAlphaTmemLevels ? be : ld_alpha_s(k);
                            const int kn = k + 1 < nlay ? k + 1 : k;
                            tmem_ld2(tA + 2 * kn, A, B);
                            ld_alpha_t(kn, be);
                            up = Ak * up + Bk;                                  // F_up(k+1)
                            stage[((k - kc) * 2 + 0) * kStageStride + lane] = up;
                            stage[((k - kc) * 2 + 1) * kStageStride + lane] = bek * up;
                            tmem_wait_ld();
                        }
                    }
                    __syncwarp();
                    {
                        const int kk = kc + ((lane & 15) >> 1);
                        FT hs;
                        const FT sum = row_sum(hs);
                        if (lane < 16 && kk < kend) accs[((lane & 1) ? DN : UP) * kAccStride + kk + 1] += sum;
                        if (spectral && kk < kend) band_add((lane & 1) ? DN : UP, kk + 1, hs);
                    }
                    __syncwarp();
                }
            }
            if (spectral) flush_bands(false);
        }
        __syncwarp();

        if (split_log2 > 0) {   // warp-uniform
            float* scr = P.split_scratch + ((size_t)col << split_log2) * kScr;
            float* mine = scr + share * kScr;
            for (int i = lane; i < 3 * kAccStride; i += 32) __stcg(mine + i, accs[i]);
            if (lane == 0) __stcg(mine + 3 * kAccStride, (float)n_cloudy);
            // Ordering without __threadfence(): that is MEMBAR.SC.GPU + CCTL.IVALL, and invalidating the SM's L1 twice per
            // work item throws the k-distribution gathers of all 12 warps out of the cache.  The partial sums go around
            // L1 (st.cg / ld.cg); the warp's stores are ordered before lane 0's RELEASE increment (MEMBAR.ALL.GPU, no
            // invalidation) by the warp barrier; the last arriver's loads depend on the counter value through the branch.
            __syncwarp();
            unsigned prev = 0;
            if (lane == 0)
                asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(P.split_flags + col) : "memory");
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev != (1u << split_log2) - 1u) continue;          // another share of this column is still on its way
            if (lane == 0) P.split_flags[col] = 0u;                  // ready for the next launch
            for (int i = lane; i < 3 * kAccStride; i += 32) {
                float t = 0.f;
                for (int sh = 0; sh < (1 << split_log2); ++sh) t += __ldcg(scr + sh * kScr + i);
                accs[i] = t;
            }
            float nc_f = 0.f;
            for (int sh = 0; sh < (1 << split_log2); ++sh) nc_f += __ldcg(scr + sh * kScr + 3 * kAccStride);
            n_cloudy = (int)nc_f;
            __syncwarp();
        }

        // ---------------- epilogue: (nlev, ncol) presentation, net, scaling, diagnostics ----------------
#pragma unroll
        for (int i = 0; i < (kMaxLay + 32) / 32; ++i) {
            const int lev = lane + 32 * i;
            if (lev < nlev) {
                const size_t o = (size_t)col * nlev + lev;
                FT up = accs[UP * kAccStride + lev], dn = accs[DN * kAccStride + lev], dr = accs[DIR * kAccStride + lev];
                if (!day) { up = dn = dr = 0.f; }
                FT net = up - dn;
                if (P.io.metric_scaling != nullptr) {
                    FT sc = __ldg(P.io.metric_scaling + o);
                    up *= sc; dn *= sc; net *= sc; dr *= sc;
                }
                P.io.out_up[o] = up; P.io.out_dn[o] = dn; P.io.out_net[o] = net;
                if (!LWG) P.io.out_dir[o] = dr;
                if (P.io.out_total_net != nullptr) P.io.out_total_net[o] = P.io.add_net[o] + net;
            }
        }
        if (lane == 0 && P.io.cld_cover != nullptr && HAS_CLD) P.io.cld_cover[col] = __fdiv_rn(FT(n_cloudy), FT(NGPT));
        __syncwarp();
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_smem, 512u);
}

}  // namespace rb

/* rrtmgp_b200.h -- C ABI of the B200-native column-radiation engine.
 *
 * Drop-in boundary for the `update_fluxes!` hot path of CliMA/RRTMGP.jl (reference v1.0.0;
 * citations are file:line in the reference tree).  The reference reaches its device code
 * through Julia multiple dispatch, not a plugin ABI; the seam this library replaces is the
 * set of methods `ext/RRTMGPCUDAExt.jl:33-45` specialises on the CUDA device
 * (`rte_lw_2stream_solve!`, `rte_sw_2stream_solve!`, `rte_lw_noscat_solve!`,
 * `compute_col_gas!`) plus the Layer-2 orchestration around them
 * (`src/api/update_fluxes.jl:223-281`).  A Julia host `ccall`s these entry points with the
 * `CuPtr`s of the arrays its `RRTMGPSolver` already owns; in this
 * repository the same ABI is driven through ctypes + torch device tensors.
 *
 * Conventions
 *   - every function returns an `rrtmgp_b200_status` (0 = ok); nothing throws, nothing
 *     allocates after `create`/`load_luts`, nothing synchronises the device (contract of
 *     src/api/update_fluxes.jl:215-218);
 *   - all array pointers are DEVICE pointers owned by the caller, element type `float` or
 *     `double` per `config.dtype`;
 *   - layouts are the reference's: Julia `(vertical, ncol)` column-major == C `[ncol][vertical]`
 *     (SURVEY.md Appendix B); level/layer 1 is the surface (src/rte/longwave_2stream.jl:262-267);
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 */
#ifndef RRTMGP_B200_H
#define RRTMGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRTMGP_B200_ABI_VERSION 1

typedef enum {
    RRTMGP_B200_OK = 0,
    RRTMGP_B200_ERR_INVALID_ARG = 1,   /* bad config / null pointer / unsupported combination */
    RRTMGP_B200_ERR_BAD_LUT_PACK = 2,  /* magic/version/size mismatch or missing table */
    RRTMGP_B200_ERR_NOT_READY = 3,     /* update called before load_luts / bind */
    RRTMGP_B200_ERR_CUDA = 4,          /* a CUDA runtime call failed (see rrtmgp_b200_last_cuda_error) */
    RRTMGP_B200_ERR_UNSUPPORTED = 5    /* e.g. nlay too large for the kernel's register tiling */
} rrtmgp_b200_status;

/* radiation_methods.jl:19,30,49,67 (GrayRadiation has no g-point axis and stays on the host) */
enum { RRTMGP_B200_CLEAR_SKY = 0, RRTMGP_B200_ALL_SKY = 1, RRTMGP_B200_ALL_SKY_WITH_CLEAR = 2 };
/* optical_props.jl:13,48 */
enum { RRTMGP_B200_TWO_STREAM = 0, RRTMGP_B200_ONE_SCALAR = 1 };
/* VolumeMixingRatios.jl:34-43 (VmrGM) and :75-78 (Vmr) */
enum { RRTMGP_B200_VMR_GM = 0, RRTMGP_B200_VMR_FULL = 1 };

/* AbstractInterpolation / AbstractBottomExtrapolation (src/api/interpolation.jl:40-135) */
enum { RRTMGP_B200_NO_INTERPOLATION = 0, RRTMGP_B200_ARITHMETIC_MEAN = 1, RRTMGP_B200_GEOMETRIC_MEAN = 2,
       RRTMGP_B200_UNIFORM_Z = 3, RRTMGP_B200_UNIFORM_P = 4, RRTMGP_B200_BEST_FIT = 5 };
enum { RRTMGP_B200_SAME_AS_INTERPOLATION = 0, RRTMGP_B200_USE_SURFACE_TEMP_AT_BOTTOM = 1, RRTMGP_B200_HYDROSTATIC_BOTTOM = 2 };

/* Mirrors what `RRTMGPSolver(grid_params, radiation_method, params, ...; op_lw, n_gauss_angles,
 * spectral_fluxes, deep_atmosphere_inverse_scaling)` fixes at construction
 * (src/api/solver.jl:136-331, src/api/grid_params.jl:38-54, src/Parameters.jl:6-14). */
typedef struct {
    int32_t abi_version;        /* RRTMGP_B200_ABI_VERSION */
    int32_t device;             /* CUDA device ordinal */
    int32_t dtype;              /* 0 = Float32, 1 = Float64 */
    int32_t ncol;
    int32_t nlay;               /* layers incl. the isothermal boundary layer, if any */
    int32_t ngas;               /* length of the gas axis of `vmr` */
    int32_t vmr_kind;           /* RRTMGP_B200_VMR_* */
    int32_t method;             /* RRTMGP_B200_CLEAR_SKY / ALL_SKY / ALL_SKY_WITH_CLEAR */
    int32_t aerosol_radiation;  /* 0/1 */
    int32_t op_lw;              /* TWO_STREAM or ONE_SCALAR (no-scattering LW) */
    int32_t n_gauss_angles;     /* 1..4, ONE_SCALAR only (solver.jl:159-171) */
    int32_t ice_rgh;            /* 1..3 (AtmosphericStates.jl:236-248) */
    int32_t spectral_fluxes;    /* 0/1: also fill per-band fluxes (Fluxes.jl:170-215) */
    int32_t isothermal_boundary_layer; /* 0/1: top layer is filled by prepare (grid_adaptation.jl:135-150) */
    int64_t col_offset;         /* global index of column 0 of this handle (column shards; keys McICA) */
    double grav, molmass_dryair, molmass_water, avogad; /* Parameters.jl:6-14 */
} rrtmgp_b200_config_t;

/* Device pointers of the arrays the reference's solver owns (src/api/solver.jl:216-272).
 * Optional pointers may be NULL.  nlev = nlay + 1. */
typedef struct {
    /* --- AtmosphericState (AtmosphericStates.jl:70-81), mutated in place by prepare --- */
    void* layerdata;      /* [ncol][nlay][4] (col_dry, p_lay, t_lay, rel_hum) */
    void* p_lev;          /* [ncol][nlev] */
    void* t_lev;          /* [ncol][nlev] */
    void* t_sfc;          /* [ncol] */
    void* vmr_h2o;        /* [ncol][nlay]            (VMR_GM) */
    void* vmr_o3;         /* [ncol][nlay]            (VMR_GM) */
    void* vmr;            /* VMR_GM: [ngas]; VMR_FULL: [ncol][nlay][ngas] */
    void* lat;            /* [ncol] degrees or NULL (gas_optics.jl:29-33) */
    /* CloudState (AtmosphericStates.jl:236-248); all NULL when there is no cloud state */
    void* cld_r_eff_liq;  /* [ncol][nlay] */
    void* cld_r_eff_ice;
    void* cld_path_liq;
    void* cld_path_ice;
    void* cld_frac;
    void* cld_cover_lw;   /* [ncol] out, optional */
    void* cld_cover_sw;   /* [ncol] out, optional */
    /* AerosolState (AtmosphericStates.jl:292-298); NULL when there is no aerosol state */
    void* aero_mass;      /* [ncol][nlay][15] */
    void* aero_size;      /* [ncol][nlay][15] */
    void* aod_sw_ext;     /* [ncol] out, optional */
    void* aod_sw_sca;     /* [ncol] out, optional */
    /* --- boundary conditions (BCs.jl:17-23,40-47) --- */
    void* sfc_emis;       /* [ncol][n_bnd_lw] */
    void* inc_flux_lw;    /* [n_gpt_lw][ncol] or NULL */
    void* cos_zenith;     /* [ncol] */
    void* toa_flux;       /* [ncol] */
    void* sfc_alb_direct; /* [ncol][n_bnd_sw] */
    void* sfc_alb_diffuse;/* [ncol][n_bnd_sw] */
    void* metric_scaling; /* [ncol][nlev] deep_atmosphere_inverse_scaling or NULL (Fluxes.jl:295-304) */
    /* --- outputs: the (nlev, ncol) presentation arrays the getters expose (Fluxes.jl:355-374) --- */
    void* lw_flux_up;  void* lw_flux_dn;  void* lw_flux_net;                      /* [ncol][nlev] */
    void* sw_flux_up;  void* sw_flux_dn;  void* sw_flux_net;  void* sw_flux_dn_dir;
    void* net_flux;                                                               /* lw_net + sw_net */
    /* clear-sky snapshots, ALL_SKY_WITH_CLEAR only (update_fluxes.jl:39-65,101-128) */
    void* clear_lw_flux_up; void* clear_lw_flux_dn; void* clear_lw_flux_net;
    void* clear_sw_flux_up; void* clear_sw_flux_dn; void* clear_sw_flux_net; void* clear_sw_flux_dn_dir;
    void* clear_net_flux;
    /* per-band fluxes, spectral_fluxes only: Julia (nlev, ncol, n_bnd) == C [n_bnd][ncol][nlev] */
    void* lw_band_flux_up; void* lw_band_flux_dn; void* lw_band_flux_net;
    void* sw_band_flux_up; void* sw_band_flux_dn; void* sw_band_flux_net;
} rrtmgp_b200_buffers_t;

/* Table dimensions and the scalars a host needs before allocating (LookUpTables.jl:130-143,185-201). */
typedef struct {
    int32_t n_gpt_lw, n_bnd_lw, n_gpt_sw, n_bnd_sw, ngas, iband_550nm;
    double p_ref_min, t_ref_min, t_ref_max, solar_src_tot;
} rrtmgp_b200_lut_info_t;

typedef struct rrtmgp_b200_handle rrtmgp_b200_handle_t;

/* RRTMGPSolver(...) construction: validates the option combination (solver.jl:159-193) and
 * sizes every internal workspace.  One handle per GPU / per column shard. */
int rrtmgp_b200_create(const rrtmgp_b200_config_t* cfg, rrtmgp_b200_handle_t** out);
void rrtmgp_b200_destroy(rrtmgp_b200_handle_t* h);

/* lookup_tables(grid_params, method) (ext/RRTMGPNCDatasetsExt.jl:93-133): `pack` is a HOST
 * pointer to a flat LUT pack (rrtmgp.jl_b200/lutpack.py; built from the rrtmgp-data NetCDF files by
 * rrtmgp.jl_b200/tables.py or from a loaded LookupBundle by `lut_pack` in julia/RRTMGPB200Ext.jl); tables
 * are converted to `dtype`, re-laid out g-point-fastest and uploaded.  The cloud (`cld_lw/`, `cld_sw/`) and
 * aerosol (`aero_lw/`, `aero_sw/`) sections may be absent when the handle's method never reads them (clear
 * sky / no aerosol radiation, as `lookup_tables` loads them); otherwise RRTMGP_B200_ERR_BAD_LUT_PACK. */
int rrtmgp_b200_load_luts(rrtmgp_b200_handle_t* h, const void* pack, size_t nbytes);
int rrtmgp_b200_lut_info(const rrtmgp_b200_handle_t* h, rrtmgp_b200_lut_info_t* out);

/* Registers the caller-owned device arrays (the struct is copied). */
int rrtmgp_b200_bind(rrtmgp_b200_handle_t* h, const rrtmgp_b200_buffers_t* bufs);

/* The `interpolation` / `bottom_extrapolation` / `center_z` / `face_z` keywords of RRTMGPSolver (solver.jl:136-147,
 * 183-193): from now on prepare_atmosphere! first fills p_lev / t_lev of the domain faces from the layer values
 * (interpolate_levels!, grid_adaptation.jl:87-113; interp! / extrap!, interpolation.jl:176-252).  `center_z`
 * [ncol][nlay] and `face_z` [ncol][nlev] are caller-owned device arrays, required by BEST_FIT and HYDROSTATIC_BOTTOM
 * (else NULL); cp_d, R_d are Parameters.cp_d / R_d, used by the two bottom-only extrapolations. */
int rrtmgp_b200_set_level_interpolation(rrtmgp_b200_handle_t* h, int32_t interpolation, int32_t bottom_extrapolation,
                                        const void* center_z, const void* face_z, double cp_d, double R_d);

/* prepare_atmosphere!(s) (update_fluxes.jl:252-281): level interpolation, boundary layer fill, clip!, col_dry; in place. */
int rrtmgp_b200_prepare_atmosphere(rrtmgp_b200_handle_t* h, void* stream);

/* The steps of prepare_atmosphere! one by one, for hosts that call the public functions of
 * src/api/grid_adaptation.jl themselves; `steps` is an OR of the bits below and the selected steps run in
 * prepare_atmosphere!'s order.  INTERPOLATE_LEVELS is a no-op unless set_level_interpolation chose a scheme, and
 * BOUNDARY_LAYER unless the handle has an isothermal boundary layer (as in update_fluxes.jl:256-270). */
typedef enum {
    RRTMGP_B200_STEP_INTERPOLATE_LEVELS = 1,   /* interpolate_levels!            grid_adaptation.jl:87-113  */
    RRTMGP_B200_STEP_BOUNDARY_LAYER = 2,       /* add_isothermal_boundary_layer! grid_adaptation.jl:135-195 */
    RRTMGP_B200_STEP_CLIP = 4,                 /* clip!                          grid_adaptation.jl:232-258 */
    RRTMGP_B200_STEP_CONCENTRATIONS = 8        /* update_concentrations!         grid_adaptation.jl:278-293 */
} rrtmgp_b200_prepare_step;
int rrtmgp_b200_prepare_steps(rrtmgp_b200_handle_t* h, uint32_t steps, void* stream);
/* update_lw_fluxes!(s) / update_sw_fluxes!(s) / update_net_fluxes!(s) (update_fluxes.jl:12-16,74-78,165-194).
 * `have_seed == 0` mirrors `seedval = nothing`: an internal per-call counter keys the McICA draws. */
int rrtmgp_b200_update_lw_fluxes(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, void* stream);
int rrtmgp_b200_update_sw_fluxes(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, void* stream);
int rrtmgp_b200_update_net_fluxes(rrtmgp_b200_handle_t* h, void* stream);
/* update_fluxes!(s, seedval) (update_fluxes.jl:223-233) = prepare + lw + sw + net, async on `stream`. */
int rrtmgp_b200_update_fluxes(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, void* stream);

/* The same update restricted to columns [col_begin, col_begin + col_count): lets a host pipeline
 * H2D copy / compute / D2H copy of column chunks on several streams (the reference has no such call; its
 * columns are independent, ext/cuda/rte_longwave_2stream.jl:101).  Requires an explicit seed so that all
 * chunks of one step sample consistently; results are identical to one full-range call. */
int rrtmgp_b200_update_fluxes_range(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, int64_t col_begin,
                                    int32_t col_count, void* stream);

/* compute_relative_humidity!(...) (src/optics/column_amounts.jl:52-76, gas_optics.jl:58-80): a host
 * duty in the reference (grid_adaptation.jl:267-270); writes layerdata[..][3]. */
int rrtmgp_b200_compute_relative_humidity(rrtmgp_b200_handle_t* h, void* stream);

/* heating_rate(s) (src/api/standalone.jl:106-124, GrayAtmosphere.jl:152-167): (grav / cp_d) dF_net/dp of the domain
 * layers from any [ncol][nlev] net-flux array and the bound p_lev; `heating_rate` is [ncol][domain nlay]. */
int rrtmgp_b200_heating_rate(rrtmgp_b200_handle_t* h, const void* flux_net, void* heating_rate, double cp_d, void* stream);

/* Diagnostics: kernels launched by the last update_* call; last CUDA error string. */
/* validate_inputs(s) (src/api/validation.jl:56-74; opt-in through `check_values[]`, :14): every bound input
 * against its physical range -- pressures / temperatures positive and finite, cos_zenith in [-1, 1], TOA flux
 * >= 0, emissivity and albedos in [0, 1], mixing ratios >= 0.  `*failed` receives an OR of the bits below
 * (0 = all good).  Unlike every other entry point this one SYNCHRONISES `stream` (it returns a host value), as
 * the reference's reduction does. */
typedef enum {
    RRTMGP_B200_BAD_LEVEL_PRESSURE = 1 << 0,
    RRTMGP_B200_BAD_LEVEL_TEMPERATURE = 1 << 1,
    RRTMGP_B200_BAD_LAYER_PRESSURE = 1 << 2,
    RRTMGP_B200_BAD_LAYER_TEMPERATURE = 1 << 3,
    RRTMGP_B200_BAD_SURFACE_TEMPERATURE = 1 << 4,
    RRTMGP_B200_BAD_COS_ZENITH = 1 << 5,
    RRTMGP_B200_BAD_TOA_SW_FLUX_DN = 1 << 6,
    RRTMGP_B200_BAD_SURFACE_EMISSIVITY = 1 << 7,
    RRTMGP_B200_BAD_DIRECT_SW_SURFACE_ALBEDO = 1 << 8,
    RRTMGP_B200_BAD_DIFFUSE_SW_SURFACE_ALBEDO = 1 << 9,
    RRTMGP_B200_BAD_VMR_H2O = 1 << 10,
    RRTMGP_B200_BAD_VMR_O3 = 1 << 11,
    RRTMGP_B200_BAD_VMR = 1 << 12
} rrtmgp_b200_invalid_input;
int rrtmgp_b200_validate_inputs(rrtmgp_b200_handle_t* h, uint32_t* failed, void* stream);

/* ---- multi-GPU: column shards across the ranks of one box, one process per GPU ----
 * The reference has no multi-GPU path (docs/src/howto/gpu.md:69-84: one device per process, the host model
 * decomposes the columns); what a sharded host needs beyond its own shard is the concatenation of the presented
 * (nlev, ncol) flux views (getters.jl:320-470), i.e. one all-gather per view.  Because `ncol` is the slowest axis
 * of every view, rank r's result is rows [r ncol, (r + 1) ncol) of the gathered array.
 *
 *   comm_unique_id   rank 0 creates the NCCL id; the host ships its 128 bytes to every rank (MPI_Bcast, a file, ...).
 *   comm_init        joins the communicator (libnccl.so.2 is dlopen'ed: the library has no link-time dependency),
 *                    allocates the gathered arrays ([nranks ncol][nlev] each, 8 of them: rrtmgp_b200_gathered_t),
 *                    and opens every peer's copy through CUDA IPC so results can be pushed over NVLink by the copy
 *                    engines.  Every rank must use the same ncol, nlay and dtype.
 *   update_fluxes_gathered   update_fluxes! + the gather, overlapped: the three longwave views travel while the
 *                    shortwave kernel runs, the shortwave / net views per column chunk as it finishes (chunks shrink so
 *                    that only a short last push is exposed); copies are cudaMemcpyAsync on four side streams (DMA
 *                    engines -- the persistent kernels leave no SM for a collective kernel), framed by two one-element NCCL all-reduces (nobody still reads the previous
 *                    step's arrays / everybody's pushes have landed).  On return (stream order) every rank holds all
 *                    columns.
 *   all_gather_fluxes   the plain alternative: one grouped ncclAllGather of the eight views as they are now. */
typedef struct {
    void *lw_flux_up, *lw_flux_dn, *lw_flux_net, *sw_flux_up, *sw_flux_dn, *sw_flux_net, *sw_flux_dn_dir, *net_flux;
} rrtmgp_b200_gathered_t;
#define RRTMGP_B200_UNIQUE_ID_BYTES 128
int rrtmgp_b200_comm_unique_id(void* id_out, size_t nbytes);
int rrtmgp_b200_comm_init(rrtmgp_b200_handle_t* h, const void* unique_id, int32_t rank, int32_t nranks);
int rrtmgp_b200_gathered_buffers(const rrtmgp_b200_handle_t* h, rrtmgp_b200_gathered_t* out);
int rrtmgp_b200_update_fluxes_gathered(rrtmgp_b200_handle_t* h, uint64_t seed, int have_seed, void* stream);
int rrtmgp_b200_all_gather_fluxes(rrtmgp_b200_handle_t* h, void* stream);
int rrtmgp_b200_comm_destroy(rrtmgp_b200_handle_t* h);

/* Measurement aid (no reference counterpart): FP32 CUDA-core peak of `device` from a stream of
 * independent scalar FFMA and of packed FFMA2, in TFLOP/s -- the denominator bench.py reports `roofline_fp32` against.
 * Synchronous; a few milliseconds. */
int rrtmgp_b200_measure_fp32_peak(int32_t device, double* ffma_tflops, double* ffma2_tflops);

int rrtmgp_b200_last_launch_count(const rrtmgp_b200_handle_t* h);
const char* rrtmgp_b200_last_cuda_error(const rrtmgp_b200_handle_t* h);
const char* rrtmgp_b200_strerror(int status);
int rrtmgp_b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RRTMGP_B200_H */

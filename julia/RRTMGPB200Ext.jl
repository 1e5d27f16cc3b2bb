# RRTMGPB200Ext.jl -- reference-side binding of librrtmgp_b200.so (include/rrtmgp_b200.h).
#
# UNTESTED HERE: this image has no Julia.  The same C ABI is exercised through ctypes by
# rrtmgp.jl_b200/_lib.py + solver.py, from which this file is mechanically derived.
#
# Drop-in seam: the same methods `ext/RRTMGPCUDAExt.jl:33-45` specialises on the CUDA device,
# plus the Layer-2 orchestration of `src/api/update_fluxes.jl:223-281`.  A host keeps its
# `RRTMGPSolver` (all arrays stay owned by Julia, `src/api/solver.jl:216-272`); `B200Engine`
# wraps one library handle bound to the `CuPtr`s of those arrays.  After `bind!`, replace
#     RRTMGP.update_fluxes!(solver, seed)      by      update_fluxes!(engine, seed)
# and keep reading results through the usual getters (`net_flux(solver)`, ...): the engine
# writes straight into the `(nlev, ncol)` presentation arrays the getters expose
# (`src/optics/Fluxes.jl:355-374`), so the `(ncol, nlev)` compute buffers, the transposed
# state cache and the presentation copies of the reference are simply unused.
module RRTMGPB200Ext

using CUDA
import Adapt
import RRTMGP
import RRTMGP: RRTMGPSolver

const LIB = get(ENV, "RRTMGP_B200_LIB", "librrtmgp_b200.so")

# rrtmgp_b200_config_t (include/rrtmgp_b200.h)
struct Config
    abi_version::Int32; device::Int32; dtype::Int32; ncol::Int32; nlay::Int32; ngas::Int32
    vmr_kind::Int32; method::Int32; aerosol_radiation::Int32; op_lw::Int32; n_gauss_angles::Int32
    ice_rgh::Int32; spectral_fluxes::Int32; isothermal_boundary_layer::Int32
    col_offset::Int64
    grav::Float64; molmass_dryair::Float64; molmass_water::Float64; avogad::Float64
end

# rrtmgp_b200_buffers_t: 48 device pointers in header order
const NBUF = 48
mutable struct Buffers
    p::NTuple{NBUF, CuPtr{Cvoid}}
end

mutable struct B200Engine
    handle::Ptr{Cvoid}
end

const ABI_VERSION = 1
function __init__()
    v = ccall((:rrtmgp_b200_abi_version, LIB), Cint, ())
    v == ABI_VERSION || error("librrtmgp_b200.so has ABI version $v, this binding was written for $ABI_VERSION")
end

check(st::Cint) = st == 0 || error(unsafe_string(ccall((:rrtmgp_b200_strerror, LIB), Cstring, (Cint,), st)))
# with the handle at hand a CUDA failure carries the runtime's own message (rrtmgp_b200_last_cuda_error)
function check(st::Cint, handle::Ptr{Cvoid})
    st == 0 && return true
    msg = unsafe_string(ccall((:rrtmgp_b200_strerror, LIB), Cstring, (Cint,), st))
    st == 4 && (msg *= ": " * unsafe_string(ccall((:rrtmgp_b200_last_cuda_error, LIB), Cstring, (Ptr{Cvoid},), handle)))
    error(msg)
end

# rrtmgp_b200_lut_info_t: what `load_luts` found (the host sizes its boundary-condition arrays by these)
struct LutInfo
    n_gpt_lw::Int32; n_bnd_lw::Int32; n_gpt_sw::Int32; n_bnd_sw::Int32; ngas::Int32; iband_550nm::Int32
    p_ref_min::Float64; t_ref_min::Float64; t_ref_max::Float64; solar_src_tot::Float64
end

devptr(::Nothing) = CuPtr{Cvoid}(0)
devptr(a) = reinterpret(CuPtr{Cvoid}, pointer(parent(a)))   # parent(): getters hand out SubArray views

method_code(::RRTMGP.ClearSkyRadiation) = 0
method_code(::RRTMGP.AllSkyRadiation) = 1
method_code(::RRTMGP.AllSkyRadiationWithClearSkyDiagnostics) = 2

"""
    B200Engine(s::RRTMGPSolver, lut_pack::Vector{UInt8}; col_offset = 0)

`lut_pack` is the flat table pack; produce it once from `Adapt.adapt(Array, s.lookups)` with
`write_lut_pack` below (array-by-array dump in the post-load layouts of `src/optics/LookUpTables.jl`).
"""
function B200Engine(s::RRTMGPSolver, lut_pack::Vector{UInt8}; col_offset = 0)
    FT = eltype(s.grid_params)
    as = s.as
    rm = s.radiation_method
    vmr = as.vmr
    cfg = Config(1, CUDA.deviceid(), FT === Float64 ? 1 : 0, s.grid_params.ncol, s.grid_params.nlay,
                 vmr isa RRTMGP.VolumeMixingRatios.VmrGM ? length(vmr.vmr) : size(vmr.vmr, 1),
                 vmr isa RRTMGP.VolumeMixingRatios.VmrGM ? 0 : 1, method_code(rm), rm.aerosol_radiation ? 1 : 0,
                 s.lws.op isa RRTMGP.Optics.OneScalar ? 1 : 0,
                 s.lws isa RRTMGP.RTE.NoScatLWRTE ? s.lws.angle_disc.n_gauss_angles : 1,
                 isnothing(as.cloud_state) ? 2 : as.cloud_state.ice_rgh,
                 (hasproperty(s.lws, :band_flux) && !isnothing(s.lws.band_flux)) ? 1 : 0, s.grid_params.isothermal_boundary_layer ? 1 : 0,
                 col_offset, s.params.grav, s.params.molmass_dryair, s.params.molmass_water, s.params.avogad)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rrtmgp_b200_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h))
    check(ccall((:rrtmgp_b200_load_luts, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t), h[], lut_pack, length(lut_pack)))
    e = B200Engine(h[])
    finalizer(x -> ccall((:rrtmgp_b200_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), e)
    bind!(e, s)
    # interpolation / bottom_extrapolation / center_z / face_z of the solver (src/api/solver.jl:136-147,183-193)
    check(ccall((:rrtmgp_b200_set_level_interpolation, LIB), Cint,
                (Ptr{Cvoid}, Int32, Int32, CuPtr{Cvoid}, CuPtr{Cvoid}, Float64, Float64), e.handle,
                interpolation_code(s.interpolation), bottom_code(s.bottom_extrapolation), devptr(s.center_z), devptr(s.face_z),
                RRTMGP.Parameters.cp_d(s.params), RRTMGP.Parameters.R_d(s.params)))
    return e
end

interpolation_code(::RRTMGP.NoInterpolation) = 0
interpolation_code(::RRTMGP.ArithmeticMean) = 1
interpolation_code(::RRTMGP.GeometricMean) = 2
interpolation_code(::RRTMGP.UniformZ) = 3
interpolation_code(::RRTMGP.UniformP) = 4
interpolation_code(::RRTMGP.BestFit) = 5
bottom_code(::RRTMGP.SameAsInterpolation) = 0
bottom_code(::RRTMGP.UseSurfaceTempAtBottom) = 1
bottom_code(::RRTMGP.HydrostaticBottom) = 2

function bind!(e::B200Engine, s::RRTMGPSolver)
    as = s.as; cs = as.cloud_state; ae = as.aerosol_state
    gm = as.vmr isa RRTMGP.VolumeMixingRatios.VmrGM
    pf, cf = s.presented_flux_lw, s.presented_flux_sw
    clw, csw = s.clear_flux_lw, s.clear_flux_sw                  # `nothing` unless WithClearSkyDiagnostics
    bl = hasproperty(s.lws, :band_flux) ? s.lws.band_flux : nothing
    bs = s.sws.band_flux
    f(x, name) = isnothing(x) ? nothing : getfield(x, name)
    ptrs = (
        devptr(as.layerdata), devptr(as.p_lev), devptr(as.t_lev), devptr(as.t_sfc),
        devptr(gm ? as.vmr.vmr_h2o : nothing), devptr(gm ? as.vmr.vmr_o3 : nothing), devptr(as.vmr.vmr), devptr(as.lat),
        devptr(f(cs, :cld_r_eff_liq)), devptr(f(cs, :cld_r_eff_ice)), devptr(f(cs, :cld_path_liq)), devptr(f(cs, :cld_path_ice)),
        devptr(f(cs, :cld_frac)), devptr(f(cs, :cld_cover_lw)), devptr(f(cs, :cld_cover_sw)),
        devptr(f(ae, :aero_mass)), devptr(f(ae, :aero_size)), devptr(f(ae, :aod_sw_ext)), devptr(f(ae, :aod_sw_sca)),
        devptr(s.lws.bcs.sfc_emis), devptr(s.lws.bcs.inc_flux), devptr(s.sws.bcs.cos_zenith), devptr(s.sws.bcs.toa_flux),
        devptr(s.sws.bcs.sfc_alb_direct), devptr(s.sws.bcs.sfc_alb_diffuse), devptr(s.deep_atmosphere_inverse_scaling),
        devptr(pf.flux_up), devptr(pf.flux_dn), devptr(pf.flux_net),
        devptr(cf.flux_up), devptr(cf.flux_dn), devptr(cf.flux_net), devptr(cf.flux_dn_dir), devptr(s.net_flux_buffer),
        devptr(f(clw, :flux_up)), devptr(f(clw, :flux_dn)), devptr(f(clw, :flux_net)),
        devptr(f(csw, :flux_up)), devptr(f(csw, :flux_dn)), devptr(f(csw, :flux_net)), devptr(f(csw, :flux_dn_dir)),
        devptr(s.clear_net_flux_buffer),
        devptr(f(bl, :flux_up)), devptr(f(bl, :flux_dn)), devptr(f(bl, :flux_net)),
        devptr(f(bs, :flux_up)), devptr(f(bs, :flux_dn)), devptr(f(bs, :flux_net)),
    )
    check(ccall((:rrtmgp_b200_bind, LIB), Cint, (Ptr{Cvoid}, Ref{Buffers}), e.handle, Buffers(ptrs)))
end

# update_fluxes!(s, seedval) (src/api/update_fluxes.jl:223-233): async on the task-local CUDA stream,
# no allocation, no host sync, returns nothing.
function update_fluxes!(e::B200Engine, seedval = nothing)
    st = CUDA.stream().handle
    check(ccall((:rrtmgp_b200_update_fluxes, LIB), Cint, (Ptr{Cvoid}, UInt64, Cint, Ptr{Cvoid}),
                e.handle, isnothing(seedval) ? UInt64(0) : UInt64(seedval), isnothing(seedval) ? 0 : 1, st))
    return nothing
end
# update_lw_fluxes!(s) / update_sw_fluxes!(s) / update_net_fluxes!(s) (src/api/update_fluxes.jl:12-16, 74-78, 165-194): the
# per-band-type steps of update_fluxes! for a host that calls them one by one (after prepare_atmosphere!)
seed_args(seedval) = (isnothing(seedval) ? UInt64(0) : UInt64(seedval), Cint(isnothing(seedval) ? 0 : 1))
update_lw_fluxes!(e::B200Engine, seedval = nothing) =
    (check(ccall((:rrtmgp_b200_update_lw_fluxes, LIB), Cint, (Ptr{Cvoid}, UInt64, Cint, Ptr{Cvoid}), e.handle, seed_args(seedval)..., CUDA.stream().handle), e.handle); nothing)
update_sw_fluxes!(e::B200Engine, seedval = nothing) =
    (check(ccall((:rrtmgp_b200_update_sw_fluxes, LIB), Cint, (Ptr{Cvoid}, UInt64, Cint, Ptr{Cvoid}), e.handle, seed_args(seedval)..., CUDA.stream().handle), e.handle); nothing)
update_net_fluxes!(e::B200Engine) =
    (check(ccall((:rrtmgp_b200_update_net_fluxes, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.handle, CUDA.stream().handle), e.handle); nothing)
# a column range of the bound arrays on a caller-chosen stream (host pipelines that overlap copies with compute)
update_fluxes_range!(e::B200Engine, seedval, col0::Integer, count::Integer, stream = CUDA.stream()) =
    (check(ccall((:rrtmgp_b200_update_fluxes_range, LIB), Cint, (Ptr{Cvoid}, UInt64, Cint, Int64, Int32, Ptr{Cvoid}), e.handle,
                 seed_args(seedval)..., col0, count, stream.handle), e.handle); nothing)
lut_info(e::B200Engine) = (r = Ref{LutInfo}(); check(ccall((:rrtmgp_b200_lut_info, LIB), Cint, (Ptr{Cvoid}, Ref{LutInfo}), e.handle, r)); r[])
# kernels launched by the last update call (3 for an all-sky update_fluxes!: prepare, LW, SW)
last_launch_count(e::B200Engine) = ccall((:rrtmgp_b200_last_launch_count, LIB), Cint, (Ptr{Cvoid},), e.handle)

prepare_atmosphere!(e::B200Engine) =
    (check(ccall((:rrtmgp_b200_prepare_atmosphere, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.handle, CUDA.stream().handle)); nothing)

# the steps of prepare_atmosphere! one by one (src/api/grid_adaptation.jl:87-113, 135-195, 232-258, 278-293)
prepare_steps!(e::B200Engine, steps) =
    (check(ccall((:rrtmgp_b200_prepare_steps, LIB), Cint, (Ptr{Cvoid}, UInt32, Ptr{Cvoid}), e.handle, UInt32(steps), CUDA.stream().handle)); nothing)
interpolate_levels!(e::B200Engine) = prepare_steps!(e, 1)
add_isothermal_boundary_layer!(e::B200Engine) = prepare_steps!(e, 2)
clip!(e::B200Engine) = prepare_steps!(e, 4)
update_concentrations!(e::B200Engine) = prepare_steps!(e, 8)

# validate_inputs(s) (src/api/validation.jl:56-74); synchronises the stream (it returns a host value)
const INVALID_INPUT_NAMES = (:level_pressure, :level_temperature, :layer_pressure, :layer_temperature, :surface_temperature,
                             :cos_zenith, :toa_sw_flux_dn, :surface_emissivity, :direct_sw_surface_albedo,
                             :diffuse_sw_surface_albedo, :vmr_h2o, :vmr_o3, :vmr)
function validate_inputs(e::B200Engine)
    failed = Ref{UInt32}(0)
    check(ccall((:rrtmgp_b200_validate_inputs, LIB), Cint, (Ptr{Cvoid}, Ref{UInt32}, Ptr{Cvoid}), e.handle, failed, CUDA.stream().handle))
    for (bit, name) in enumerate(INVALID_INPUT_NAMES)
        (failed[] >> (bit - 1)) & 1 == 1 && error("RRTMGP input validation failed: `$name` contains values outside its physical range (or non-finite values).")
    end
    return nothing
end

# heating_rate(s) (src/api/standalone.jl:106-124): allocates and returns a fresh (nlay, ncol) array
function heating_rate(e::B200Engine, s::RRTMGPSolver)
    nlay = s.grid_params.nlay - Int(s.grid_params.isothermal_boundary_layer)
    hr = similar(RRTMGP.net_flux(s), nlay, s.grid_params.ncol)
    check(ccall((:rrtmgp_b200_heating_rate, LIB), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Float64, Ptr{Cvoid}), e.handle,
                devptr(s.net_flux_buffer), devptr(hr), RRTMGP.Parameters.cp_d(s.params), CUDA.stream().handle))
    return hr
end

# ---- multi-GPU: column shards, one process (MPI rank) per GPU (include/rrtmgp_b200.h "multi-GPU") ----
# The reference has no multi-GPU path (docs/src/howto/gpu.md:69-84).  Rank 0 creates the NCCL id, the host ships its
# 128 bytes to every rank (e.g. `MPI.Bcast!`), every rank calls `comm_init!`; afterwards `update_fluxes_gathered!`
# leaves the (nlev, nranks * ncol) concatenation of the eight flux views on EVERY rank, the transfers overlapped
# with the shortwave kernel (copy engines over NVLink).  The gathered arrays are owned by the library and wrapped,
# not copied.
comm_unique_id() = (id = Vector{UInt8}(undef, 128); check(ccall((:rrtmgp_b200_comm_unique_id, LIB), Cint, (Ptr{UInt8}, Csize_t), id, 128)); id)

const GATHERED_NAMES = (:lw_flux_up, :lw_flux_dn, :lw_flux_net, :sw_flux_up, :sw_flux_dn, :sw_flux_net, :sw_flux_dn_dir, :net_flux)

function comm_init!(e::B200Engine, s::RRTMGPSolver, unique_id::Vector{UInt8}, rank::Integer, nranks::Integer)
    length(unique_id) == 128 || error("unique_id must be 128 bytes")
    check(ccall((:rrtmgp_b200_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), e.handle, unique_id, rank, nranks))
    ptrs = Ref(ntuple(_ -> CuPtr{Cvoid}(0), 8))
    check(ccall((:rrtmgp_b200_gathered_buffers, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.handle, ptrs))
    FT = eltype(s.grid_params)
    dims = (s.grid_params.nlay + 1, nranks * s.grid_params.ncol)            # (nlev, nranks * ncol), column-major
    return NamedTuple{GATHERED_NAMES}(ntuple(i -> unsafe_wrap(CuArray, reinterpret(CuPtr{FT}, ptrs[][i]), dims), 8))
end

function update_fluxes_gathered!(e::B200Engine, seedval = nothing)
    check(ccall((:rrtmgp_b200_update_fluxes_gathered, LIB), Cint, (Ptr{Cvoid}, UInt64, Cint, Ptr{Cvoid}), e.handle,
                isnothing(seedval) ? UInt64(0) : UInt64(seedval), isnothing(seedval) ? 0 : 1, CUDA.stream().handle))
    return nothing
end
all_gather_fluxes!(e::B200Engine) =
    (check(ccall((:rrtmgp_b200_all_gather_fluxes, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.handle, CUDA.stream().handle)); nothing)
comm_destroy!(e::B200Engine) = (check(ccall((:rrtmgp_b200_comm_destroy, LIB), Cint, (Ptr{Cvoid},), e.handle)); nothing)

# relative humidity as the engine's kernel (a host duty in the reference, grid_adaptation.jl:267-270)
compute_relative_humidity!(e::B200Engine) =
    (check(ccall((:rrtmgp_b200_compute_relative_humidity, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), e.handle, CUDA.stream().handle)); nothing)

# name -> Array pairs of a loaded `LookupBundle` (src/api/lookup_bundle.jl:29-46), i.e. exactly what
# `lookup_tables(grid_params, method)` built from the rrtmgp-data artifact, in the entry names
# `rrtmgp_b200_load_luts` reads.  `rrtmgp.jl_b200/tables.py` produces the same entries from the NetCDF files
# without Julia.  Cloud / aerosol sections are written only when the bundle holds them.
function lut_arrays(b)
    h = x -> Array(Adapt.adapt(Array, x))
    out = Pair{String, Array}[]
    gas(pre, l, sw) = begin
        push!(out, "$pre/key_species" => h(l.key_species), "$pre/kmajor" => h(l.kmajor),
              "$pre/bnd_lims_gpt" => h(l.band_data.bnd_lims_gpt), "$pre/major_gpt2bnd" => h(l.band_data.major_gpt2bnd),
              "$pre/bnd_lims_wn" => h(l.band_data.bnd_lims_wn),
              "$pre/ln_p_ref" => h(l.ref_points.ln_p_ref),          # the struct keeps only log(p_ref)
              "$pre/t_ref" => h(l.ref_points.t_ref), "$pre/vmr_ref" => h(l.ref_points.vmr_ref),
              "$pre/idx_h2o" => Int32[l.idx_h2o],
              "$pre/params" => Float64[l.p_ref_tropo, l.p_ref_min, l.t_ref_min, l.t_ref_max, sw ? l.solar_src_tot : 0])
        for (tag, m) in (("lower", l.minor_lower), ("upper", l.minor_upper))
            push!(out, "$pre/minor_$tag/bnd_st" => h(m.bnd_st), "$pre/minor_$tag/gpt_st" => h(m.gpt_st),
                  "$pre/minor_$tag/gasdata" => h(m.gasdata), "$pre/minor_$tag/kminor" => h(m.kminor))
        end
        if sw
            push!(out, "$pre/rayl_lower" => h(l.rayl_lower), "$pre/rayl_upper" => h(l.rayl_upper),
                  "$pre/solar_src_scaled" => h(l.solar_src_scaled))
        else
            push!(out, "$pre/planck_fraction" => h(l.planck.planck_fraction), "$pre/t_planck" => h(l.planck.t_planck),
                  "$pre/tot_planck" => h(l.planck.tot_planck))
        end
    end
    gas("lw", b.lookup_lw, false)
    gas("sw", b.lookup_sw, true)
    for (tag, c) in (("cld_lw", b.lookup_lw_cld), ("cld_sw", b.lookup_sw_cld))
        isnothing(c) && continue
        push!(out, "$tag/dims" => Int32.(h(c.dims)), "$tag/bounds" => h(c.bounds), "$tag/liqdata" => h(c.liqdata),
              "$tag/icedata" => h(c.icedata), "$tag/bnd_lims_wn" => h(c.bnd_lims_wn))
    end
    for (tag, a) in (("aero_lw", b.lookup_lw_aero), ("aero_sw", b.lookup_sw_aero))
        isnothing(a) && continue
        push!(out, "$tag/dims" => Int32.(h(a.dims)), "$tag/size_bin_limits" => h(a.size_bin_limits),
              "$tag/rh_levels" => h(a.rh_levels), "$tag/dust" => h(a.dust), "$tag/sea_salt" => h(a.sea_salt),
              "$tag/sulfate" => h(a.sulfate), "$tag/black_carbon_rh" => h(a.black_carbon_rh),
              "$tag/black_carbon" => h(a.black_carbon), "$tag/organic_carbon_rh" => h(a.organic_carbon_rh),
              "$tag/organic_carbon" => h(a.organic_carbon), "$tag/bnd_lims_wn" => h(a.bnd_lims_wn),
              "$tag/iband_550nm" => Int32[a.iband_550nm])
    end
    return out
end
lut_pack(s::RRTMGPSolver) = write_lut_pack(lut_arrays(s.lookups))

# LUT pack writer: name -> Array in the post-load layouts of src/optics/LookUpTables.jl
# (format: rrtmgp.jl_b200/lutpack.py; names: rrtmgp.jl_b200/synthetic.py `make_lut_arrays`).
function write_lut_pack(arrays::Vector{Pair{String, Array}})
    io = IOBuffer()
    n = length(arrays)
    off = cld(32 + 80n, 64) * 64
    entries = IOBuffer(); blobs = IOBuffer()
    for (name, a) in arrays
        isint = eltype(a) <: Integer
        data = isint ? Int32.(a) : Float64.(a)
        write(entries, rpad(name, 32, '\0')); write(entries, UInt32(isint ? 1 : 0), UInt32(ndims(a)))
        write(entries, UInt32.(vcat(collect(size(a)), ones(Int, 6 - ndims(a)))))
        write(entries, UInt64(off), UInt64(sizeof(data)))
        write(blobs, data); pad = mod(-sizeof(data), 64); write(blobs, zeros(UInt8, pad))
        off += sizeof(data) + pad
    end
    write(io, "RRTMGPB200LUT\0\0\0"); write(io, UInt32(1), UInt32(n), UInt64(off))
    write(io, take!(entries)); write(io, zeros(UInt8, mod(-position(io), 64))); write(io, take!(blobs))
    return take!(io)
end

end # module

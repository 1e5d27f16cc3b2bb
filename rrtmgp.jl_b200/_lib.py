"""ctypes binding of the C ABI in include/rrtmgp_b200.h (the same symbols a Julia host `ccall`s).

There is deliberately NO fallback: if `librrtmgp_b200.so` is missing or a call fails, this
module raises -- the product path never routes through the CPU oracle or torch ops.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(_CSRC, "librrtmgp_b200.so")

ABI_VERSION = 1
OK, ERR_INVALID_ARG, ERR_BAD_LUT_PACK, ERR_NOT_READY, ERR_CUDA, ERR_UNSUPPORTED = range(6)
CLEAR_SKY, ALL_SKY, ALL_SKY_WITH_CLEAR = 0, 1, 2
TWO_STREAM, ONE_SCALAR = 0, 1
VMR_GM, VMR_FULL = 0, 1
STEP_INTERPOLATE_LEVELS, STEP_BOUNDARY_LAYER, STEP_CLIP, STEP_CONCENTRATIONS = 1, 2, 4, 8
# bits of rrtmgp_b200_validate_inputs, named by the getter to inspect (src/api/validation.jl:56-74)
INVALID_INPUT_NAMES = ("level_pressure", "level_temperature", "layer_pressure", "layer_temperature", "surface_temperature",
                       "cos_zenith", "toa_sw_flux_dn", "surface_emissivity", "direct_sw_surface_albedo",
                       "diffuse_sw_surface_albedo", "vmr_h2o", "vmr_o3", "vmr")


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "device", "dtype", "ncol", "nlay", "ngas", "vmr_kind", "method", "aerosol_radiation",
        "op_lw", "n_gauss_angles", "ice_rgh", "spectral_fluxes", "isothermal_boundary_layer")] + [
        ("col_offset", C.c_int64), ("grav", C.c_double), ("molmass_dryair", C.c_double),
        ("molmass_water", C.c_double), ("avogad", C.c_double)]


BUFFER_FIELDS = (
    "layerdata", "p_lev", "t_lev", "t_sfc", "vmr_h2o", "vmr_o3", "vmr", "lat",
    "cld_r_eff_liq", "cld_r_eff_ice", "cld_path_liq", "cld_path_ice", "cld_frac", "cld_cover_lw", "cld_cover_sw",
    "aero_mass", "aero_size", "aod_sw_ext", "aod_sw_sca",
    "sfc_emis", "inc_flux_lw", "cos_zenith", "toa_flux", "sfc_alb_direct", "sfc_alb_diffuse", "metric_scaling",
    "lw_flux_up", "lw_flux_dn", "lw_flux_net", "sw_flux_up", "sw_flux_dn", "sw_flux_net", "sw_flux_dn_dir", "net_flux",
    "clear_lw_flux_up", "clear_lw_flux_dn", "clear_lw_flux_net", "clear_sw_flux_up", "clear_sw_flux_dn",
    "clear_sw_flux_net", "clear_sw_flux_dn_dir", "clear_net_flux",
    "lw_band_flux_up", "lw_band_flux_dn", "lw_band_flux_net", "sw_band_flux_up", "sw_band_flux_dn", "sw_band_flux_net")


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BUFFER_FIELDS]


GATHERED_FIELDS = ("lw_flux_up", "lw_flux_dn", "lw_flux_net", "sw_flux_up", "sw_flux_dn", "sw_flux_net", "sw_flux_dn_dir", "net_flux")
UNIQUE_ID_BYTES = 128


class Gathered(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in GATHERED_FIELDS]


class LutInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_gpt_lw", "n_bnd_lw", "n_gpt_sw", "n_bnd_sw", "ngas", "iband_550nm")] + [
        (n, C.c_double) for n in ("p_ref_min", "t_ref_min", "t_ref_max", "solar_src_tot")]


# AbstractInterpolation / AbstractBottomExtrapolation (src/api/interpolation.jl:40-135)
INTERPOLATIONS = {"NoInterpolation": 0, "ArithmeticMean": 1, "GeometricMean": 2, "UniformZ": 3, "UniformP": 4, "BestFit": 5}
BOTTOM_EXTRAPOLATIONS = {"SameAsInterpolation": 0, "UseSurfaceTempAtBottom": 1, "HydrostaticBottom": 2}

# every symbol include/rrtmgp_b200.h declares
EXPORTS = ("rrtmgp_b200_create", "rrtmgp_b200_destroy", "rrtmgp_b200_load_luts", "rrtmgp_b200_lut_info",
           "rrtmgp_b200_bind", "rrtmgp_b200_prepare_atmosphere", "rrtmgp_b200_prepare_steps", "rrtmgp_b200_update_lw_fluxes",
           "rrtmgp_b200_update_sw_fluxes", "rrtmgp_b200_update_net_fluxes", "rrtmgp_b200_update_fluxes",
           "rrtmgp_b200_update_fluxes_range",
           "rrtmgp_b200_set_level_interpolation", "rrtmgp_b200_heating_rate",
           "rrtmgp_b200_compute_relative_humidity", "rrtmgp_b200_validate_inputs", "rrtmgp_b200_measure_fp32_peak",
           "rrtmgp_b200_comm_unique_id", "rrtmgp_b200_comm_init", "rrtmgp_b200_gathered_buffers",
           "rrtmgp_b200_update_fluxes_gathered", "rrtmgp_b200_all_gather_fluxes", "rrtmgp_b200_comm_destroy",
           "rrtmgp_b200_last_launch_count",
           "rrtmgp_b200_last_cuda_error", "rrtmgp_b200_strerror", "rrtmgp_b200_abi_version")


def build_ext(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-j", str(max(1, min(8, os.cpu_count() or 1))), "-C", _CSRC] + (["-B"] if force else [])
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        H = C.c_void_p
        L.rrtmgp_b200_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
        L.rrtmgp_b200_destroy.argtypes = [H]
        L.rrtmgp_b200_destroy.restype = None
        L.rrtmgp_b200_load_luts.argtypes = [H, C.c_char_p, C.c_size_t]
        L.rrtmgp_b200_lut_info.argtypes = [H, C.POINTER(LutInfo)]
        L.rrtmgp_b200_bind.argtypes = [H, C.POINTER(Buffers)]
        L.rrtmgp_b200_prepare_atmosphere.argtypes = [H, C.c_void_p]
        L.rrtmgp_b200_prepare_steps.argtypes = [H, C.c_uint32, C.c_void_p]
        L.rrtmgp_b200_validate_inputs.argtypes = [H, C.POINTER(C.c_uint32), C.c_void_p]
        for n in ("rrtmgp_b200_update_lw_fluxes", "rrtmgp_b200_update_sw_fluxes", "rrtmgp_b200_update_fluxes"):
            getattr(L, n).argtypes = [H, C.c_uint64, C.c_int, C.c_void_p]
        L.rrtmgp_b200_update_net_fluxes.argtypes = [H, C.c_void_p]
        L.rrtmgp_b200_update_fluxes_range.argtypes = [H, C.c_uint64, C.c_int, C.c_int64, C.c_int32, C.c_void_p]
        L.rrtmgp_b200_compute_relative_humidity.argtypes = [H, C.c_void_p]
        L.rrtmgp_b200_set_level_interpolation.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.rrtmgp_b200_heating_rate.argtypes = [H, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.rrtmgp_b200_comm_unique_id.argtypes = [C.c_void_p, C.c_size_t]
        L.rrtmgp_b200_comm_init.argtypes = [H, C.c_void_p, C.c_int32, C.c_int32]
        L.rrtmgp_b200_gathered_buffers.argtypes = [H, C.POINTER(Gathered)]
        L.rrtmgp_b200_update_fluxes_gathered.argtypes = [H, C.c_uint64, C.c_int, C.c_void_p]
        L.rrtmgp_b200_all_gather_fluxes.argtypes = [H, C.c_void_p]
        L.rrtmgp_b200_comm_destroy.argtypes = [H]
        L.rrtmgp_b200_measure_fp32_peak.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.rrtmgp_b200_last_launch_count.argtypes = [H]
        L.rrtmgp_b200_last_cuda_error.argtypes = [H]
        L.rrtmgp_b200_last_cuda_error.restype = C.c_char_p
        L.rrtmgp_b200_strerror.argtypes = [C.c_int]
        L.rrtmgp_b200_strerror.restype = C.c_char_p
        L.rrtmgp_b200_abi_version.restype = C.c_int
        if L.rrtmgp_b200_abi_version() != ABI_VERSION:
            raise RuntimeError("librrtmgp_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


class RRTMGPB200Error(RuntimeError):
    pass


def check(status: int, handle=None):
    if status != OK:
        msg = lib().rrtmgp_b200_strerror(status).decode()
        if status == ERR_CUDA and handle:
            msg += ": " + lib().rrtmgp_b200_last_cuda_error(handle).decode()
        raise RRTMGPB200Error(f"rrtmgp_b200 status {status}: {msg}")

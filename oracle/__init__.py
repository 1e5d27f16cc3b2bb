"""ctypes front-end of the CPU oracle (oracle/rrtmgp_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, `__graft_entry__.smoke()` and bench.py's
`cpu_baseline` / `--impl reference` legs.  The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    if os.environ.get("RRTMGP_ORACLE_LIB"):      # tools/opcount.py: the operation-counting build of the same source
        return os.environ["RRTMGP_ORACLE_LIB"]
    src = os.path.join(_HERE, "rrtmgp_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _State(C.Structure):
    _fields_ = [("ncol", C.c_int), ("nlay", C.c_int), ("ngas", C.c_int), ("vmr_kind", C.c_int),
                ("ice_rgh", C.c_int), ("pad_", C.c_int), ("col_offset", C.c_longlong)] + [
        (n, C.c_void_p) for n in (
            "layerdata", "p_lev", "t_lev", "t_sfc", "vmr_h2o", "vmr_o3", "vmr", "lat",
            "cld_r_eff_liq", "cld_r_eff_ice", "cld_path_liq", "cld_path_ice", "cld_frac",
            "aero_mass", "aero_size", "sfc_emis", "inc_flux_lw", "cos_zenith", "toa_flux",
            "sfc_alb_direct", "sfc_alb_diffuse", "metric_scaling")]


_OUT_FIELDS = ("lw_up", "lw_dn", "lw_net", "sw_up", "sw_dn", "sw_net", "sw_dir", "net",
               "clear_lw_up", "clear_lw_dn", "clear_lw_net", "clear_sw_up", "clear_sw_dn",
               "clear_sw_net", "clear_sw_dir", "clear_net",
               "cld_cover_lw", "cld_cover_sw", "aod_sw_ext", "aod_sw_sca",
               "lw_band_up", "lw_band_dn", "sw_band_up", "sw_band_dn", "mask_lw", "mask_sw")


class _Out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _OUT_FIELDS]


class _Opts(C.Structure):
    _fields_ = [("method", C.c_int), ("aerosols", C.c_int), ("lw_noscat", C.c_int),
                ("n_gauss_angles", C.c_int), ("do_prepare", C.c_int), ("do_lw", C.c_int),
                ("do_sw", C.c_int), ("nthreads", C.c_int), ("seed", C.c_ulonglong),
                ("grav", C.c_double), ("molmass_dryair", C.c_double), ("molmass_water", C.c_double),
                ("avogad", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_create.restype = C.c_void_p
        _lib.oracle_create.argtypes = [C.c_char_p, C.c_size_t, C.c_int]
        _lib.oracle_destroy.argtypes = [C.c_void_p, C.c_int]
        _lib.oracle_update_fluxes.argtypes = [C.c_void_p, C.c_int, C.POINTER(_State), C.POINTER(_Out), C.POINTER(_Opts)]
        _lib.oracle_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        dp = np.ctypeslib.ndpointer(np.float64, flags="C")
        _lib.oracle_loc_lower_eq.argtypes = [C.c_double, C.c_double, C.c_int, dp]
        _lib.oracle_loc_lower.argtypes = [C.c_double, dp, C.c_int]
        _lib.oracle_interp1d_equispaced.restype = C.c_double
        _lib.oracle_interp1d_equispaced.argtypes = [C.c_double, dp, dp, C.c_int]
        _lib.oracle_interp1d_loc_factor.argtypes = [C.c_double, dp, C.c_int, C.POINTER(C.c_double)]
        _lib.oracle_gauss_angles.argtypes = [C.c_int, dp, dp]
        _lib.oracle_lw_noscat_one_angle.argtypes = [C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_int,
                                                    C.c_double, C.c_double, C.c_double, dp, dp]
        _lib.oracle_lw_2stream_coeffs.argtypes = [C.c_int] + [C.c_double] * 5 + [dp]
        _lib.oracle_sw_2stream_coeffs.argtypes = [C.c_int] + [C.c_double] * 4 + [dp]
        _lib.oracle_build_cloud_mask.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_ulonglong, C.c_longlong,
                                                 C.c_int, C.c_int, C.c_void_p]
        _lib.oracle_mcica_rand.restype = C.c_double
        _lib.oracle_mcica_rand.argtypes = [C.c_ulonglong, C.c_longlong, C.c_int, C.c_int, C.c_int]
        vp = C.c_void_p
        _lib.oracle_gray_setup.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_double, C.c_double, C.c_double,
                                           C.c_double, vp, vp, vp, vp, vp, vp]
        _lib.oracle_gray_solve_lw.argtypes = [C.c_int] * 5 + [dp, C.c_double] + [vp] * 12
        _lib.oracle_gray_solve_sw.argtypes = [C.c_int] * 5 + [dp] + [vp] * 12
        _lib.oracle_gray_heating_rate.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, C.c_double, C.c_double, vp]
        _lib.oracle_gray_update_profile.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double] + [vp] * 7
        _lib.oracle_interpolate_levels.argtypes = [C.c_int] * 6 + [vp] * 5 + [C.c_double] * 3 + [vp] * 2
        _lib.oracle_compute_relative_humidity.argtypes = [C.c_int] * 3 + [vp] * 3 + [C.c_double] * 2 + [vp]
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# default physical parameters = the reference tests' overrides
# (test/all_sky_with_aerosols_utils.jl:41-43 on src/api/standalone.jl:87-97)
DEFAULT_PARAMS = dict(grav=9.80665, molmass_dryair=0.028964, molmass_water=0.018016, avogad=6.02214076e23)


class Oracle:
    """Restated reference `update_fluxes!` on the CPU for one LUT pack and one precision."""

    def __init__(self, pack: bytes, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.is_f64 = int(self.dtype == np.float64)
        self._pack = pack  # keep alive
        self._h = lib().oracle_create(pack, len(pack), self.is_f64)
        if not self._h:
            raise RuntimeError("oracle_create failed (bad LUT pack)")
        d = (C.c_int * 5)()
        lib().oracle_dims(self._h, self.is_f64, d)
        self.n_gpt_lw, self.n_bnd_lw, self.n_gpt_sw, self.n_bnd_sw, self.ngas = list(d)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h, self.is_f64)
            self._h = None

    def update_fluxes(self, state: Dict[str, np.ndarray], *, method: str = "all_sky", aerosols: bool = True,
                      lw_noscat: bool = False, n_gauss_angles: int = 1, prepare: bool = True, do_lw: bool = True,
                      do_sw: bool = True, seed: int = 0, ice_rgh: int = 2, col_offset: int = 0, nthreads: int = 0,
                      spectral: bool = False, masks: bool = False, params: Optional[dict] = None,
                      inplace: bool = False) -> Dict[str, np.ndarray]:
        """Runs prepare + LW + SW + net. `state` arrays (see synthetic.make_atmosphere) are converted to
        the oracle precision; unless `inplace`, they are copied so prepare's in-place clipping does
        not leak (the mutated copies come back under 'state')."""
        dt = self.dtype
        st = {k: (np.ascontiguousarray(v, dtype=dt) if not inplace else v) for k, v in state.items()}
        if not inplace:
            st = {k: (v.copy() if v is state.get(k) else v) for k, v in st.items()}
        ncol, nlay = st["layerdata"].shape[:2]
        nlev = nlay + 1
        s = _State()
        s.ncol, s.nlay, s.ice_rgh, s.col_offset = ncol, nlay, ice_rgh, col_offset
        if "vmr_full" in st:
            s.vmr_kind, s.ngas, s.vmr = 1, st["vmr_full"].shape[2], _ptr(st["vmr_full"])
        else:
            s.vmr_kind, s.ngas = 0, st["vmr"].shape[0]
            s.vmr, s.vmr_h2o, s.vmr_o3 = _ptr(st["vmr"]), _ptr(st["vmr_h2o"]), _ptr(st["vmr_o3"])
        for k in ("layerdata", "p_lev", "t_lev", "t_sfc", "lat", "cld_r_eff_liq", "cld_r_eff_ice", "cld_path_liq",
                  "cld_path_ice", "cld_frac", "aero_mass", "aero_size", "sfc_emis", "inc_flux_lw", "cos_zenith",
                  "toa_flux", "sfc_alb_direct", "sfc_alb_diffuse", "metric_scaling"):
            setattr(s, k, _ptr(st.get(k)))
        out: Dict[str, np.ndarray] = {}
        o = _Out()

        def alloc(name, shape, dtype=dt):
            out[name] = np.zeros(shape, dtype=dtype)
            setattr(o, name, _ptr(out[name]))

        for n in ("lw_up", "lw_dn", "lw_net", "sw_up", "sw_dn", "sw_net", "sw_dir", "net"):
            alloc(n, (ncol, nlev))
        m = {"clear_sky": 0, "all_sky": 1, "all_sky_with_clear": 2}[method]
        if m == 2:
            for n in ("clear_lw_up", "clear_lw_dn", "clear_lw_net", "clear_sw_up", "clear_sw_dn", "clear_sw_net",
                      "clear_sw_dir", "clear_net"):
                alloc(n, (ncol, nlev))
        if m >= 1 and "cld_frac" in st:
            alloc("cld_cover_lw", (ncol,))
            alloc("cld_cover_sw", (ncol,))
        if aerosols and "aero_mass" in st:
            alloc("aod_sw_ext", (ncol,))
            alloc("aod_sw_sca", (ncol,))
        if spectral:
            alloc("lw_band_up", (self.n_bnd_lw, ncol, nlev))
            alloc("lw_band_dn", (self.n_bnd_lw, ncol, nlev))
            alloc("sw_band_up", (self.n_bnd_sw, ncol, nlev))
            alloc("sw_band_dn", (self.n_bnd_sw, ncol, nlev))
        if masks and m >= 1 and "cld_frac" in st:
            alloc("mask_lw", (self.n_gpt_lw, ncol, nlay), np.uint8)
            alloc("mask_sw", (self.n_gpt_sw, ncol, nlay), np.uint8)
        p = dict(DEFAULT_PARAMS)
        p.update(params or {})
        op = _Opts(m, int(aerosols), int(lw_noscat), n_gauss_angles, int(prepare), int(do_lw), int(do_sw),
                   nthreads, seed, p["grav"], p["molmass_dryair"], p["molmass_water"], p["avogad"])
        rc = lib().oracle_update_fluxes(self._h, self.is_f64, C.byref(s), C.byref(o), C.byref(op))
        if rc != 0:
            raise RuntimeError(f"oracle_update_fluxes -> {rc}")
        out["state"] = st
        return out


# ---------------------------------------------------------------------------------------
# gray radiation (BASELINE config 1 plumbing) and unit-level wrappers
# ---------------------------------------------------------------------------------------
# src/api/standalone.jl:87-97
GRAY_PARAMS = dict(grav=9.81, molmass_dryair=0.02897, molmass_water=0.018015, gas_constant=8.314462618,
                   kappa_d=2.0 / 7.0, Stefan=5.670374419e-8, avogad=6.02214076e23)
OTP_SCHNEIDER2004 = (0, np.array([3.5, 300.0, 200.0, 60.0, 0.0]))       # gray_atmospheric_states.jl:36-43
OTP_OGORMAN2008 = (1, np.array([1.0, 0.2, 7.2, 1.8, 0.22]))             # gray_atmospheric_states.jl:75-84


def gray_setup(dtype, lat, nlay, p0=1.0e5, pe=9.0e3, params=GRAY_PARAMS):
    """`setup_gray_as_pr_grid` (src/optics/gray_atmospheric_states.jl:154-299)."""
    dt = np.dtype(dtype)
    lat = np.ascontiguousarray(lat, dtype=dt)
    ncol = lat.size
    z = lambda *s: np.zeros(s, dtype=dt)
    st = dict(lat=lat, p_lev=z(ncol, nlay + 1), p_lay=z(ncol, nlay), t_lev=z(ncol, nlay + 1), t_lay=z(ncol, nlay),
              z_lev=z(ncol, nlay + 1), t_sfc=z(ncol))
    r_d = params["gas_constant"] / params["molmass_dryair"]
    lib().oracle_gray_setup(int(dt == np.float64), ncol, nlay, _ptr(lat), p0, pe, r_d, params["grav"],
                            _ptr(st["p_lev"]), _ptr(st["p_lay"]), _ptr(st["t_lev"]), _ptr(st["t_lay"]),
                            _ptr(st["z_lev"]), _ptr(st["t_sfc"]))
    return st


def gray_solve_lw(st, otp, *, two_stream=True, sfc_emis=1.0, inc_flux=None, scaling=None, params=GRAY_PARAMS):
    dt = st["p_lev"].dtype
    ncol, nlev = st["p_lev"].shape
    emis = np.full(ncol, sfc_emis, dtype=dt)
    inc = None if inc_flux is None else np.ascontiguousarray(np.broadcast_to(inc_flux, (ncol,)), dtype=dt)
    sc = None if scaling is None else np.ascontiguousarray(scaling, dtype=dt)
    up, dn, net = (np.zeros((ncol, nlev), dtype=dt) for _ in range(3))
    lib().oracle_gray_solve_lw(int(dt == np.float64), ncol, nlev - 1, int(two_stream), otp[0],
                               np.ascontiguousarray(otp[1], dtype=np.float64), params["Stefan"], _ptr(st["lat"]),
                               _ptr(st["p_lev"]), _ptr(st["p_lay"]), _ptr(st["t_lev"]), _ptr(st["t_lay"]),
                               _ptr(st["t_sfc"]), _ptr(emis), _ptr(inc), _ptr(sc), _ptr(up), _ptr(dn), _ptr(net))
    return dict(up=up, dn=dn, net=net)


def gray_solve_sw(st, otp, *, two_stream=True, cos_zenith=0.5, toa_flux=1361.0, albedo=0.2):
    dt = st["p_lev"].dtype
    ncol, nlev = st["p_lev"].shape
    full = lambda v: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=dt), (ncol,)))
    mu0, toa, ad, af = full(cos_zenith), full(toa_flux), full(albedo), full(albedo)
    up, dn, net, dr = (np.zeros((ncol, nlev), dtype=dt) for _ in range(4))
    tau = np.zeros((ncol, nlev - 1), dtype=dt)
    lib().oracle_gray_solve_sw(int(dt == np.float64), ncol, nlev - 1, int(two_stream), otp[0],
                               np.ascontiguousarray(otp[1], dtype=np.float64), _ptr(st["lat"]), _ptr(st["p_lev"]),
                               _ptr(st["p_lay"]), _ptr(mu0), _ptr(toa), _ptr(ad), _ptr(af), _ptr(up), _ptr(dn),
                               _ptr(net), _ptr(dr), _ptr(tau))
    return dict(up=up, dn=dn, net=net, dir=dr, tau=tau)


def gray_heating_rate(net, p_lev, params=GRAY_PARAMS):
    """`compute_gray_heating_rate_kernel!` (src/optics/GrayAtmosphere.jl:152-167)."""
    dt = net.dtype
    ncol, nlev = net.shape
    cp_d = params["gas_constant"] / params["molmass_dryair"] / params["kappa_d"]
    hr = np.zeros((ncol, nlev - 1), dtype=dt)
    lib().oracle_gray_heating_rate(int(dt == np.float64), ncol, nlev - 1, _ptr(net), _ptr(p_lev),
                                   float(dt.type(params["grav"])), float(dt.type(cp_d)), _ptr(hr))
    return hr


def gray_update_profile(st, hr, dn, net, dt_seconds, params=GRAY_PARAMS):
    """`update_profile_lw_kernel!` (src/optics/GrayAtmosphere.jl:64-125); mutates t_lay/t_lev."""
    dt = net.dtype
    ncol, nlev = net.shape
    t_ex = np.zeros((ncol, nlev), dtype=dt)
    grad = np.zeros((ncol, nlev - 1), dtype=dt)
    lib().oracle_gray_update_profile(int(dt == np.float64), ncol, nlev - 1, params["Stefan"], dt_seconds, _ptr(hr),
                                     _ptr(dn), _ptr(net), _ptr(st["t_lay"]), _ptr(st["t_lev"]), _ptr(t_ex), _ptr(grad))
    return t_ex, grad


def solve_gray(dtype, nlay=60, ncol=1, *, latitude=None, inc_flux=None, scaling=None, otp=OTP_OGORMAN2008,
               params=GRAY_PARAMS, **sw):
    """`solve_gray` (src/api/standalone.jl:163-210): two-stream LW + SW on the gray pressure grid."""
    lat = latitude if latitude is not None else ([0.0] if ncol == 1 else np.linspace(-90.0, 90.0, ncol))
    st = gray_setup(dtype, lat, nlay, params=params)
    lw = gray_solve_lw(st, otp, two_stream=True, inc_flux=inc_flux, scaling=scaling, params=params)
    s = gray_solve_sw(st, otp, two_stream=True, **sw)
    if scaling is not None:
        sc = np.asarray(scaling, dtype=st["p_lev"].dtype)
        for k in ("up", "dn", "net", "dir"):
            s[k] = s[k] * sc
    net = lw["net"] + s["net"]
    return dict(state=st, lw=lw, sw=s, net=net, heating_rate=gray_heating_rate(net, st["p_lev"], params))


INTERPOLATIONS = {"none": 0, "arithmetic_mean": 1, "geometric_mean": 2, "uniform_z": 3, "uniform_p": 4, "best_fit": 5}
BOTTOM_EXTRAPOLATIONS = {"same_as_interpolation": 0, "use_surface_temp_at_bottom": 1, "hydrostatic_bottom": 2}


def interpolate_levels(p_lay, t_lay, t_sfc, interpolation, bottom_extrapolation="same_as_interpolation", *,
                       center_z=None, face_z=None, nlay=None, params=GRAY_PARAMS, p_lev=None, t_lev=None):
    """`interpolate_levels!` (src/api/grid_adaptation.jl:87-113) with `interp!` / `extrap!`
    (src/api/interpolation.jl:176-252).  Arrays are [ncol][nlay] / [ncol][nlay + 1]; `nlay` restricts the
    domain (isothermal boundary layer on top).  Returns (p_lev, t_lev)."""
    dt = p_lay.dtype
    ncol, stride = p_lay.shape
    nlay = stride if nlay is None else nlay
    cp_d = params["gas_constant"] / params["molmass_dryair"] / params["kappa_d"]
    r_d = params["gas_constant"] / params["molmass_dryair"]
    p_lev = np.zeros((ncol, stride + 1), dtype=dt) if p_lev is None else p_lev
    t_lev = np.zeros((ncol, stride + 1), dtype=dt) if t_lev is None else t_lev
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dt)
    p_lay, t_lay, t_sfc, center_z, face_z = c(p_lay), c(t_lay), c(t_sfc), c(center_z), c(face_z)
    lib().oracle_interpolate_levels(int(dt == np.float64), ncol, nlay, stride, INTERPOLATIONS[interpolation],
                                    BOTTOM_EXTRAPOLATIONS[bottom_extrapolation], _ptr(p_lay), _ptr(t_lay), _ptr(t_sfc),
                                    _ptr(center_z), _ptr(face_z), float(dt.type(params["grav"])), float(dt.type(cp_d)),
                                    float(dt.type(r_d)), _ptr(p_lev), _ptr(t_lev))
    return p_lev, t_lev


def compute_relative_humidity(p_lay, t_lay, vmr_h2o, params=DEFAULT_PARAMS):
    """`compute_relative_humidity!` (src/optics/column_amounts.jl:52-76, kernel gas_optics.jl:58-80); arrays
    [ncol][nlay] in the precision of `p_lay`."""
    dt = p_lay.dtype
    c = lambda a: np.ascontiguousarray(a, dtype=dt)
    p_lay, t_lay, vmr_h2o = c(p_lay), c(t_lay), c(vmr_h2o)
    ncol, nlay = p_lay.shape
    rh = np.zeros((ncol, nlay), dtype=dt)
    lib().oracle_compute_relative_humidity(int(dt == np.float64), ncol, nlay, _ptr(p_lay), _ptr(t_lay), _ptr(vmr_h2o),
                                           params["molmass_water"], params["molmass_dryair"], _ptr(rh))
    return rh

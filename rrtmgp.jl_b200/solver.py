"""Host-side mirror of the reference's Layer-2 interface for the hot path.

Same names and argument meaning as `src/api/{grid_params,radiation_methods,solver,
update_fluxes,getters}.jl`; Julia's `f!(s)` is `f(s)` here.  The solver owns every array
(allocated once at construction, `solver.jl:216-272`), the host writes inputs in place
through the state getters, calls `update_fluxes(s, seed)`, and reads fluxes through the
flux getters.  Arrays are torch CUDA tensors whose memory is exactly the reference's:
Julia `(nlev, ncol)` column-major == torch `[ncol, nlev]` row-major,
so a getter documented as `(nlev, ncol)` in `docs/src/getters.md:12-40` returns a
`[ncol, nlev]` tensor view here.  torch is only the allocator/stream provider; all compute
happens in librrtmgp_b200.so through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import Buffers, Config, LutInfo, check, lib

# src/api/standalone.jl:87-97
DEFAULT_PARAMETERS = dict(grav=9.81, molmass_dryair=0.02897, molmass_water=0.018015, gas_constant=8.314462618,
                          kappa_d=2.0 / 7.0, Stefan=5.670374419e-8, avogad=6.02214076e23)


def default_parameters(**overrides) -> dict:
    p = dict(DEFAULT_PARAMETERS)
    p.update(overrides)
    return p


# src/api/aerosols.jl:18-34
AEROSOL_INDEX = {"dust1": 1, "sea_salt1": 2, "sulfate": 3, "black_carbon_rh": 4, "black_carbon": 5,
                 "organic_carbon_rh": 6, "organic_carbon": 7, "dust2": 8, "dust3": 9, "dust4": 10, "dust5": 11,
                 "sea_salt2": 12, "sea_salt3": 13, "sea_salt4": 14, "sea_salt5": 15}


def canonical_aerosol_name(name: str) -> str:
    """`canonical_aerosol_name(name)` (src/api/aerosols.jl:78-88): the name itself, or an error naming the known ones."""
    if name in AEROSOL_INDEX:
        return str(name)
    raise KeyError(f"unknown aerosol name {name!r}; known names are {sorted(AEROSOL_INDEX)}")


def aerosol_index(name: str) -> int:
    return AEROSOL_INDEX[canonical_aerosol_name(name)]


def aerosol_names():
    return sorted(AEROSOL_INDEX, key=AEROSOL_INDEX.get)


def aerosol_index_map() -> dict:
    """`aerosol_index_map()` (src/api/aerosols.jl): name -> slot of the `aero_mass` / `aero_size` arrays."""
    return dict(AEROSOL_INDEX)


def gas_names_sw():
    """`gas_names_sw()` (src/api/getters.jl:566-588): the gas names `volume_mixing_ratio` accepts."""
    return ["h2o", "cfc11", "h2o_self", "co2", "cfc12", "hfc134a", "cfc22", "ch4", "hfc23", "ccl4", "hfc143a", "co",
            "no2", "n2", "o2", "o3", "h2o_frgn", "hfc32", "n2o", "cf4", "hfc125"]


def requires_z(scheme: str) -> bool:
    """`requires_z(scheme)` (src/api/interpolation.jl:139-146): BestFit and HydrostaticBottom need altitudes."""
    return scheme in ("BestFit", "HydrostaticBottom")


@dataclass(frozen=True)
class RRTMGPGridParams:
    """`RRTMGPGridParams(FT; context, domain_nlay, ncol, isothermal_boundary_layer)` (grid_params.jl:38-54)."""
    FT: type = np.float32
    domain_nlay: int = 64
    ncol: int = 1
    isothermal_boundary_layer: bool = False
    device: int = 0

    @property
    def nlay(self) -> int:
        return self.domain_nlay + int(self.isothermal_boundary_layer)


# radiation_methods.jl:19-70 (GrayRadiation has no g-point axis; it is not on this path)
@dataclass(frozen=True)
class ClearSkyRadiation:
    aerosol_radiation: bool = False


@dataclass(frozen=True)
class AllSkyRadiation:
    aerosol_radiation: bool = False
    reset_rng_seed: bool = False


@dataclass(frozen=True)
class AllSkyRadiationWithClearSkyDiagnostics:
    aerosol_radiation: bool = False
    reset_rng_seed: bool = False


_METHOD_CODE = {ClearSkyRadiation: _lib.CLEAR_SKY, AllSkyRadiation: _lib.ALL_SKY,
                AllSkyRadiationWithClearSkyDiagnostics: _lib.ALL_SKY_WITH_CLEAR}


class RRTMGPSolver:
    """`RRTMGPSolver(grid_params, radiation_method, params, ...)` (solver.jl:95-331) over the C ABI.

    `lookups` is a LUT pack (bytes; `lutpack.pack_luts`).  `op_lw` is "two_stream" or "one_scalar"
    (no-scattering longwave, with `n_gauss_angles` 1..4).  `vmr_kind` "gm" stores H2O/O3 per layer and
    the other gases as global means (`VmrGM`), "full" stores every gas per layer (`Vmr`).
    `col_offset` is the global index of this solver's first column when columns are sharded.
    """

    def __init__(self, grid_params: RRTMGPGridParams, radiation_method, params: dict, lookups: bytes, *,
                 op_lw: str = "two_stream", n_gauss_angles: int = 1, spectral_fluxes: bool = False,
                 deep_atmosphere_inverse_scaling: Optional[torch.Tensor] = None, vmr_kind: str = "gm",
                 ngas: Optional[int] = None, ice_rgh: int = 2, inc_flux_lw: bool = False, with_lat: bool = False,
                 col_offset: int = 0, interpolation: str = "NoInterpolation",
                 bottom_extrapolation: str = "SameAsInterpolation", center_z=None, face_z=None):
        if type(radiation_method) not in _METHOD_CODE:
            raise TypeError("radiation_method must be ClearSkyRadiation, AllSkyRadiation or "
                            "AllSkyRadiationWithClearSkyDiagnostics (GrayRadiation is not on this path)")
        if op_lw not in ("two_stream", "one_scalar"):
            raise ValueError("op_lw must be 'two_stream' or 'one_scalar'")
        # solver.jl:159-171
        if n_gauss_angles != 1 and op_lw != "one_scalar":
            raise ValueError(f"`n_gauss_angles = {n_gauss_angles}` applies only to the non-scattering longwave "
                             "solver; pass op_lw='one_scalar', or keep n_gauss_angles = 1")
        if interpolation not in _lib.INTERPOLATIONS or bottom_extrapolation not in _lib.BOTTOM_EXTRAPOLATIONS:
            raise ValueError(f"interpolation must be one of {sorted(_lib.INTERPOLATIONS)} and bottom_extrapolation "
                             f"one of {sorted(_lib.BOTTOM_EXTRAPOLATIONS)}")
        # solver.jl:183-193: z-based interpolation / extrapolation reads the altitudes at solve time
        if (interpolation == "BestFit" or (interpolation != "NoInterpolation" and bottom_extrapolation == "HydrostaticBottom")) \
                and (center_z is None or face_z is None):
            raise ValueError("BestFit interpolation and HydrostaticBottom extrapolation need layer and level "
                             "altitudes; pass `center_z` and `face_z` to the `RRTMGPSolver`")
        if not torch.cuda.is_available():
            raise RuntimeError("RRTMGPSolver needs a CUDA device (there is no CPU fallback)")
        self.interpolation, self.bottom_extrapolation = interpolation, bottom_extrapolation
        self.grid_params = grid_params
        self.radiation_method = radiation_method
        self.params = dict(params)
        self._lookups = lookups
        self.dtype = np.dtype(grid_params.FT)
        self.tdtype = torch.float64 if self.dtype == np.float64 else torch.float32
        self.device = torch.device("cuda", grid_params.device)
        self.op_lw = op_lw
        self.spectral_fluxes = bool(spectral_fluxes)
        ncol, nlay = grid_params.ncol, grid_params.nlay
        nlev = nlay + 1
        self._h = C.c_void_p()
        L = lib()

        # tables first: their dims size the BC arrays
        probe = Config(_lib.ABI_VERSION, grid_params.device, int(self.dtype == np.float64), ncol, nlay,
                       ngas or 64, _lib.VMR_GM if vmr_kind == "gm" else _lib.VMR_FULL, _METHOD_CODE[type(radiation_method)],
                       int(radiation_method.aerosol_radiation), _lib.TWO_STREAM if op_lw == "two_stream" else _lib.ONE_SCALAR,
                       n_gauss_angles, ice_rgh, int(spectral_fluxes), int(grid_params.isothermal_boundary_layer),
                       col_offset, params["grav"], params["molmass_dryair"], params["molmass_water"], params["avogad"])
        check(L.rrtmgp_b200_create(C.byref(probe), C.byref(self._h)))
        try:
            check(L.rrtmgp_b200_load_luts(self._h, lookups, len(lookups)), self._h)
            info = LutInfo()
            check(L.rrtmgp_b200_lut_info(self._h, C.byref(info)), self._h)
            self.lut_info = info
            if ngas is None or ngas != info.ngas:
                # re-create with the tables' gas count (the vmr gas axis must match for `Vmr`)
                ngas = max(ngas or 0, info.ngas) if vmr_kind == "gm" else info.ngas
                L.rrtmgp_b200_destroy(self._h)
                self._h = C.c_void_p()
                probe.ngas = ngas
                check(L.rrtmgp_b200_create(C.byref(probe), C.byref(self._h)))
                check(L.rrtmgp_b200_load_luts(self._h, lookups, len(lookups)), self._h)
            self.config = probe
            self.ngas = ngas
            z = lambda *shape: torch.zeros(*shape, dtype=self.tdtype, device=self.device)
            B: Dict[str, Optional[torch.Tensor]] = {k: None for k in _lib.BUFFER_FIELDS}
            B["layerdata"], B["p_lev"], B["t_lev"], B["t_sfc"] = z(ncol, nlay, 4), z(ncol, nlev), z(ncol, nlev), z(ncol)
            if vmr_kind == "gm":
                B["vmr_h2o"], B["vmr_o3"], B["vmr"] = z(ncol, nlay), z(ncol, nlay), z(ngas)
            else:
                B["vmr"] = z(ncol, nlay, ngas)
            if with_lat:
                B["lat"] = z(ncol)
            if not isinstance(radiation_method, ClearSkyRadiation):
                for k in ("cld_r_eff_liq", "cld_r_eff_ice", "cld_path_liq", "cld_path_ice", "cld_frac"):
                    B[k] = z(ncol, nlay)
                B["cld_cover_lw"], B["cld_cover_sw"] = z(ncol), z(ncol)
            if radiation_method.aerosol_radiation:
                B["aero_mass"], B["aero_size"] = z(ncol, nlay, 15), z(ncol, nlay, 15)
                B["aod_sw_ext"], B["aod_sw_sca"] = z(ncol), z(ncol)
            B["sfc_emis"] = z(ncol, info.n_bnd_lw)
            if inc_flux_lw:
                B["inc_flux_lw"] = z(info.n_gpt_lw, ncol)
            B["cos_zenith"], B["toa_flux"] = z(ncol), z(ncol)
            B["sfc_alb_direct"], B["sfc_alb_diffuse"] = z(ncol, info.n_bnd_sw), z(ncol, info.n_bnd_sw)
            if deep_atmosphere_inverse_scaling is not None:
                sc = torch.as_tensor(deep_atmosphere_inverse_scaling, dtype=self.tdtype, device=self.device)
                B["metric_scaling"] = sc.expand(ncol, nlev).contiguous()
            for k in ("lw_flux_up", "lw_flux_dn", "lw_flux_net", "sw_flux_up", "sw_flux_dn", "sw_flux_net",
                      "sw_flux_dn_dir", "net_flux"):
                B[k] = z(ncol, nlev)
            if isinstance(radiation_method, AllSkyRadiationWithClearSkyDiagnostics):
                for k in ("clear_lw_flux_up", "clear_lw_flux_dn", "clear_lw_flux_net", "clear_sw_flux_up",
                          "clear_sw_flux_dn", "clear_sw_flux_net", "clear_sw_flux_dn_dir", "clear_net_flux"):
                    B[k] = z(ncol, nlev)
            if spectral_fluxes:
                for k in ("lw_band_flux_up", "lw_band_flux_dn", "lw_band_flux_net"):
                    B[k] = z(info.n_bnd_lw, ncol, nlev)
                for k in ("sw_band_flux_up", "sw_band_flux_dn", "sw_band_flux_net"):
                    B[k] = z(info.n_bnd_sw, ncol, nlev)
            self.buffers = B
            cb = Buffers()
            for k, t in B.items():
                setattr(cb, k, None if t is None else t.data_ptr())
            check(L.rrtmgp_b200_bind(self._h, C.byref(cb)), self._h)
            # level interpolation (solver.jl:136-147; Parameters.cp_d / R_d as in src/Parameters.jl)
            self.center_z = self.face_z = None
            if center_z is not None and face_z is not None:
                self.center_z = torch.zeros(ncol, nlay, dtype=self.tdtype, device=self.device)
                self.face_z = torch.zeros(ncol, nlev, dtype=self.tdtype, device=self.device)
                zc = torch.as_tensor(np.asarray(center_z), dtype=self.tdtype, device=self.device)
                zf = torch.as_tensor(np.asarray(face_z), dtype=self.tdtype, device=self.device)
                self.center_z[:, : zc.shape[1]].copy_(zc)
                self.face_z[:, : zf.shape[1]].copy_(zf)
            if interpolation != "NoInterpolation":
                r_d = params.get("R_d", params["gas_constant"] / params["molmass_dryair"] if "gas_constant" in params else 0.0)
                cp_d = params.get("cp_d", r_d / params["kappa_d"] if "kappa_d" in params else 0.0)
                check(L.rrtmgp_b200_set_level_interpolation(
                    self._h, _lib.INTERPOLATIONS[interpolation], _lib.BOTTOM_EXTRAPOLATIONS[bottom_extrapolation],
                    None if self.center_z is None else self.center_z.data_ptr(),
                    None if self.face_z is None else self.face_z.data_ptr(), float(cp_d), float(r_d)), self._h)
        except Exception:
            L.rrtmgp_b200_destroy(self._h)
            self._h = None
            raise

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().rrtmgp_b200_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- plumbing ------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _domain(self, t: torch.Tensor, levels: bool) -> torch.Tensor:
        """getters.jl:40-46: the isothermal boundary row is masked by the view, not by the kernel."""
        n = self.grid_params.domain_nlay + (1 if levels else 0)
        return t[:, :n]

    def set_state(self, arrays: Dict[str, "np.ndarray"]) -> None:
        """Convenience: copy host arrays (keys of `synthetic.make_atmosphere`) into the solver's arrays."""
        alias = {"vmr_full": "vmr"}
        for k, v in arrays.items():
            dst = self.buffers.get(alias.get(k, k))
            if dst is None:
                continue
            src = torch.as_tensor(np.ascontiguousarray(v, dtype=self.dtype))
            view = dst
            if dst.dim() >= 2 and src.shape != dst.shape and src.shape[0] == dst.shape[0]:
                view = dst[:, : src.shape[1]]  # domain rows only (boundary layer is filled by prepare)
            view.copy_(src, non_blocking=False)

    @property
    def last_launch_count(self) -> int:
        return lib().rrtmgp_b200_last_launch_count(self._h)


# -- running a solve (update_fluxes.jl) ----------------------------------------------------
def _seed_args(s: RRTMGPSolver, seedval):
    # update_fluxes.jl:150-156: the seed is honoured only when the method asks for reproducible seeding
    use = seedval is not None and getattr(s.radiation_method, "reset_rng_seed", False)
    return (C.c_uint64(int(seedval) & (2 ** 64 - 1)) if use else C.c_uint64(0)), int(use)


def update_fluxes(s: RRTMGPSolver, seedval=None) -> None:
    """`update_fluxes!(s, seedval)` (update_fluxes.jl:223-233): async on the current CUDA stream."""
    if check_values.value:   # update_fluxes.jl:224
        validate_inputs(s)
    seed, have = _seed_args(s, seedval)
    check(lib().rrtmgp_b200_update_fluxes(s._h, seed, have, s._stream()), s._h)


def update_fluxes_range(s: RRTMGPSolver, seedval: int, col_begin: int, col_count: int, stream=None) -> None:
    """`update_fluxes!` restricted to columns [col_begin, col_begin + col_count) on `stream` (a
    `torch.cuda.Stream`; default: the current one).  Columns are independent, so any partition of the
    column range gives the same result as one full call with the same seed."""
    st = C.c_void_p((stream or torch.cuda.current_stream(s.device)).cuda_stream)
    check(lib().rrtmgp_b200_update_fluxes_range(s._h, C.c_uint64(int(seedval) & (2 ** 64 - 1)), 1, col_begin,
                                                col_count, st), s._h)


# -- multi-GPU: column shards, one process per GPU (include/rrtmgp_b200.h "multi-GPU"; sharding.py) ---------------
class _DevicePtr:
    """A device allocation owned by librrtmgp_b200.so, exposed to torch through `__cuda_array_interface__`."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def comm_unique_id() -> bytes:
    """The NCCL id rank 0 creates and ships to every rank (`rrtmgp_b200_comm_unique_id`)."""
    buf = (C.c_char * _lib.UNIQUE_ID_BYTES)()
    check(lib().rrtmgp_b200_comm_unique_id(buf, _lib.UNIQUE_ID_BYTES))
    return bytes(buf)


def comm_init(s: RRTMGPSolver, unique_id: bytes, rank: int, nranks: int) -> Dict[str, torch.Tensor]:
    """Joins the communicator and returns the gathered `(nlev, nranks * ncol)` flux views of this rank
    (`[nranks * ncol, nlev]` tensors over memory the library owns; also kept as `s.gathered`)."""
    if len(unique_id) != _lib.UNIQUE_ID_BYTES:
        raise ValueError(f"unique_id must be {_lib.UNIQUE_ID_BYTES} bytes")
    check(lib().rrtmgp_b200_comm_init(s._h, unique_id, rank, nranks), s._h)
    g = _lib.Gathered()
    check(lib().rrtmgp_b200_gathered_buffers(s._h, C.byref(g)), s._h)
    shape = (nranks * s.grid_params.ncol, s.grid_params.nlay + 1)
    typestr = "<f8" if s.dtype == np.float64 else "<f4"
    s.gathered = {k: torch.as_tensor(_DevicePtr(getattr(g, k), shape, typestr), device=s.device) for k in _lib.GATHERED_FIELDS}
    s.comm_rank, s.comm_nranks = rank, nranks
    return s.gathered


def update_fluxes_gathered(s: RRTMGPSolver, seedval=None) -> None:
    """`update_fluxes!` on this rank's columns + the all-gather of the eight flux views into `s.gathered`, overlapped
    (longwave views travel under the shortwave kernel; copy engines over NVLink)."""
    seed, have = _seed_args(s, seedval)
    check(lib().rrtmgp_b200_update_fluxes_gathered(s._h, seed, have, s._stream()), s._h)


def all_gather_fluxes(s: RRTMGPSolver) -> None:
    """One grouped ncclAllGather of the eight local flux views into `s.gathered` (no compute)."""
    check(lib().rrtmgp_b200_all_gather_fluxes(s._h, s._stream()), s._h)


def comm_destroy(s: RRTMGPSolver) -> None:
    s.gathered = None
    check(lib().rrtmgp_b200_comm_destroy(s._h), s._h)


INPUT_KEYS = ("layerdata", "p_lev", "t_lev", "t_sfc", "vmr_h2o", "vmr_o3", "cld_r_eff_liq", "cld_r_eff_ice",
              "cld_path_liq", "cld_path_ice", "cld_frac", "aero_mass", "aero_size", "sfc_emis", "cos_zenith", "toa_flux",
              "sfc_alb_direct", "sfc_alb_diffuse", "lat", "metric_scaling")
OUTPUT_KEYS = ("lw_flux_up", "lw_flux_dn", "lw_flux_net", "sw_flux_up", "sw_flux_dn", "sw_flux_net", "sw_flux_dn_dir",
               "net_flux", "cld_cover_lw", "cld_cover_sw", "aod_sw_ext", "aod_sw_sca")


class HostPipeline:
    """Host-resident driving of one solver: per step, every per-column input is copied from pinned host memory,
    `update_fluxes!` runs, and every flux / diagnostic is copied back -- as column chunks on a few CUDA streams so
    that the H2D copy of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap (PCIe and the
    SMs are otherwise idle in turn).  This is the end-to-end call for a host whose state lives in host memory;
    a GPU-resident host (the CliMA case) calls `update_fluxes` directly."""

    def __init__(self, s: RRTMGPSolver, n_chunks: int = 8, n_streams: int = 3):
        self.s = s
        ncol = s.grid_params.ncol
        n_chunks = max(1, min(n_chunks, ncol))
        step = -(-ncol // n_chunks)
        # whole "waves" of the persistent kernels (one 12-warp CTA per SM) per chunk avoid a ragged tail per chunk
        wave = 12 * torch.cuda.get_device_properties(s.device).multi_processor_count
        if step > 2 * wave:
            step = max(wave, (step // wave) * wave)
        # ramp up (1, 2, 4 waves, ...) so that the kernels start after a short first copy, and ramp down so that
        # only a short last device-to-host copy is exposed
        sizes, a = [], 0
        ramp = [wave << i for i in range(8) if (wave << i) < step] if step > 2 * wave else []
        for r in ramp:
            if ncol - a > 2 * step:
                sizes.append(r); a += r
        tail = [r for r in reversed(ramp)]
        reserved = sum(tail) if ncol - a > sum(tail) + step else 0
        while ncol - a - reserved > 0:
            n = min(step, ncol - a - reserved)
            sizes.append(n); a += n
        if reserved:
            sizes += tail
        self.chunks, a = [], 0
        for n in sizes:
            self.chunks.append((a, n)); a += n
        assert a == ncol
        self.streams = [torch.cuda.Stream(device=s.device) for _ in range(max(1, n_streams))]
        full = lambda k: (k == "vmr" and s.config.vmr_kind == _lib.VMR_FULL)
        self.in_keys = [k for k in INPUT_KEYS + ("vmr",) if s.buffers.get(k) is not None and (k != "vmr" or full(k))]
        self.out_keys = [k for k in OUTPUT_KEYS if s.buffers.get(k) is not None]
        if s.config.method == _lib.ALL_SKY_WITH_CLEAR:
            self.out_keys += [k for k in s.buffers if k.startswith("clear_") and s.buffers[k] is not None]
        self.host_in = {k: torch.empty_like(s.buffers[k], device="cpu").pin_memory() for k in self.in_keys}
        self.host_out = {k: torch.empty_like(s.buffers[k], device="cpu").pin_memory() for k in self.out_keys}
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.host_in.values())
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.host_out.values())

    def load_host_inputs(self, arrays) -> None:
        alias = {"vmr_full": "vmr"}
        for k, v in arrays.items():
            k = alias.get(k, k)
            if k in self.host_in:
                self.host_in[k].copy_(torch.as_tensor(np.ascontiguousarray(v, dtype=self.s.dtype)))
            elif k == "vmr" and self.s.config.vmr_kind == _lib.VMR_GM:
                self.s.buffers["vmr"].copy_(torch.as_tensor(np.ascontiguousarray(v, dtype=self.s.dtype)))

    def update_fluxes(self, seedval: int = 0) -> None:
        """One end-to-end step; returns after the results are in `self.host_out` (pinned host tensors)."""
        s = self.s
        cur = torch.cuda.current_stream(s.device)
        for st in self.streams:
            st.wait_stream(cur)
        for i, (a, n) in enumerate(self.chunks):
            st = self.streams[i % len(self.streams)]
            with torch.cuda.stream(st):
                for k in self.in_keys:
                    s.buffers[k][a:a + n].copy_(self.host_in[k][a:a + n], non_blocking=True)
                update_fluxes_range(s, seedval, a, n, st)
                for k in self.out_keys:
                    self.host_out[k][a:a + n].copy_(s.buffers[k][a:a + n], non_blocking=True)
        for st in self.streams:
            st.synchronize()


def prepare_atmosphere(s: RRTMGPSolver) -> None:
    check(lib().rrtmgp_b200_prepare_atmosphere(s._h, s._stream()), s._h)


class _Toggle:
    """`RRTMGP.check_values[]` (src/api/validation.jl:14): `check_values.value = True` makes `update_fluxes` call
    `validate_inputs` before each solve.  Off by default; the check synchronises the stream."""
    value = False


check_values = _Toggle()


def validate_inputs(s: RRTMGPSolver) -> None:
    """`validate_inputs(s)` (src/api/validation.jl:56-74): raises on the first input outside its physical range."""
    failed = C.c_uint32(0)
    check(lib().rrtmgp_b200_validate_inputs(s._h, C.byref(failed), s._stream()), s._h)
    for bit, name in enumerate(_lib.INVALID_INPUT_NAMES):
        if failed.value & (1 << bit):
            raise ValueError(f"RRTMGP input validation failed: `{name}` contains values outside its physical range "
                             "(or non-finite values). Inspect the corresponding getter on the solver.")


# the steps of prepare_atmosphere! as the public functions of src/api/grid_adaptation.jl
def interpolate_levels(s: RRTMGPSolver) -> None:
    """`interpolate_levels!` (grid_adaptation.jl:87-113); a no-op with `NoInterpolation`."""
    check(lib().rrtmgp_b200_prepare_steps(s._h, _lib.STEP_INTERPOLATE_LEVELS, s._stream()), s._h)


def add_isothermal_boundary_layer(s: RRTMGPSolver) -> None:
    """`add_isothermal_boundary_layer!` (grid_adaptation.jl:135-195); a no-op without the extra layer."""
    check(lib().rrtmgp_b200_prepare_steps(s._h, _lib.STEP_BOUNDARY_LAYER, s._stream()), s._h)


def clip(s: RRTMGPSolver) -> None:
    """`clip!` (grid_adaptation.jl:232-258): state into the lookup-table ranges, in place."""
    check(lib().rrtmgp_b200_prepare_steps(s._h, _lib.STEP_CLIP, s._stream()), s._h)


def update_concentrations(s: RRTMGPSolver) -> None:
    """`update_concentrations!` (grid_adaptation.jl:278-293): dry-air column amounts from `p_lev` and `vmr_h2o`."""
    check(lib().rrtmgp_b200_prepare_steps(s._h, _lib.STEP_CONCENTRATIONS, s._stream()), s._h)


def update_lw_fluxes(s: RRTMGPSolver, seedval=None) -> None:
    seed, have = _seed_args(s, seedval)
    check(lib().rrtmgp_b200_update_lw_fluxes(s._h, seed, have, s._stream()), s._h)


def update_sw_fluxes(s: RRTMGPSolver, seedval=None) -> None:
    seed, have = _seed_args(s, seedval)
    check(lib().rrtmgp_b200_update_sw_fluxes(s._h, seed, have, s._stream()), s._h)


def update_net_fluxes(s: RRTMGPSolver) -> None:
    check(lib().rrtmgp_b200_update_net_fluxes(s._h, s._stream()), s._h)


def compute_relative_humidity(s: RRTMGPSolver) -> None:
    """`compute_relative_humidity!` (column_amounts.jl:52-76): a host duty (grid_adaptation.jl:267-270)."""
    check(lib().rrtmgp_b200_compute_relative_humidity(s._h, s._stream()), s._h)


# -- state getters (getters.jl:79-152,160-330) ----------------------------------------------
# table ranges the state is clipped to (src/api/grid_adaptation.jl:24-56; `clip!` :232-258)
def get_p_min(s): return float(s.lut_info.p_ref_min)
def get_t_min(s): return float(s.lut_info.t_ref_min)
def get_t_max(s): return float(s.lut_info.t_ref_max)


def center_z(s):
    """`center_z(s)` / `face_z(s)` (getters.jl:255-262): altitudes given at construction, or None."""
    return None if s.center_z is None else s._domain(s.center_z, False)


def face_z(s):
    return None if s.face_z is None else s._domain(s.face_z, True)


def layer_pressure(s): return s._domain(s.buffers["layerdata"][:, :, 1], False)
def layer_temperature(s): return s._domain(s.buffers["layerdata"][:, :, 2], False)
def layer_relative_humidity(s): return s._domain(s.buffers["layerdata"][:, :, 3], False)
def level_pressure(s): return s._domain(s.buffers["p_lev"], True)
def level_temperature(s): return s._domain(s.buffers["t_lev"], True)
def surface_temperature(s): return s.buffers["t_sfc"]
def surface_emissivity(s): return s.buffers["sfc_emis"]
def latitude(s): return s.buffers["lat"]
def cos_zenith(s): return s.buffers["cos_zenith"]
def toa_sw_flux_dn(s): return s.buffers["toa_flux"]
def toa_lw_flux_dn(s): return s.buffers["inc_flux_lw"]
def direct_sw_surface_albedo(s): return s.buffers["sfc_alb_direct"]
def diffuse_sw_surface_albedo(s): return s.buffers["sfc_alb_diffuse"]
def deep_atmosphere_inverse_scaling(s): return s.buffers["metric_scaling"]
def isothermal_boundary_layer(s): return s.grid_params.isothermal_boundary_layer
def radiation_method(s): return s.radiation_method


def _need(s, key, what):
    t = s.buffers.get(key)
    if t is None:
        raise ValueError(f"{what} is not available for {type(s.radiation_method).__name__}")
    return t


def cloud_liquid_effective_radius(s): return s._domain(_need(s, "cld_r_eff_liq", "cloud state"), False)
def cloud_ice_effective_radius(s): return s._domain(_need(s, "cld_r_eff_ice", "cloud state"), False)
def cloud_liquid_water_path(s): return s._domain(_need(s, "cld_path_liq", "cloud state"), False)
def cloud_ice_water_path(s): return s._domain(_need(s, "cld_path_ice", "cloud state"), False)
def cloud_fraction(s): return s._domain(_need(s, "cld_frac", "cloud state"), False)


def aerosol_column_mass_density(s, name: str):
    return s._domain(_need(s, "aero_mass", "aerosol state")[:, :, aerosol_index(name) - 1], False)


def aerosol_radius(s, name: str):
    return s._domain(_need(s, "aero_size", "aerosol state")[:, :, aerosol_index(name) - 1], False)


def volume_mixing_ratio(s, ig: int):
    """`volume_mixing_ratio(s, gas)` (getters.jl:600-640) by 1-based gas index of the lookup tables."""
    if s.config.vmr_kind == _lib.VMR_FULL:
        return s._domain(s.buffers["vmr"][:, :, ig - 1], False)
    if ig == 1:
        return s._domain(s.buffers["vmr_h2o"], False)
    if ig == 3:
        return s._domain(s.buffers["vmr_o3"], False)
    return s.buffers["vmr"][ig - 1]


def set_volume_mixing_ratio(s, ig: int, value) -> None:
    t = volume_mixing_ratio(s, ig)
    t.copy_(torch.as_tensor(value, dtype=s.tdtype, device=s.device).expand(t.shape)) if t.dim() else t.fill_(float(value))


# -- flux and diagnostic getters (getters.jl:340-520) ------------------------------------------
def lw_flux_up(s): return s._domain(s.buffers["lw_flux_up"], True)
def lw_flux_dn(s): return s._domain(s.buffers["lw_flux_dn"], True)
def lw_flux_net(s): return s._domain(s.buffers["lw_flux_net"], True)
def sw_flux_up(s): return s._domain(s.buffers["sw_flux_up"], True)
def sw_flux_dn(s): return s._domain(s.buffers["sw_flux_dn"], True)
def sw_flux_net(s): return s._domain(s.buffers["sw_flux_net"], True)
def sw_direct_flux_dn(s): return s._domain(s.buffers["sw_flux_dn_dir"], True)
def net_flux(s): return s._domain(s.buffers["net_flux"], True)
def clear_lw_flux_up(s): return s._domain(_need(s, "clear_lw_flux_up", "clear-sky flux"), True)
def clear_lw_flux_dn(s): return s._domain(_need(s, "clear_lw_flux_dn", "clear-sky flux"), True)
def clear_lw_flux_net(s): return s._domain(_need(s, "clear_lw_flux_net", "clear-sky flux"), True)
def clear_sw_flux_up(s): return s._domain(_need(s, "clear_sw_flux_up", "clear-sky flux"), True)
def clear_sw_flux_dn(s): return s._domain(_need(s, "clear_sw_flux_dn", "clear-sky flux"), True)
def clear_sw_flux_net(s): return s._domain(_need(s, "clear_sw_flux_net", "clear-sky flux"), True)
def clear_sw_direct_flux_dn(s): return s._domain(_need(s, "clear_sw_flux_dn_dir", "clear-sky flux"), True)
def clear_net_flux(s): return s._domain(_need(s, "clear_net_flux", "clear-sky flux"), True)
def lw_cloud_cover(s): return _need(s, "cld_cover_lw", "cloud cover")
def sw_cloud_cover(s): return _need(s, "cld_cover_sw", "cloud cover")
def aod_sw_extinction(s): return _need(s, "aod_sw_ext", "aerosol optical depth")
def aod_sw_scattering(s): return _need(s, "aod_sw_sca", "aerosol optical depth")
def spectral_lw_flux_up(s): return _need(s, "lw_band_flux_up", "spectral flux")[:, :, : s.grid_params.domain_nlay + 1]
def spectral_lw_flux_dn(s): return _need(s, "lw_band_flux_dn", "spectral flux")[:, :, : s.grid_params.domain_nlay + 1]
def spectral_lw_flux_net(s): return _need(s, "lw_band_flux_net", "spectral flux")[:, :, : s.grid_params.domain_nlay + 1]
def spectral_sw_flux_up(s): return _need(s, "sw_band_flux_up", "spectral flux")[:, :, : s.grid_params.domain_nlay + 1]
def spectral_sw_flux_dn(s): return _need(s, "sw_band_flux_dn", "spectral flux")[:, :, : s.grid_params.domain_nlay + 1]
def spectral_sw_flux_net(s): return _need(s, "sw_band_flux_net", "spectral flux")[:, :, : s.grid_params.domain_nlay + 1]


def heating_rate(s) -> torch.Tensor:
    """`heating_rate(s)` (standalone.jl:106-124): (g / cp) dF_net/dp on the domain layers, from the engine's
    kernel; allocates and returns a fresh `(ncol, nlay)` array like the reference."""
    p = s.params
    cp_d = p.get("cp_d", p["gas_constant"] / p["molmass_dryair"] / p["kappa_d"])
    hr = torch.empty(s.grid_params.ncol, s.grid_params.domain_nlay, dtype=s.tdtype, device=s.device)
    check(lib().rrtmgp_b200_heating_rate(s._h, s.buffers["net_flux"].data_ptr(), hr.data_ptr(), float(cp_d), s._stream()), s._h)
    return hr

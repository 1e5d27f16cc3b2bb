"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/rrtmgp_b200.h declares,
and its argument validation answers with status codes (no GPU compute is attempted)."""
import ctypes as C
import os
import re

import pytest

import rrtmgp_b200 as R
from rrtmgp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    R.build_ext()
    return _lib.lib()


def test_header_and_exports_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "rrtmgp_b200.h")).read()
    declared = set(re.findall(r"\b(rrtmgp_b200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert getattr(raw, name) is not None


def test_abi_version_and_strerror(lib):
    assert lib.rrtmgp_b200_abi_version() == _lib.ABI_VERSION
    assert lib.rrtmgp_b200_strerror(0) == b"ok"
    for code in range(1, 6):
        assert lib.rrtmgp_b200_strerror(code) not in (b"ok", b"unknown status")
    assert lib.rrtmgp_b200_strerror(99) == b"unknown status"


def _cfg(**kw):
    base = dict(abi_version=_lib.ABI_VERSION, device=0, dtype=0, ncol=8, nlay=64, ngas=19, vmr_kind=0, method=1,
                aerosol_radiation=1, op_lw=0, n_gauss_angles=1, ice_rgh=2, spectral_fluxes=0,
                isothermal_boundary_layer=0, col_offset=0, grav=9.81, molmass_dryair=0.02897,
                molmass_water=0.018015, avogad=6.02214076e23)
    base.update(kw)
    return _lib.Config(**base)


@pytest.mark.parametrize("bad", [dict(abi_version=99), dict(dtype=2), dict(ncol=0), dict(nlay=1), dict(method=3),
                                 dict(vmr_kind=5), dict(op_lw=2), dict(n_gauss_angles=0), dict(n_gauss_angles=5),
                                 dict(n_gauss_angles=2, op_lw=0),   # solver.jl:159-171
                                 dict(ice_rgh=0), dict(grav=0.0)])
def test_create_rejects_invalid_configs(lib, bad):
    h = C.c_void_p()
    cfg = _cfg(**bad)
    assert lib.rrtmgp_b200_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID_ARG
    assert not h.value


def test_create_rejects_unsupported_nlay(lib):
    h = C.c_void_p()
    cfg = _cfg(nlay=200)
    assert lib.rrtmgp_b200_create(C.byref(cfg), C.byref(h)) == _lib.ERR_UNSUPPORTED


def test_null_handles_are_errors_not_crashes(lib):
    assert lib.rrtmgp_b200_update_fluxes(None, 0, 0, None) == _lib.ERR_INVALID_ARG
    assert lib.rrtmgp_b200_prepare_atmosphere(None, None) == _lib.ERR_INVALID_ARG
    assert lib.rrtmgp_b200_load_luts(None, b"x", 1) == _lib.ERR_INVALID_ARG
    assert lib.rrtmgp_b200_last_launch_count(None) == 0
    lib.rrtmgp_b200_destroy(None)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under rrtmgp.jl_b200/ may reference it."""
    pkg = os.path.join(ROOT, "rrtmgp.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_solver_fails_loudly_without_cuda():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        R.RRTMGPSolver(R.RRTMGPGridParams(FT=np.float32, domain_nlay=8, ncol=2), R.ClearSkyRadiation(),
                       R.default_parameters(), b"")


def test_aerosol_and_gas_naming_helpers():
    """`aerosol_names` / `aerosol_index` / `aerosol_index_map` / `canonical_aerosol_name` (src/api/aerosols.jl) and
    `gas_names_sw` (src/api/getters.jl:566-588); the ingestion module carries the same aerosol map
    (ext/lookup_constructors.jl:47-58)."""
    import rrtmgp_b200 as R
    assert R.aerosol_names()[0] == "dust1" and len(R.aerosol_names()) == 15
    assert R.aerosol_index("sea_salt3") == 13 and R.aerosol_index_map() == R.tables.AEROSOL_INDEX
    assert R.canonical_aerosol_name("sulfate") == "sulfate"
    with pytest.raises(KeyError, match="known names"):
        R.aerosol_index("soot")
    names = R.gas_names_sw()
    assert len(names) == 21 and set(R.synthetic.GAS_NAMES) | {"h2o_self", "h2o_frgn"} == set(names)
    assert R.requires_z("BestFit") and R.requires_z("HydrostaticBottom") and not R.requires_z("UniformP")


def test_comm_entry_points_validate_arguments(lib):
    """The multi-GPU entry points (include/rrtmgp_b200.h "multi-GPU") answer with status codes without a GPU: a
    short id buffer, NULL handles."""
    small = (C.c_char * 16)()
    assert lib.rrtmgp_b200_comm_unique_id(small, 16) == _lib.ERR_INVALID_ARG
    assert lib.rrtmgp_b200_comm_unique_id(None, 128) == _lib.ERR_INVALID_ARG
    uid = (C.c_char * 128)()
    assert lib.rrtmgp_b200_comm_init(None, uid, 0, 1) == _lib.ERR_INVALID_ARG
    g = _lib.Gathered()
    assert lib.rrtmgp_b200_gathered_buffers(None, C.byref(g)) == _lib.ERR_INVALID_ARG
    assert lib.rrtmgp_b200_comm_destroy(None) == _lib.ERR_INVALID_ARG
    assert lib.rrtmgp_b200_update_fluxes_gathered(None, 0, 1, None) == _lib.ERR_INVALID_ARG
    assert lib.rrtmgp_b200_all_gather_fluxes(None, None) == _lib.ERR_INVALID_ARG
    assert C.sizeof(_lib.Gathered) == 8 * C.sizeof(C.c_void_p) and _lib.UNIQUE_ID_BYTES == 128


def test_julia_shim_binds_every_entry_point():
    """INTEGRATION.md's reference-side binding (`julia/RRTMGPB200Ext.jl`, untested here: no Julia) names every function
    `include/rrtmgp_b200.h` declares, except the one that only the benchmark uses."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "rrtmgp_b200.h")).read()
    shim = open(os.path.join(root, "julia", "RRTMGPB200Ext.jl")).read()
    declared = set(re.findall(r"\b(rrtmgp_b200_[a-z0-9_]+)\s*\(", header))
    bound = set(re.findall(r":(rrtmgp_b200_[a-z0-9_]+)", shim))
    assert declared - bound <= {"rrtmgp_b200_measure_fp32_peak"}, sorted(declared - bound)
    assert bound <= declared, sorted(bound - declared)

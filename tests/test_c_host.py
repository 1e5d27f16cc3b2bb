"""A plain C99 program is a host of the C ABI too: `tests/c_host/abi_probe.c` compiles against `include/rrtmgp_b200.h`
with `gcc -std=c99 -pedantic` (the header is C, not C++), links `librrtmgp_b200.so`, and reports the struct layouts.
They must equal the ctypes mirror (`rrtmgp.jl_b200/_lib.py`) field by field -- the same check the Julia shim's
`struct Config` / `Buffers` (julia/RRTMGPB200Ext.jl:25-37) relies on -- and the argument validation must behave the
same from C as from Python.  No GPU is touched."""
import ctypes as C
import os
import re
import subprocess

import pytest

import rrtmgp_b200 as R
from rrtmgp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    if not os.path.exists(_lib.LIB_PATH):
        R.build_ext()
    exe = str(tmp_path_factory.mktemp("c_host") / "abi_probe")
    csrc = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_host", "abi_probe.c"), "-o", exe,
                           "-L", csrc, "-lrrtmgp_b200", f"-Wl,-rpath,{csrc}"])
    out = subprocess.check_output([exe], text=True)
    return dict(line.split(" ", 1) for line in out.strip().splitlines())


def test_struct_layouts_match_the_ctypes_mirror(probe):
    for name, cls in (("config", _lib.Config), ("buffers", _lib.Buffers), ("lut_info", _lib.LutInfo)):
        assert int(probe[f"sizeof.{name}"]) == C.sizeof(cls), name
        for key, off in probe.items():
            m = re.fullmatch(rf"{name}\.(\w+)", key)
            if m:
                assert getattr(cls, m.group(1)).offset == int(off), key
    assert int(probe["sizeof.buffers"]) == 8 * len(_lib.BUFFER_FIELDS)     # NBUF of the Julia shim
    assert int(probe["abi_version"]) == _lib.ABI_VERSION


def test_julia_shim_declares_the_same_layout():
    """The Julia `struct Config` lists the header's fields in order, and NBUF equals the number of buffer pointers."""
    jl = open(os.path.join(ROOT, "julia", "RRTMGPB200Ext.jl")).read()
    body = jl[jl.index("struct Config"):jl.index("end", jl.index("struct Config"))]
    fields = re.findall(r"(\w+)::(Int32|Int64|Float64)", body)
    want = [(n, {C.c_int32: "Int32", C.c_int64: "Int64", C.c_double: "Float64"}[t]) for n, t in _lib.Config._fields_]
    assert fields == want
    assert int(re.search(r"const NBUF = (\d+)", jl).group(1)) == len(_lib.BUFFER_FIELDS)


def test_argument_validation_from_c(probe):
    assert int(probe["create.bad_abi"]) == _lib.ERR_INVALID_ARG and probe["create.bad_abi.handle_null"] == "1"
    assert int(probe["create.angles_without_noscat"]) == _lib.ERR_INVALID_ARG      # solver.jl:159-171
    assert int(probe["create.null"]) == _lib.ERR_INVALID_ARG and int(probe["update.null"]) == _lib.ERR_INVALID_ARG
    assert probe["strerror.ok"] == "ok" and "LUT pack" in probe["strerror.bad_pack"]

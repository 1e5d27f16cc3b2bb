#!/usr/bin/env python
"""Operation count of the restated reference algorithm on the headline workload ("pin by op-counting
the oracle").  Builds oracle/libopcount.so (oracle/opcount.cpp: the oracle with its Float64 type replaced by a counting
wrapper), runs `update_fluxes` single-threaded on a few columns of the all-sky-with-aerosols benchmark atmosphere and
writes profiles/oracle_opcount.json: operations per column by kind, for the LW and the SW solve.

CPU only; test / measurement infrastructure like the oracle itself."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "oracle", "libopcount.so")
subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-fno-strict-aliasing",
                       os.path.join(ROOT, "oracle", "opcount.cpp"), "-o", LIB])
os.environ["RRTMGP_ORACLE_LIB"] = LIB

import numpy as np  # noqa: E402

import rrtmgp_b200 as R  # noqa: E402
import oracle  # noqa: E402

NAMES = ("add", "mul", "div", "fma_like", "sqrt", "exp", "expm1", "log", "trig", "pow", "cmp", "minmax")
PARAMS = dict(grav=9.80665, molmass_dryair=0.028964, molmass_water=0.018016)


def read(reset=True):
    buf = (C.c_ulonglong * 12)()
    oracle.lib().oracle_opcount_read(buf, int(reset))
    return dict(zip(NAMES, [int(x) for x in buf]))


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncol", type=int, default=48)
    ap.add_argument("--nlay", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "oracle_opcount.json"))
    args = ap.parse_args()
    ncol, nlay = args.ncol, args.nlay
    pack = R.synthetic.make_lut_pack(seed=7)
    st = R.synthetic.make_atmosphere(ncol, nlay, seed=20260101, cld_frac=1.0, cos_zenith=0.86, dtype=np.float64)
    o = oracle.Oracle(pack, np.float64)
    read()
    out = {"workload": f"all_sky_with_aerosols, nlay={nlay}, 256 LW + 224 SW g-points, cld_frac=1, cos_zenith=0.86, {ncol} columns "
                       "(synthetic LUT pack seed 7); counts per column", "per_column": {}}
    for tag, kw in (("prepare", dict(do_lw=False, do_sw=False)), ("lw", dict(prepare=False, do_sw=False)),
                    ("sw", dict(prepare=False, do_lw=False))):
        o.update_fluxes(st, seed=1, params=PARAMS, nthreads=1, **kw)
        c = read()
        per = {k: v / ncol for k, v in c.items()}
        per["flops_add_mul_div"] = per["add"] + per["mul"] + per["div"]
        per["transcendental"] = per["sqrt"] + per["exp"] + per["expm1"] + per["log"] + per["trig"] + per["pow"]
        out["per_column"][tag] = per
    tot = {k: sum(out["per_column"][t][k] for t in ("prepare", "lw", "sw")) for k in out["per_column"]["lw"]}
    out["per_column"]["total"] = tot
    out["note"] = ("restated reference algorithm (oracle/rrtmgp_oracle.cpp, g-point-outer loops as in the reference's CPU path): "
                   "every +, -, *, / on the floating-point type counts one, math functions are counted by kind, comparisons "
                   "and min/max separately; the reference recomputes the per-layer interpolation indices and fractions for every "
                   "g-point, which the engine hoists (DESIGN.md §4), so this is an upper bound of the necessary work")
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    for t, per in out["per_column"].items():
        print(t, {k: round(v) for k, v in per.items()})


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Regenerates tests/golden/allsky_real_dims.npz: Float64 oracle fluxes for 8 seeded all-sky-with-aerosols
columns on the seed-7 synthetic tables (real table dimensions).  The reference is Julia and cannot run in
this image, so these vectors pin the RESTATED reference (oracle/), itself pinned by the reference's
data-free tests and by tests/npref.py; they guard against silent changes of oracle, generators and engine.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rrtmgp_b200 as R  # noqa: E402
from oracle import Oracle  # noqa: E402

CASE = dict(ncol=8, nlay=64, seed_state=424242, seed_lut=7, seed_mcica=2026)


def inputs():
    return R.synthetic.make_atmosphere(CASE["ncol"], CASE["nlay"], seed=CASE["seed_state"], dtype=np.float64,
                                       cld_frac=None, cos_zenith=None)


if __name__ == "__main__":
    pack = R.synthetic.make_lut_pack(seed=CASE["seed_lut"])
    r = Oracle(pack, np.float64).update_fluxes(inputs(), seed=CASE["seed_mcica"], method="all_sky_with_clear")
    keep = {k: r[k] for k in ("lw_up", "lw_dn", "sw_up", "sw_dn", "sw_dir", "net", "clear_lw_up", "clear_sw_dn",
                              "cld_cover_lw", "cld_cover_sw", "aod_sw_ext", "aod_sw_sca")}
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "allsky_real_dims.npz"), **keep)
    print({k: v.shape for k, v in keep.items()})

"""The C++ oracle against an independent numpy restatement of the same Julia sources (tests/npref.py)."""
import numpy as np
import pytest

import npref
import rrtmgp_b200 as R
from oracle import Oracle


def _compare(arrays, pack, st, cols, **kw):
    o = Oracle(pack, np.float64).update_fluxes(st, seed=1, **kw)
    prepared = dict(st)
    prepared.update({k: o["state"][k] for k in ("layerdata", "p_lev", "t_lev", "vmr_h2o")})
    for c in cols:
        r = npref.solve_column(arrays, prepared, c, clouds=kw.get("method", "all_sky") != "clear_sky",
                               aerosols=kw.get("aerosols", True))
        for k in ("lw_up", "lw_dn", "sw_up", "sw_dn", "sw_dir"):
            scale = max(1.0, np.abs(o[k][c]).max())
            assert np.abs(r[k] - o[k][c]).max() <= 1e-10 * scale, (c, k)
        if "aod_sw_ext" in o:
            np.testing.assert_allclose(r["aod"], [o["aod_sw_ext"][c], o["aod_sw_sca"][c]], rtol=1e-12)


def test_small_tables_all_sky_with_aerosols():
    dims = R.synthetic.LutDims(n_bnd_lw=4, n_bnd_sw=4, gpts_lw=[4, 6, 3, 5], gpts_sw=[5, 2, 6, 4],
                               nsize_liq=8, nsize_ice=7, nrh=9)
    arrays = R.synthetic.make_lut_arrays(seed=5, dims=dims)
    pack = R.lutpack.pack_luts(arrays)
    st = R.synthetic.make_atmosphere(9, 24, dtype=np.float64, n_bnd_lw=4, n_bnd_sw=4, cld_frac=1.0, cos_zenith=None,
                                     z_top=40.0e3)
    _compare(arrays, pack, st, range(9))


def test_small_tables_clear_sky():
    arrays = R.synthetic.make_lut_arrays(seed=11, dims=R.synthetic.SMALL_DIMS)
    pack = R.lutpack.pack_luts(arrays)
    st = R.synthetic.make_atmosphere(4, 30, dtype=np.float64, n_bnd_lw=3, n_bnd_sw=3, clouds=False, aerosols=False)
    _compare(arrays, pack, st, range(4), method="clear_sky", aerosols=False)


def test_real_dims_two_columns():
    arrays = R.synthetic.make_lut_arrays(seed=7)
    pack = R.lutpack.pack_luts(arrays)
    st = R.synthetic.make_atmosphere(3, 64, dtype=np.float64, cld_frac=1.0)
    _compare(arrays, pack, st, (0, 1))

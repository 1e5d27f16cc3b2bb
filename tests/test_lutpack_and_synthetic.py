"""LUT pack container and the structural faithfulness of the synthetic tables."""
import numpy as np
import pytest

import rrtmgp_b200 as R


def test_pack_roundtrip_preserves_shapes_and_values():
    arrays = R.synthetic.make_lut_arrays(seed=3, dims=R.synthetic.SMALL_DIMS)
    back = R.lutpack.unpack_luts(R.lutpack.pack_luts(arrays))
    assert set(back) == set(arrays)
    for k, v in arrays.items():
        assert back[k].shape == np.asarray(v).shape, k
        np.testing.assert_array_equal(back[k], np.asarray(v).astype(back[k].dtype))


def test_pack_rejects_corruption():
    buf = R.lutpack.pack_luts({"x": np.arange(4.0)})
    with pytest.raises(ValueError):
        R.lutpack.unpack_luts(b"garbage" + buf[7:])
    with pytest.raises(ValueError):
        R.lutpack.unpack_luts(buf[:-8])


def test_real_dims_match_the_reference_tables():
    """docs/src/Optics.md:125-126,191-211: n_eta 9, 59 reference pressures (60 kmajor nodes),
    14 reference temperatures, 196 Planck temperatures, 256/224 g-points in 16/14 bands."""
    a = R.synthetic.make_lut_arrays(seed=7)
    assert a["lw/kmajor"].shape == (9, 60, 14, 256) and a["sw/kmajor"].shape == (9, 60, 14, 224)
    assert a["lw/planck_fraction"].shape == (9, 60, 14, 256)
    assert a["lw/tot_planck"].shape == (196, 16) and a["lw/p_ref"].shape == (59,)
    assert a["lw/key_species"].shape == (2, 2, 16) and a["sw/key_species"].shape == (2, 2, 14)
    assert a["sw/rayl_lower"].shape == (9, 14, 224)
    # duplicated tropopause node: p_ref_tropo is the 13th reference pressure
    assert a["lw/params"][0] == a["lw/p_ref"][12]
    # Planck fractions sum to one over the g-points of a band; solar fractions sum to one
    pf = a["lw/planck_fraction"]
    lims = a["lw/bnd_lims_gpt"]
    for b in range(16):
        s = pf[..., lims[0, b] - 1:lims[1, b]].sum(axis=3)
        np.testing.assert_allclose(s, 1.0, rtol=1e-12)
    assert a["sw/solar_src_scaled"].sum() == pytest.approx(1.0, rel=1e-12)
    # band-integrated Planck function adds up to sigma T^4 / pi
    t = a["lw/t_planck"]
    np.testing.assert_allclose(np.pi * a["lw/tot_planck"].sum(axis=1), 5.670374419e-8 * t ** 4, rtol=5e-3)
    # 550 nm lies in a shortwave band
    assert 1 <= a["aero_sw/iband_550nm"][0] <= 14 and a["aero_lw/iband_550nm"][0] == 0


def test_minor_csr_is_consistent():
    a = R.synthetic.make_lut_arrays(seed=7)
    for pre in ("lw", "sw"):
        g2b = a[f"{pre}/major_gpt2bnd"]
        for tag in ("minor_lower", "minor_upper"):
            bst, gst = a[f"{pre}/{tag}/bnd_st"], a[f"{pre}/{tag}/gpt_st"]
            gd, km = a[f"{pre}/{tag}/gasdata"], a[f"{pre}/{tag}/kminor"]
            assert gd.shape[0] == 4 and gd.shape[1] == bst[-1] - 1
            n_per_g = np.diff(gst)
            np.testing.assert_array_equal(n_per_g, np.diff(bst)[g2b - 1])
            assert km.shape[2] >= gst[-1] - 1
            assert set(np.unique(gd[2])) <= {0, 1} and set(np.unique(gd[3])) <= {0, 1}
        # every flag combination of compute_tau_minor (gas_optics.jl:383-399) is exercised
        combos = {tuple(r[1:] > 0) for r in a[f"{pre}/minor_lower/gasdata"].T}
        assert len(combos) >= 3


def test_atmosphere_recipe():
    st = R.synthetic.make_atmosphere(30, 64, cld_frac=None, cos_zenith=None)
    assert st["layerdata"].shape == (30, 64, 4) and st["p_lev"].shape == (30, 65)
    assert (np.diff(st["p_lev"], axis=1) < 0).all()          # level 1 is the surface
    assert (st["cld_frac"][2::3] == 0).all()                  # every third column is cloud free
    assert ((st["aero_mass"] > 0).sum(axis=2) == 1).all()     # one species per (layer, column)
    assert (st["cos_zenith"] <= 0).any() and (st["cos_zenith"] > 0).any()
    rh = R.synthetic.relative_humidity(st["layerdata"][:, :, 1], st["layerdata"][:, :, 2], st["vmr_h2o"])
    np.testing.assert_array_equal(rh, st["layerdata"][:, :, 3])


def test_float32_parity_gate_rules():
    """The per-column Float32 gate of the GPU suite (tests/helpers.py::gate_f32) on hand-made arrays: within the
    threshold passes; above it passes only beside a comparably wrong Float32 oracle in the SAME column; the strict
    variant also wants the Float32 oracle above the threshold there, or the engine within a tenth of the threshold of it."""
    import numpy as np
    import pytest
    from helpers import LEDGER, gate_f32
    ref = np.zeros((4, 5))
    tol = 0.1
    ok = ref + 0.05
    n0 = len(LEDGER)
    assert gate_f32("k", ok, ref, tol)["bar"] == "threshold"
    eng = ref.copy(); eng[2, 1] = 0.14                       # one column above the threshold
    with pytest.raises(AssertionError):
        gate_f32("k", eng, ref, tol)                         # no Float32 oracle to excuse it
    o32 = ref.copy(); o32[1, 1] = 0.2                        # the oracle is wrong in ANOTHER column: no excuse
    with pytest.raises(AssertionError):
        gate_f32("k", eng, ref, tol, o32)
    o32 = ref.copy(); o32[2, 1] = 0.12                       # same column, engine <= 1.5 x oracle error, oracle above the threshold
    row = gate_f32("k", eng, ref, tol, o32, strict=True)
    assert row["passed"] and row["columns_excused_by_f32_oracle"] == 1 and row["engine_max_err_column"] == 2
    o32[2, 1] = 0.099                                        # oracle just below the threshold, engine 0.14: |e - o32| = 0.041 > 0.01
    gate_f32("k", eng, ref, tol, o32)                        # the plain rule excuses it (0.14 <= 1.5 * 0.099)
    with pytest.raises(AssertionError):
        gate_f32("k", eng, ref, tol, o32, strict=True)       # the strict rule does not
    eng[2, 1] = 0.105                                        # engine tracks the oracle to within a tenth of the threshold
    assert gate_f32("k", eng, ref, tol, o32, strict=True)["passed"]
    eng[2, 1] = 0.31; o32[2, 1] = 0.2                        # more than 1.5 x the oracle's error: never
    with pytest.raises(AssertionError):
        gate_f32("k", eng, ref, tol, o32)
    del LEDGER[n0:]                                          # these rows are not parity evidence

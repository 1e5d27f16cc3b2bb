"""Properties of the restated reference on the synthetic spectral tables (CPU only).

The reference's Fortran-flux comparisons need the rrtmgp-data artifact (absent here), so the spectral
path is pinned by the identities the reference itself asserts on real data: clear vs all-sky ordering
(test/all_sky_with_aerosols_utils.jl:190-197), AOD ordering (:364-367), band sums (:233-248), zenith
edge cases (test/cos_zenith_edge_cases.jl:199-240), McICA semantics (test/partial_cloud_fraction.jl:
113-193), Float32 <-> Float64 consistency (test/float32_consistency.jl:53-62), and read-path
invariance (test/clear_sky_utils.jl:149-160 -> here: VmrGM vs Vmr storage, column sharding)."""
import numpy as np
import pytest

import rrtmgp_b200 as R
from oracle import Oracle, lib

FLUX = ("lw_up", "lw_dn", "lw_net", "sw_up", "sw_dn", "sw_net", "sw_dir", "net")


@pytest.fixture(scope="module")
def o64(real_pack):
    return Oracle(real_pack, np.float64)


@pytest.fixture(scope="module")
def o32(real_pack):
    return Oracle(real_pack, np.float32)


def test_fluxes_are_physical(o64):
    st = R.synthetic.make_atmosphere(24, 64, dtype=np.float64)
    r = o64.update_fluxes(st, seed=1)
    for k in FLUX:
        assert np.isfinite(r[k]).all(), k
    assert (r["lw_up"] > 0).all() and (r["lw_dn"] >= 0).all() and (r["sw_dn"] >= r["sw_dir"] - 1e-9).all()
    np.testing.assert_allclose(r["lw_net"], r["lw_up"] - r["lw_dn"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(r["net"], r["lw_net"] + r["sw_net"], rtol=0, atol=1e-12)
    # surface LW emission ~ sigma T^4 with emissivity 0.98 (+ reflected downwelling)
    sig = 5.670374419e-8 * r["state"]["t_sfc"] ** 4
    assert (np.abs(r["lw_up"][:, 0] - (0.98 * sig + 0.02 * r["lw_dn"][:, 0])) < 0.01 * sig).all()
    # TOA direct beam = F0 mu0
    np.testing.assert_allclose(r["sw_dir"][:, -1], st["toa_flux"] * st["cos_zenith"], rtol=1e-9)
    # net SW absorbed by the column is non-negative, OLR below surface emission
    assert (r["sw_dn"][:, -1] - r["sw_up"][:, -1] >= 0).all() and (r["lw_up"][:, -1] < r["lw_up"][:, 0]).all()


def test_clear_vs_all_sky_ordering_and_aod(o64):
    st = R.synthetic.make_atmosphere(48, 64, dtype=np.float64)
    r = o64.update_fluxes(st, seed=2, method="all_sky_with_clear")
    assert (r["clear_lw_up"][:, -1] >= r["lw_up"][:, -1] - 1e-9).all()      # clouds lower the OLR
    assert (r["sw_up"][:, -1] >= r["clear_sw_up"][:, -1] - 1e-9).all()      # and brighten the planet
    cloudy = (st["cld_frac"] > 0).any(axis=1)
    assert (r["clear_lw_up"][cloudy, -1] > r["lw_up"][cloudy, -1]).all()
    np.testing.assert_array_equal(r["clear_lw_up"][~cloudy], r["lw_up"][~cloudy])
    assert (r["aod_sw_ext"] >= r["aod_sw_sca"]).all() and (r["aod_sw_sca"] >= 0).all()
    # the diagnostics solve is exactly the clear-sky method (aerosols included)
    rc = o64.update_fluxes(st, seed=2, method="clear_sky")
    for k in FLUX:
        np.testing.assert_array_equal(r["clear_" + k], rc[k])


def test_band_fluxes_sum_to_broadband(o64):
    st = R.synthetic.make_atmosphere(8, 64, dtype=np.float64)
    r = o64.update_fluxes(st, seed=3, spectral=True)
    np.testing.assert_allclose(r["lw_band_up"].sum(0), r["lw_up"], rtol=1e-12)
    np.testing.assert_allclose(r["lw_band_dn"].sum(0), r["lw_dn"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(r["sw_band_up"].sum(0), r["sw_up"], rtol=1e-12)
    np.testing.assert_allclose(r["sw_band_dn"].sum(0), r["sw_dn"], rtol=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cos_zenith_edge_cases(real_pack, dtype):
    o = Oracle(real_pack, dtype)
    st = R.synthetic.make_atmosphere(4, 64, dtype=dtype)
    st["cos_zenith"][:] = np.array([0.5, 0.0, 1e-10, -0.5], dtype=dtype)
    r = o.update_fluxes(st, seed=4)
    for k in FLUX:
        assert np.isfinite(r[k]).all(), k
    for k in ("sw_up", "sw_dn", "sw_net", "sw_dir"):
        assert (r[k][[1, 3]] == 0).all()
    assert (r["sw_dn"][0] > 0).all()
    # night columns still sample masks and optics: cloud cover and AOD are reported
    assert r["cld_cover_sw"][1] == r["cld_cover_lw"][1] and r["aod_sw_ext"][3] > 0


def test_mcica_semantics(o64):
    st = R.synthetic.make_atmosphere(60, 64, dtype=np.float64, cld_frac=None, aerosols=False)
    a = o64.update_fluxes(st, seed=42, aerosols=False, masks=True)
    b = o64.update_fluxes(st, seed=42, aerosols=False, masks=True)
    c = o64.update_fluxes(st, seed=43, aerosols=False, masks=True)
    np.testing.assert_array_equal(a["net"], b["net"])                 # seeded: reproducible
    assert not np.array_equal(a["mask_lw"], c["mask_lw"])              # another seed: another sample
    assert not np.array_equal(a["mask_lw"], a["mask_sw"][:, :, :])     # LW and SW draw separately
    for k in ("cld_cover_lw", "cld_cover_sw"):
        assert ((a[k] >= 0) & (a[k] <= 1)).all()
    cf = st["cld_frac"]
    assert not a["mask_lw"][:, cf == 0].any()                          # never cloudy where cld_frac == 0
    # sampled cloud frequency tracks the cloud fraction (max-random overlap is unbiased per layer)
    freq = a["mask_lw"].mean(axis=0)
    sel = cf > 0
    assert abs(freq[sel].mean() - cf[sel].mean()) < 0.02
    # maximum overlap: adjacent cloudy layers are at least as correlated as independent draws
    m = a["mask_lw"].astype(bool)
    both = (m[:, :, 1:] & m[:, :, :-1]).mean()
    indep = (m[:, :, 1:].mean(axis=0) * m[:, :, :-1].mean(axis=0)).mean()
    assert both > indep
    # cld_frac = 1 is deterministic (every draw r >= 0 passes)
    st1 = R.synthetic.make_atmosphere(20, 64, dtype=np.float64, cld_frac=1.0, aerosols=False)
    x = o64.update_fluxes(st1, seed=1, aerosols=False)
    y = o64.update_fluxes(st1, seed=2, aerosols=False)
    np.testing.assert_array_equal(x["net"], y["net"])


def test_mcica_uniform_generator():
    u = np.array([lib().oracle_mcica_rand(7, c, s, g, k) for c in range(40) for s in (0, 1) for g in (1, 77) for k in range(1, 40)])
    assert ((u >= 0) & (u < 1)).all()
    assert abs(u.mean() - 0.5) < 0.02 and abs(u.var() - 1 / 12) < 0.01
    assert len(np.unique(u)) == u.size


def test_float32_consistency_spectral(o32, o64):
    """test/float32_consistency.jl:53-62 thresholds (LW 1e-3; SW 1.2e-1 cloudy) on the synthetic tables.

    The reference floors the two-stream eigenvalue at k_min = sqrt(eps(FT)) (src/Numerics.jl:24), so
    Float32 and Float64 solve slightly different equations where 2(1 - ssa)(gamma1 + gamma2) < 3.5e-4,
    i.e. for almost conservative Rayleigh g-points.  The synthetic SW tables have a few such g-points
    in a few columns; there the reference's own Float32 arithmetic departs from Float64 by up to
    ~0.3 W/m2.  The CI threshold is therefore asserted on the 99th percentile of the per-column error
    and the maximum is bounded separately."""
    st = R.synthetic.make_atmosphere(200, 64)
    a, b = o32.update_fluxes(st, seed=5), o64.update_fluxes(st, seed=5)
    d = lambda k: np.abs(a[k].astype(np.float64) - b[k]).max(axis=1)
    assert max(d("lw_up").max(), d("lw_dn").max()) <= 1.0e-3
    sw = np.maximum(d("sw_up"), d("sw_dn"))
    assert np.percentile(sw, 99) <= 1.2e-1
    assert sw.max() <= 0.5


def test_vmr_storage_and_sharding_invariance(o64):
    """Read-path invariance: VmrGM vs Vmr storage and any column sharding give bitwise equal fluxes."""
    gm = R.synthetic.make_atmosphere(30, 64, dtype=np.float64, cld_frac=None)
    full = R.synthetic.make_atmosphere(30, 64, dtype=np.float64, cld_frac=None, vmr_kind="full")
    a, b = o64.update_fluxes(gm, seed=9), o64.update_fluxes(full, seed=9)
    for k in FLUX:
        np.testing.assert_array_equal(a[k], b[k])
    from rrtmgp_b200.sharding import shard_range, shard_state
    parts = []
    for rank in range(4):
        lo, hi = shard_range(30, rank, 4)
        parts.append(o64.update_fluxes(shard_state(gm, 30, rank, 4), seed=9, col_offset=lo))
    for k in FLUX + ("cld_cover_lw", "aod_sw_ext"):
        np.testing.assert_array_equal(np.concatenate([p[k] for p in parts]), a[k])


def test_prepare_clips_and_computes_col_dry(o64):
    """clip! (grid_adaptation.jl:232-258) and compute_col_gas_kernel! (gas_optics.jl:16-41)."""
    st = R.synthetic.make_atmosphere(6, 64, dtype=np.float64, with_lat=True)
    st["vmr_h2o"][0, 3] = -1e-3
    st["t_lev"][1, 0] = 400.0
    st["layerdata"][2, 5, 2] = 100.0
    st["p_lev"][3, -1] = 0.01
    r = o64.update_fluxes(st, seed=1)
    s = r["state"]
    assert s["vmr_h2o"][0, 3] == 0 and s["t_lev"][1, 0] == 355.0 and s["layerdata"][2, 5, 2] == 160.0
    assert s["p_lev"][3, -1] == pytest.approx(109663.0 * np.exp(-0.2 * 58))
    g0 = 9.80665 - 0.02586 * np.cos(2 * np.pi * st["lat"] / 180)          # sic: gas_optics.jl:32
    dp = s["p_lev"][:, :-1] - s["p_lev"][:, 1:]
    expect = dp * 6.02214076e23 / (1e4 * (0.028964 + 0.018016 * s["vmr_h2o"]) * g0[:, None])
    np.testing.assert_allclose(s["layerdata"][:, :, 0], expect, rtol=1e-13)
    for k in FLUX:
        assert np.isfinite(r[k]).all()


def test_noscat_vs_two_stream_and_angles(o64):
    """The two LW solvers agree to a few W/m2 (test/runtests.jl:46-49 tolerates 4.5-5 between them); more
    quadrature angles change the flux only slightly (test/angular_discretization.jl)."""
    st = R.synthetic.make_atmosphere(12, 64, dtype=np.float64)
    two = o64.update_fluxes(st, seed=1, method="clear_sky", aerosols=False)
    one = o64.update_fluxes(st, seed=1, method="clear_sky", aerosols=False, lw_noscat=True)
    four = o64.update_fluxes(st, seed=1, method="clear_sky", aerosols=False, lw_noscat=True, n_gauss_angles=4)
    assert np.abs(two["lw_up"] - one["lw_up"]).max() < 5.0 and np.abs(two["lw_dn"] - one["lw_dn"]).max() < 5.0
    assert 0 < np.abs(four["lw_up"] - one["lw_up"]).max() < 5.0

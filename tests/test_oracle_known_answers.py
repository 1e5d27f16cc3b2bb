"""Pins the CPU oracle against the reference's own data-free known-answer tests.

Each test restates a testset of the reference (file:line in the docstring) with the same
inputs and the same pass criteria, run against oracle/rrtmgp_oracle.cpp.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import lib


# ------------------------------------------------------------------------------------------
# test/optics_utils.jl:5-41 -- exact equalities
# ------------------------------------------------------------------------------------------
def test_loc_lower_and_interp1d_exact_values():
    L = lib()
    xeq = np.arange(0.0, 1.5 + 1e-12, 0.05)  # Vector(0:0.05:1.5)
    xeq = np.array([i * 0.05 for i in range(31)])
    neq, dx = xeq.size, 0.05
    assert L.oracle_loc_lower_eq(-0.3, dx, neq, xeq) == 1
    assert L.oracle_loc_lower_eq(1.55, dx, neq, xeq) == neq - 1
    assert L.oracle_loc_lower_eq(0.72, dx, neq, xeq) == 15
    assert L.oracle_loc_lower_eq(1.1, dx, neq, xeq) == 23

    x = np.concatenate([np.array([i * 0.05 for i in range(17)]), np.array([0.825 + i * 0.025 for i in range(28)])])
    n = x.size
    assert L.oracle_loc_lower(-0.3, x, n) == 1
    assert L.oracle_loc_lower(1.55, x, n) == n - 1
    assert L.oracle_loc_lower(0.72, x, n) == 15
    assert L.oracle_loc_lower(1.02, x, n) == 25
    assert L.oracle_loc_lower(1.10, x, n) == 29

    yeq = 3 * xeq + 4
    f = L.oracle_interp1d_equispaced
    assert f(-0.3, xeq, yeq, neq) == yeq[0]
    assert f(1.55, xeq, yeq, neq) == yeq[-1]
    assert f(0.72, xeq, yeq, neq) == pytest.approx(yeq[14] * (1 - 0.4) + yeq[15] * 0.4, rel=1e-14)
    assert f(1.10, xeq, yeq, neq) == pytest.approx(yeq[22], rel=1e-14)

    fac = C.c_double()
    assert (L.oracle_interp1d_loc_factor(-0.3, x, n, C.byref(fac)), fac.value) == (1, 0.0)
    assert (L.oracle_interp1d_loc_factor(1.55, x, n, C.byref(fac)), fac.value) == (n - 1, 1.0)
    loc = L.oracle_interp1d_loc_factor(1.02, x, n, C.byref(fac))
    assert loc == 25 and fac.value == pytest.approx(0.8, rel=1e-9)


# ------------------------------------------------------------------------------------------
# test/angular_discretization.jl
# ------------------------------------------------------------------------------------------
def _angles(n):
    D, w = np.zeros(4), np.zeros(4)
    lib().oracle_gauss_angles(n, D, w)
    return D[:n], w[:n]


def _two_E3(tau, n=100_001):
    """2 E3(tau) by Simpson quadrature in mu (test/angular_discretization.jl:31-39)."""
    mu = np.linspace(0.0, 1.0, n)
    f = np.zeros(n)
    f[1:] = np.exp(-tau / mu[1:]) * mu[1:]
    h = 1.0 / (n - 1)
    s = f[0] + f[-1] + 4 * f[1:-1:2].sum() + 2 * f[2:-1:2].sum()
    return 2 * s * h / 3


def test_quadrature_weights_and_secants():
    """:43-67"""
    for n in range(1, 5):
        D, w = _angles(n)
        assert w.sum() == pytest.approx(1.0, rel=1e-8)
        assert (w > 0).all() and (D > 1).all()
        assert (np.diff(D) < 0).all() or n == 1


def test_more_angles_integrate_the_hemisphere_better():
    """:72-88"""
    taus = (0.05, 0.2, 0.5, 1.0, 2.0, 5.0)
    worst = []
    for n in range(1, 5):
        D, w = _angles(n)
        worst.append(max(abs((w * np.exp(-t * D)).sum() - _two_E3(t)) for t in taus))
    assert worst == sorted(worst, reverse=True)
    assert worst[0] > 1e-2
    assert worst[3] < 1e-3


def test_one_angle_transport_exact_for_isothermal_layer():
    """:102-153: one angle gives exactly pi w B (1 - exp(-tau D)) at the surface (rtol 1e-14);
    the four-angle sum is within 2e-3 of pi B (1 - 2 E3(tau))."""
    B, tau_layer = 0.5, 0.7
    tau = np.array([tau_layer])
    lay = np.array([B])
    lev = np.array([B, B])

    def sfc_dn(Ds, w):
        up, dn = np.zeros(2), np.zeros(2)
        lib().oracle_lw_noscat_one_angle(1, tau, lay, lev, B, 1.0, 0, 0.0, Ds, w, up, dn)
        return dn[0]

    total = 0.0
    for n in range(1, 5):
        D, w = _angles(n)
        if n == 4:
            total = sum(sfc_dn(D[i], w[i]) for i in range(n))
        for i in range(n):
            expected = np.pi * w[i] * B * (1 - np.exp(-tau_layer * D[i]))
            assert sfc_dn(D[i], w[i]) == pytest.approx(expected, rel=1e-14)
    exact = np.pi * B * (1 - _two_E3(tau_layer))
    assert total == pytest.approx(exact, rel=2e-3)


def test_single_angle_default_secant():
    """AngularDiscretizations.jl:42-43 (also SURVEY.md Appendix A.17)."""
    D, w = _angles(1)
    assert D[0] == 1.0 / 0.6096748751 and w[0] == 1.0


# ------------------------------------------------------------------------------------------
# test/gray_atm_utils.jl
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("two_stream", [False, True])
def test_gray_sw_direct_beam(dtype, two_stream):
    """:144-234: surface direct flux == F0 mu0 exp(-sum(tau)/mu0), rel < 1e-3."""
    ncol, nlay = 9, 60
    st = oracle.gray_setup(dtype, np.linspace(-90, 90, ncol), nlay)
    mu0 = np.cos(np.pi / 180 * 52.95)
    r = oracle.gray_solve_sw(st, oracle.OTP_OGORMAN2008, two_stream=two_stream, cos_zenith=mu0, toa_flux=1407.679,
                             albedo=0.1)
    exact = 1407.679 * mu0 * np.exp(-r["tau"][0].astype(np.float64).sum() / mu0)
    assert abs(r["dir"][0, 0] - exact) / exact < 1e-3


@pytest.mark.parametrize("two_stream", [False, True])
def test_gray_lw_radiative_equilibrium(two_stream):
    """:28-142: integrate the gray LW problem (Schneider 2004 optical depth, 6 h steps) to
    radiative equilibrium (max |dF_net| < 1e-5); the level temperatures then match the
    radiative-equilibrium profile implied by the fluxes to < 0.1 K, for both LW solvers."""
    dtype, ncol, nlay = np.float64, 9, 60
    st = oracle.gray_setup(dtype, np.linspace(-90, 90, ncol), nlay)
    dt = 60.0 * 60.0 * 6.0
    nsteps = int(365 * 40 * 4)
    t_ex = None
    err = np.inf
    for _ in range(nsteps):
        f = oracle.gray_solve_lw(st, oracle.OTP_SCHNEIDER2004, two_stream=two_stream, sfc_emis=1.0)
        hr = oracle.gray_heating_rate(f["net"], st["p_lev"])
        t_ex, grad = oracle.gray_update_profile(st, hr, f["dn"], f["net"], dt)
        err = grad.max()
        if err < 1e-5:
            break
    assert err < 1e-5
    assert np.abs(t_ex - st["t_lev"]).max() < 0.1


# ------------------------------------------------------------------------------------------
# test/api_contract.jl
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_incident_longwave_flux_boundary_condition(dtype):
    """:198-221"""
    r0 = oracle.solve_gray(dtype, nlay=60, ncol=4)
    assert (r0["lw"]["dn"][:, -1] == 0).all()
    r = oracle.solve_gray(dtype, nlay=60, ncol=4, inc_flux=25.0)
    np.testing.assert_allclose(r["lw"]["dn"][:, -1], 25.0, rtol=1e-6)
    assert (r["lw"]["dn"][:, 0] > r0["lw"]["dn"][:, 0]).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_metric_scaling_multiplies_fluxes(dtype):
    """:226-260"""
    nlay, ncol = 20, 4
    base = oracle.solve_gray(dtype, nlay=nlay, ncol=ncol)
    sc = oracle.solve_gray(dtype, nlay=nlay, ncol=ncol, scaling=np.full((ncol, nlay + 1), 2.0))
    rt = 1e-6 if dtype == np.float32 else 1e-12
    for grp, keys in (("lw", ("up", "dn", "net")), ("sw", ("up", "dn", "dir", "net"))):
        for k in keys:
            np.testing.assert_allclose(sc[grp][k], 2 * base[grp][k], rtol=rt)
    np.testing.assert_allclose(sc["net"], 2 * base["net"], rtol=rt)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_heating_rate_is_net_flux_divergence(dtype):
    """:307-320"""
    out = oracle.solve_gray(dtype, nlay=20, ncol=3)
    p = oracle.GRAY_PARAMS
    cp = p["gas_constant"] / p["molmass_dryair"] / p["kappa_d"]
    F, pl = out["net"].astype(np.float64), out["state"]["p_lev"].astype(np.float64)
    expected = (p["grav"] / cp) * (F[:, 1:] - F[:, :-1]) / (pl[:, 1:] - pl[:, :-1])
    np.testing.assert_allclose(out["heating_rate"], expected, rtol=2e-5 if dtype == np.float32 else 1e-12)


# ------------------------------------------------------------------------------------------
# test/float32_consistency.jl:67-77,209-212 -- gray F32 vs F64 ratchet
# ------------------------------------------------------------------------------------------
def test_gray_float32_consistency():
    o32 = oracle.solve_gray(np.float32, nlay=60, ncol=8)
    o64 = oracle.solve_gray(np.float64, nlay=60, ncol=8)
    d = lambda a, b: np.abs(a.astype(np.float64) - b).max()
    assert max(d(o32["lw"]["net"], o64["lw"]["net"]), d(o32["sw"]["net"], o64["sw"]["net"])) <= 1e-3
    assert d(o32["heating_rate"], o64["heating_rate"]) <= 1e-8


def test_operation_count_build_of_the_oracle(tmp_path):
    """tools/opcount.py (SURVEY.md §8d "pin by op-counting the oracle"): the counting build runs the same source, and
    its counters show the structure of the reference algorithm -- one `log` per (layer, g-point) cell (the pressure
    interpolation is redone for every g-point, gas_optics.jl:100-115), the longwave two-stream coefficients evaluated
    twice per cell (longwave_2stream.jl:279,313: two `exp`, two `expm1`)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "ops.json"
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "opcount.py"), "--ncol", "2", "--nlay", "8", "--out", str(out)],
                          stdout=subprocess.DEVNULL)
    per = json.load(open(out))["per_column"]
    cells_lw, cells_sw = 8 * 256, 8 * 224
    assert per["lw"]["log"] == cells_lw and per["sw"]["log"] == cells_sw
    assert per["lw"]["exp"] == 2 * cells_lw and per["lw"]["expm1"] == 2 * cells_lw
    assert per["lw"]["flops_add_mul_div"] > 100 * cells_lw and per["sw"]["flops_add_mul_div"] > 100 * cells_sw

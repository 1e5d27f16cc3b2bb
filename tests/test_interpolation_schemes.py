"""Level <- layer interpolation and bottom extrapolation: the reference's closed-form unit tests
(test/interpolation_schemes.jl:35-333, test/grid_adaptation.jl:85-120) restated against the oracle's
`interpolate_levels` (src/api/interpolation.jl:176-252, src/api/grid_adaptation.jl:87-113).  The engine's
kernel is checked against the oracle in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import oracle
from oracle import GRAY_PARAMS as PRM

G = PRM["grav"]
R_D = PRM["gas_constant"] / PRM["molmass_dryair"]
CP_D = R_D / PRM["kappa_d"]


def _three_layer(dt, p, t, z=None):
    """Columns of 2 or 3 layers so that face 1 (0-based) is interp!(below = layer 0, above = layer 1)."""
    return np.asarray(p, dtype=dt)[None, :], np.asarray(t, dtype=dt)[None, :]


def _interp(dt, scheme, pd, Td, pu, Tu, **kw):
    # face index 1 of a 2-layer column is the interior face between the two layers
    p_lay, t_lay = _three_layer(dt, [pd, pu], [Td, Tu])
    p_lev, t_lev = oracle.interpolate_levels(p_lay, t_lay, np.array([300.0], dtype=dt), scheme, **kw)
    return p_lev[0, 1], t_lev[0, 1]


def _extrap_bottom(dt, scheme, bottom, p1, T1, p2, T2, Ts, **kw):
    p_lay, t_lay = _three_layer(dt, [p1, p2], [T1, T2])
    p_lev, t_lev = oracle.interpolate_levels(p_lay, t_lay, np.array([Ts], dtype=dt), scheme, bottom, **kw)
    return p_lev[0, 0], t_lev[0, 0]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_interp_closed_forms(dt):
    """test/interpolation_schemes.jl:35-98"""
    rt = 1e-5 if dt == np.float32 else 1e-12
    for pd, Td, pu, Tu in ((1.0e5, 300.0, 8.0e4, 280.0), (9.0e4, 290.0, 7.0e4, 270.0)):
        p, T = _interp(dt, "uniform_z", pd, Td, pu, Tu)
        assert T == pytest.approx((Td + Tu) / 2, rel=rt)
        assert min(pd, pu) < p < max(pd, pu) and min(Td, Tu) < T < max(Td, Tu)
        p, T = _interp(dt, "uniform_z", pd, Td, pu, Td)           # isothermal limit: geometric mean of p
        assert T == pytest.approx(Td, rel=rt) and p == pytest.approx(np.sqrt(pd * pu), rel=rt)
        p, T = _interp(dt, "uniform_p", pd, Td, pu, Tu)
        assert p == pytest.approx((pd + pu) / 2, rel=rt) and min(Td, Tu) < T < max(Td, Tu)
        p, T = _interp(dt, "uniform_p", pd, Td, pu, Td)
        assert T == pytest.approx(Td, rel=rt) and p == pytest.approx((pd + pu) / 2, rel=rt)
        p, T = _interp(dt, "arithmetic_mean", pd, Td, pu, Tu)
        assert p == pytest.approx((pd + pu) / 2, rel=rt) and T == pytest.approx((Td + Tu) / 2, rel=rt)
        p, T = _interp(dt, "geometric_mean", pd, Td, pu, Tu)
        assert p == pytest.approx(np.sqrt(pd * pu), rel=rt) and T == pytest.approx(np.sqrt(Td * Tu), rel=rt)
    # BestFit: T linear in z, p between the layers
    zc = np.array([[0.0, 1000.0]], dtype=dt); zf = np.array([[-500.0, 500.0, 1500.0]], dtype=dt)
    p, T = _interp(dt, "best_fit", 1.0e5, 300.0, 8.0e4, 280.0, center_z=zc, face_z=zf)
    assert T == pytest.approx(290.0, rel=rt) and 8.0e4 < p < 1.0e5


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_extrap_closed_forms(dt):
    """test/interpolation_schemes.jl:100-259"""
    rt = 2e-5 if dt == np.float32 else 1e-12
    p1, T1, p2, T2, Ts = 9.0e4, 285.0, 8.0e4, 275.0, 300.0
    p, T = _extrap_bottom(dt, "geometric_mean", "same_as_interpolation", p1, T1, p2, T2, Ts)
    assert T / T1 == pytest.approx(np.sqrt(T1 / T2), rel=rt) and p / p1 == pytest.approx(np.sqrt(p1 / p2), rel=rt)
    assert T > T1 and p > p1
    p, T = _extrap_bottom(dt, "uniform_z", "same_as_interpolation", p1, T1, p2, T2, Ts)
    assert T == pytest.approx((3 * T1 - T2) / 2, rel=rt) and p > p1
    p, T = _extrap_bottom(dt, "uniform_z", "same_as_interpolation", p1, T1, p2, T1, Ts)
    assert T == pytest.approx(T1, rel=rt) and p == pytest.approx(np.sqrt(p1 * p2), rel=rt)
    p, T = _extrap_bottom(dt, "uniform_p", "same_as_interpolation", p1, T1, p2, T2, Ts)
    assert p == pytest.approx((3 * p1 - p2) / 2, rel=rt) and T > T1
    p, T = _extrap_bottom(dt, "uniform_p", "same_as_interpolation", p1, T1, p2, T1, Ts)
    assert T == pytest.approx(T1, rel=rt) and p == pytest.approx((3 * p1 - p2) / 2, rel=rt)
    p, T = _extrap_bottom(dt, "arithmetic_mean", "use_surface_temp_at_bottom", p1, T1, p2, T2, Ts)
    assert T == dt(Ts) and p > p1
    assert p == pytest.approx(p1 * (Ts / T1) ** (CP_D / R_D), rel=rt)
    p, T = _extrap_bottom(dt, "arithmetic_mean", "use_surface_temp_at_bottom", p1, T1, p2, T2, T1)
    assert p == pytest.approx(p1, rel=rt)
    zc = np.array([[500.0, 1500.0]], dtype=dt); zf = np.array([[0.0, 1000.0, 2000.0]], dtype=dt)
    p, T = _extrap_bottom(dt, "arithmetic_mean", "hydrostatic_bottom", p1, T1, p2, T2, Ts, center_z=zc, face_z=zf)
    assert T == pytest.approx(T1 + (G / CP_D) * 500.0, rel=rt) and p > p1
    zf0 = np.array([[500.0, 1000.0, 2000.0]], dtype=dt)   # identity limit: z = z+
    p, T = _extrap_bottom(dt, "arithmetic_mean", "hydrostatic_bottom", p1, T1, p2, T2, Ts, center_z=zc, face_z=zf0)
    assert T == pytest.approx(T1, rel=rt) and p == pytest.approx(p1, rel=rt)
    p, T = _extrap_bottom(dt, "best_fit", "same_as_interpolation", p1, T1, p2, T2, Ts, center_z=zc, face_z=zf)
    assert T == pytest.approx(T1 + (T2 - T1) * (0.0 - 500.0) / 1000.0, rel=rt) and p > p1


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_dry_adiabat_is_reproduced_by_best_fit_and_hydrostatic_bottom(dt):
    """test/interpolation_schemes.jl:261-305: BestFit + HydrostaticBottom are exact on a dry adiabat."""
    T0, p0 = 300.0, 1.0e5
    T_ad = lambda z: T0 - (G / CP_D) * z
    p_ad = lambda z: p0 * (T_ad(z) / T0) ** (CP_D / R_D)
    nlay, ncol = 8, 2
    z_lev = np.linspace(0.0, 8000.0, nlay + 1)
    z_lay = 0.5 * (z_lev[:-1] + z_lev[1:])
    face_z = np.tile(z_lev, (ncol, 1)).astype(dt)
    center_z = np.tile(z_lay, (ncol, 1)).astype(dt)
    p_lev, t_lev = oracle.interpolate_levels(p_ad(center_z).astype(dt), T_ad(center_z).astype(dt),
                                             np.full(ncol, 300.0, dtype=dt), "best_fit", "hydrostatic_bottom",
                                             center_z=center_z, face_z=face_z)
    rtol = float(np.sqrt(np.finfo(dt).eps))
    np.testing.assert_allclose(t_lev, T_ad(face_z), rtol=rtol)
    np.testing.assert_allclose(p_lev, p_ad(face_z), rtol=rtol)


def test_uniform_p_needs_distinct_layer_pressures():
    """test/interpolation_schemes.jl:307-333: equal pressures give NaN temperatures, not silently wrong values."""
    with np.errstate(all="ignore"):
        p, T = _interp(np.float64, "uniform_p", 500.0, 250.0, 500.0, 260.0)
    assert p == 500.0 and np.isnan(T)


def test_arithmetic_mean_whole_column_and_boundary_layer_domain():
    """test/grid_adaptation.jl:85-120: interior faces are layer means, boundary faces the linear continuation;
    with an isothermal boundary layer only the domain faces are written."""
    rng = np.random.default_rng(0)
    ncol, nlay = 3, 6
    p_lay = np.sort(rng.uniform(1e3, 1e5, (ncol, nlay)))[:, ::-1].copy()
    t_lay = rng.uniform(200, 300, (ncol, nlay))
    p_lev, t_lev = oracle.interpolate_levels(p_lay, t_lay, np.full(ncol, 290.0), "arithmetic_mean")
    np.testing.assert_allclose(p_lev[:, 1:-1], 0.5 * (p_lay[:, :-1] + p_lay[:, 1:]))
    np.testing.assert_allclose(t_lev[:, 0], 1.5 * t_lay[:, 0] - 0.5 * t_lay[:, 1])
    np.testing.assert_allclose(p_lev[:, -1], 1.5 * p_lay[:, -1] - 0.5 * p_lay[:, -2])
    sentinel = np.full((ncol, nlay + 1), -1.0)
    p2, t2 = oracle.interpolate_levels(p_lay, t_lay, np.full(ncol, 290.0), "arithmetic_mean", nlay=nlay - 1,
                                       p_lev=sentinel.copy(), t_lev=sentinel.copy())
    assert (p2[:, -1] == -1.0).all() and (t2[:, -1] == -1.0).all()
    np.testing.assert_allclose(p2[:, nlay - 1], 1.5 * p_lay[:, nlay - 2] - 0.5 * p_lay[:, nlay - 3])
    # NoInterpolation leaves the levels alone
    p3, _ = oracle.interpolate_levels(p_lay, t_lay, np.full(ncol, 290.0), "none", p_lev=sentinel.copy(), t_lev=sentinel.copy())
    assert (p3 == -1.0).all()

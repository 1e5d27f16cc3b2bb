"""Seeded synthetic inputs: structurally faithful lookup tables and all-sky columns.

Neither the rrtmgp-data artifact nor a NetCDF reader exists in this image
(SURVEY.md §0), so benchmarks and parity tests run on *synthetic* tables that keep
the structure of the real ones -- two-atmosphere pressure axis with the duplicated
tropopause node, key-species pairs including `(0,0)->(2,2)` and dry-air members,
minor-gas CSR lists with every `scales_with_density` / `scale_by_complement` /
scaling-gas combination, Planck fractions and solar fractions that sum to one,
non-uniform aerosol RH levels -- with magnitudes chosen so that fluxes land at
O(100) W/m2 and the reference's W/m2 tolerances mean something.

Array shapes follow the reference's Julia dimension order (first index fastest):
`ext/lookup_constructors.jl:186-190,299-311,653-665,727-751,18-81`.

Atmospheric columns follow the reference's all-sky-with-aerosols test recipe
(`test/read_all_sky_with_aerosols.jl:44-48,76-81,84-102,138-153,186`) on top of the
analytic `standard_atmosphere` profiles (`src/api/atmosphere_profile.jl:44-97`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .lutpack import pack_luts

# Canonical gas slots (`ext/lookup_constructors.jl:5-16`: h2o -> 1, o3 -> 3).
GAS_NAMES = ["h2o", "co2", "o3", "n2o", "co", "ch4", "o2", "n2", "ccl4", "cfc11", "cfc12",
             "cfc22", "hfc143a", "hfc125", "hfc23", "hfc32", "hfc134a", "cf4", "no2"]

LW_BAND_WN = [10, 250, 500, 630, 700, 820, 980, 1080, 1180, 1390, 1480, 1800, 2080, 2250,
              2390, 2680, 3250]
SW_BAND_WN = [(820, 2680), (2680, 3250), (3250, 4000), (4000, 4650), (4650, 5150),
              (5150, 6150), (6150, 7700), (7700, 8050), (8050, 12850), (12850, 16000),
              (16000, 22650), (22650, 29000), (29000, 38000), (38000, 50000)]


@dataclass
class LutDims:
    n_eta: int = 9
    n_p_ref: int = 59
    n_t_ref: int = 14
    n_t_plnk: int = 196
    ngas: int = 19
    n_bnd_lw: int = 16
    n_bnd_sw: int = 14
    gpt_per_bnd: int = 16
    nsize_liq: int = 20
    nsize_ice: int = 18
    nrghice: int = 3
    nbin: int = 5
    nrh: int = 36
    # minor absorbers per band in the lower atmosphere (real tables have up to ~6; more than 4 LW / 3 SW
    # exercises the second 128-bit slot group of the fast kernels)
    minor_lower_lw: int = 3
    minor_lower_sw: int = 2
    # optional explicit g-points per band (lists) for reduced-resolution style tables
    gpts_lw: Optional[List[int]] = None
    gpts_sw: Optional[List[int]] = None

    def lw_gpts(self):
        return self.gpts_lw or [self.gpt_per_bnd] * self.n_bnd_lw

    def sw_gpts(self):
        return self.gpts_sw or [self.gpt_per_bnd] * self.n_bnd_sw


REAL_DIMS = LutDims()
SMALL_DIMS = LutDims(n_bnd_lw=3, n_bnd_sw=3, nsize_liq=8, nsize_ice=7, nrh=9)


def _planck_band(t, wn1, wn2, n=400):
    """Band-integrated Planck radiance [W m-2 sr-1] between wavenumbers wn1..wn2 [cm-1]."""
    h, c, kb = 6.62607015e-34, 2.99792458e8, 1.380649e-23
    nu = np.linspace(wn1, wn2, n) * 100.0  # m-1
    x = h * c * nu[None, :] / (kb * np.asarray(t)[:, None])
    b = 2 * h * c * c * nu[None, :] ** 3 / np.expm1(x)
    return np.trapezoid(b, nu, axis=1)


def _band_tables(rng, d: LutDims, n_bnd, gpts, lw: bool):
    """Gas-optics tables shared by LW and SW."""
    n_gpt = int(sum(gpts))
    n_eta, n_p, n_t = d.n_eta, d.n_p_ref + 1, d.n_t_ref
    out = {}
    lims = np.zeros((2, n_bnd), dtype=np.int32)
    gpt2bnd = np.zeros(n_gpt, dtype=np.int32)
    g0 = 0
    for b in range(n_bnd):
        lims[0, b], lims[1, b] = g0 + 1, g0 + gpts[b]
        gpt2bnd[g0:g0 + gpts[b]] = b + 1
        g0 += gpts[b]
    out["bnd_lims_gpt"] = lims
    out["major_gpt2bnd"] = gpt2bnd

    # reference pressures / temperatures (docs/src/Optics.md:191-211)
    p_ref = 109663.0 * np.exp(-0.2 * np.arange(d.n_p_ref))
    t_ref = 160.0 + 15.0 * np.arange(n_t)
    out["p_ref"] = p_ref
    out["t_ref"] = t_ref
    i_trop = min(12, d.n_p_ref - 2)
    p_ref_tropo = p_ref[i_trop]
    # kmajor pressure axis: nodes 0..i_trop lower, then i_trop..end upper (duplicated node)
    p_nodes = np.concatenate([p_ref[: i_trop + 1], p_ref[i_trop:]])
    assert p_nodes.size == n_p

    # reference vmr (2, ngas+1, n_t): slot 1 = dry air
    vmr_ref = np.zeros((2, d.ngas + 1, n_t))
    vmr_ref[:, 0, :] = 1.0
    base = {1: 5e-3, 2: 3.6e-4, 3: 2e-7, 4: 3.1e-7, 5: 1.2e-7, 6: 1.7e-6, 7: 0.209, 8: 0.781}
    for ig in range(1, d.ngas + 1):
        v0 = base.get(ig, 1e-10 * ig)
        tfac = np.exp(0.02 * (t_ref - 250.0)) if ig == 1 else 1.0 + 0.001 * (t_ref - 250.0)
        vmr_ref[0, ig, :] = v0 * tfac
        vmr_ref[1, ig, :] = v0 * tfac * (0.02 if ig == 1 else (20.0 if ig == 3 else 1.0))
    out["vmr_ref"] = vmr_ref

    # key species (2, 2, n_bnd); includes dry-air members and a (0,0) pair rewritten to (2,2)
    lower_pairs = [(1, 2), (1, 2), (1, 0), (1, 3), (1, 4), (2, 2), (1, 6), (1, 2), (2, 0), (1, 7)]
    upper_pairs = [(1, 2), (2, 2), (3, 2), (3, 0), (2, 4), (2, 2), (6, 2), (3, 2), (2, 0), (7, 2)]
    ks = np.zeros((2, 2, n_bnd), dtype=np.int32)
    for b in range(n_bnd):
        ks[:, 0, b] = lower_pairs[b % len(lower_pairs)]
        ks[:, 1, b] = upper_pairs[(b * 3 + 1) % len(upper_pairs)]
    out["key_species"] = ks

    # k-distribution: sorted in g within a band, spanning ~5 decades
    kmajor = np.zeros((n_eta, n_p, n_t, n_gpt))
    eta = np.linspace(0.0, 1.0, n_eta)
    lp = np.log(p_nodes / 1.0e5)
    tt = (t_ref - 250.0) / 100.0
    g0 = 0
    for b in range(n_bnd):
        ng = gpts[b]
        k0 = 10.0 ** rng.uniform(-25.0, -22.6) if lw else 10.0 ** rng.uniform(-27.5, -25.5)
        span = rng.uniform(3.0, 5.0) if lw else rng.uniform(2.0, 4.5)
        a = rng.uniform(0.2, 0.9)       # pressure broadening exponent
        bt = rng.uniform(-0.8, 1.2)     # temperature sensitivity
        ce = rng.uniform(-0.8, 0.8)     # eta tilt
        for j in range(ng):
            gfrac = (j + 0.5) / ng
            amp = k0 * 10.0 ** (span * gfrac ** 1.5)
            pe = a * (1.0 - 0.7 * gfrac)
            f = (np.exp(pe * lp)[None, :, None] * np.exp(bt * tt)[None, None, :]
                 * (1.0 + ce * (eta - 0.5))[:, None, None])
            noise = 1.0 + 0.1 * rng.standard_normal((n_eta, n_p, n_t))
            kmajor[:, :, :, g0 + j] = amp * f * np.clip(noise, 0.5, 1.5)
        g0 += ng
    out["kmajor"] = kmajor

    # minor absorbers, CSR over bands (lookup_constructors.jl:207-311 post-reorder layout:
    # contributors ordered g-point-major, then absorber-in-band)
    def minor(tag, counts, templates):
        bnd_st = np.ones(n_bnd + 1, dtype=np.int32)
        gasdata = []
        for b in range(n_bnd):
            for i in range(counts[b]):
                gasdata.append(templates[(b * 2 + i) % len(templates)])
            bnd_st[b + 1] = bnd_st[b] + counts[b]
        gasdata = np.array(gasdata, dtype=np.int32).T.reshape(4, -1)
        gpt_st = np.ones(n_gpt + 1, dtype=np.int32)
        for g in range(n_gpt):
            gpt_st[g + 1] = gpt_st[g] + counts[gpt2bnd[g] - 1]
        n_contrib = int(gpt_st[-1] - 1)
        kmin = np.zeros((n_eta, n_t, max(n_contrib, 1)))
        for g in range(n_gpt):
            b = gpt2bnd[g] - 1
            gfrac = (g + 1 - lims[0, b] + 0.5) / gpts[b]
            for i in range(counts[b]):
                row = gasdata[:, bnd_st[b] - 1 + i]
                # magnitude from a target column optical depth: tau ~ kminor * (typical scaling)
                typ = lambda ig: base.get(int(ig), 1e-10 * int(ig))
                scal = typ(row[0]) * 2.0e25
                if row[2] == 1:
                    scal *= 2.0  # 0.01 * p / T
                    if row[1] > 0:
                        scal *= 1.0 if row[3] == 1 else typ(row[1])
                mag = 10.0 ** rng.uniform(-2.5, -0.3 if lw else -1.0) / scal
                f = mag * (0.2 + gfrac) * (1.0 + 0.5 * (eta - 0.5))[:, None] \
                    * np.exp(rng.uniform(-0.5, 0.5) * tt)[None, :]
                kmin[:, :, gpt_st[g] - 1 + i] = f * (1.0 + 0.05 * rng.standard_normal((n_eta, n_t)))
        out[f"minor_{tag}/bnd_st"] = bnd_st
        out[f"minor_{tag}/gpt_st"] = gpt_st
        out[f"minor_{tag}/gasdata"] = gasdata
        out[f"minor_{tag}/kminor"] = kmin

    # (idx_gas, idx_scaling_gas, scales_with_density, scale_by_complement)
    templ_lower = [(1, 1, 1, 0), (1, 1, 1, 1), (2, 0, 0, 0), (8, 8, 1, 0), (4, 0, 0, 0),
                   (5, 0, 0, 0), (7, 0, 1, 0), (10, 0, 0, 0), (3, 0, 0, 0), (6, 0, 0, 0),
                   (8, 1, 1, 1), (11, 0, 0, 0)]
    templ_upper = [(2, 0, 0, 0), (3, 0, 0, 0), (7, 8, 1, 0), (4, 0, 0, 0), (9, 0, 0, 0),
                   (7, 0, 1, 0), (5, 0, 0, 0)]
    if lw:
        cl = [d.minor_lower_lw] * n_bnd
        cu = [1 + (b % 2) for b in range(n_bnd)]
    else:
        cl = [d.minor_lower_sw] * n_bnd
        cu = [1] * n_bnd
    if n_bnd > 2:
        cu[-1] = 0  # one band without upper-atmosphere minors (n == 0 path)
    minor("lower", cl, templ_lower)
    minor("upper", cu, templ_upper)

    out["idx_h2o"] = np.array([1], dtype=np.int32)
    t_min, t_max = t_ref[0], t_ref[-1]
    out["params"] = np.array([p_ref_tropo, p_ref.min(), t_min, t_max, 0.0])
    return out


def make_lut_arrays(seed: int = 7, dims: LutDims = REAL_DIMS) -> Dict[str, np.ndarray]:
    """All tables as `name -> ndarray` with reference (Julia-order) shapes."""
    d = dims
    rng = np.random.default_rng(seed)
    arrays: Dict[str, np.ndarray] = {}

    # ---------------- longwave ----------------
    gl = d.lw_gpts()
    lw = _band_tables(rng, d, d.n_bnd_lw, gl, lw=True)
    n_gpt_lw = int(sum(gl))
    wn_edges = np.interp(np.linspace(0, 16, d.n_bnd_lw + 1), np.arange(17), LW_BAND_WN)
    lw["bnd_lims_wn"] = np.stack([wn_edges[:-1], wn_edges[1:]])
    t_planck = 160.0 + np.arange(d.n_t_plnk) * (195.0 / max(d.n_t_plnk - 1, 1))
    lw["t_planck"] = t_planck
    lw["tot_planck"] = np.stack(
        [_planck_band(t_planck, wn_edges[b], wn_edges[b + 1]) for b in range(d.n_bnd_lw)], axis=1)
    pf = np.zeros((d.n_eta, d.n_p_ref + 1, d.n_t_ref, n_gpt_lw))
    g0 = 0
    for b in range(d.n_bnd_lw):
        ng = gl[b]
        w = np.sin(np.pi * (np.arange(ng) + 0.5) / (2 * ng)) ** 0.5  # quadrature-like weights
        w = w[::-1] * rng.uniform(0.8, 1.2, ng)
        f = w[None, None, None, :] * (1.0 + 0.15 * rng.standard_normal(
            (d.n_eta, d.n_p_ref + 1, d.n_t_ref, ng))).clip(0.4, 1.6)
        pf[..., g0:g0 + ng] = f / f.sum(axis=3, keepdims=True)
        g0 += ng
    lw["planck_fraction"] = pf
    for k, v in lw.items():
        arrays[f"lw/{k}"] = v

    # ---------------- shortwave ----------------
    gs = d.sw_gpts()
    sw = _band_tables(rng, d, d.n_bnd_sw, gs, lw=False)
    n_gpt_sw = int(sum(gs))
    idx = np.linspace(0, len(SW_BAND_WN) - 1, d.n_bnd_sw).round().astype(int)
    sw_wn = np.array([SW_BAND_WN[i] for i in idx], dtype=float).T  # (2, n_bnd)
    if d.n_bnd_sw < len(SW_BAND_WN):
        # keep a band containing 550 nm (18182 cm-1)
        sw_wn[:, d.n_bnd_sw // 2] = (16000.0, 22650.0)
    sw["bnd_lims_wn"] = sw_wn
    solar_band = np.array([_planck_band(np.array([5772.0]), a, b, 2000)[0] for a, b in sw_wn.T])
    solar_band = solar_band / solar_band.sum()
    ssrc = np.zeros(n_gpt_sw)
    rayl_lo = np.zeros((d.n_eta, d.n_t_ref, n_gpt_sw))
    rayl_up = np.zeros_like(rayl_lo)
    g0 = 0
    for b in range(d.n_bnd_sw):
        ng = gs[b]
        w = rng.uniform(0.5, 1.5, ng)
        ssrc[g0:g0 + ng] = solar_band[b] * w / w.sum()
        wn_mid = 0.5 * (sw_wn[0, b] + sw_wn[1, b])
        r0 = 4.4e-27 * (wn_mid / 18182.0) ** 4
        for j in range(ng):
            pert = 1.0 + 0.02 * rng.standard_normal((d.n_eta, d.n_t_ref))
            rayl_lo[:, :, g0 + j] = r0 * pert
            rayl_up[:, :, g0 + j] = r0 * (1.0 + 0.02 * rng.standard_normal((d.n_eta, d.n_t_ref)))
        g0 += ng
    sw["rayl_lower"], sw["rayl_upper"] = rayl_lo, rayl_up
    solar_tot = 1360.85
    sw["solar_src_scaled"] = ssrc / ssrc.sum()
    sw["params"][4] = solar_tot
    for k, v in sw.items():
        arrays[f"sw/{k}"] = v

    # ---------------- clouds (LookUpCld: lookup_constructors.jl:727-751) ----------------
    def cloud(tag, nbnd, wn, is_sw):
        radliq = (2.5, 21.5)
        radice = (5.0, 90.0)  # radii (diameters halved at load)
        rl = np.linspace(*radliq, d.nsize_liq)
        ri = np.linspace(*radice, d.nsize_ice)
        liq = np.zeros((3 * d.nsize_liq, nbnd))
        ice = np.zeros((3 * d.nsize_ice, nbnd, d.nrghice))
        for b in range(nbnd):
            x = b / max(nbnd - 1, 1)
            ext_l = 1.5 / rl * (1.0 + 0.1 * np.sin(3 * x + rl / 7.0))
            ext_i = 1.64 / ri * (1.0 + 0.1 * np.cos(2 * x + ri / 30.0))
            if is_sw:
                ssa_l = 1.0 - 10.0 ** (-5.0 + 4.2 * (1 - x)) * (1 + rl / 20.0)
                ssa_i = 1.0 - 10.0 ** (-4.5 + 3.8 * (1 - x)) * (1 + ri / 60.0)
                asy_l = 0.80 + 0.07 * rl / 21.5 + 0.02 * x
                asy_i = 0.74 + 0.12 * ri / 90.0 + 0.03 * x
            else:
                ssa_l = 0.45 + 0.3 * x + 0.1 * rl / 21.5
                ssa_i = 0.40 + 0.3 * x + 0.15 * ri / 90.0
                asy_l = 0.75 + 0.15 * rl / 21.5
                asy_i = 0.70 + 0.2 * ri / 90.0
            liq[:, b] = np.concatenate([ext_l, np.clip(ssa_l, 0, 1), np.clip(asy_l, 0, 0.98)])
            for r in range(d.nrghice):
                ice[:, b, r] = np.concatenate([ext_i * (1 + 0.03 * r), np.clip(ssa_i, 0, 1),
                                               np.clip(asy_i - 0.03 * r, 0, 0.98)])
        arrays[f"{tag}/dims"] = np.array([nbnd, d.nrghice, d.nsize_liq, d.nsize_ice, 2], dtype=np.int32)
        arrays[f"{tag}/bounds"] = np.array([radliq[0], radliq[1], radice[0], radice[1]])
        arrays[f"{tag}/liqdata"] = liq
        arrays[f"{tag}/icedata"] = ice
        arrays[f"{tag}/bnd_lims_wn"] = wn

    cloud("cld_lw", d.n_bnd_lw, lw["bnd_lims_wn"], False)
    cloud("cld_sw", d.n_bnd_sw, sw["bnd_lims_wn"], True)

    # ---------------- MERRA aerosols (lookup_constructors.jl:18-81) ----------------
    def aerosol(tag, nbnd, wn, is_sw):
        nbin, nrh = d.nbin, d.nrh
        lims = np.array([[0.1, 1.0], [1.0, 1.8], [1.8, 3.0], [3.0, 6.0], [6.0, 10.0]][:nbin]).T
        u = np.linspace(0.0, 1.0, nrh)
        rh = 0.99 * (1.0 - (1.0 - u) ** 1.8)  # non-uniform, denser near saturation, in [0, 0.99]
        rh[0] = 0.0
        growth = 1.0 + 2.5 * rh ** 3

        def props(ext0, ssa0, asy0, shape):
            e = ext0 * (1.0 + 0.2 * rng.standard_normal(shape)).clip(0.5, 1.5)
            s = np.clip(ssa0 + 0.03 * rng.standard_normal(shape), 0.02, 0.999)
            g = np.clip(asy0 + 0.03 * rng.standard_normal(shape), 0.05, 0.9)
            return e, s, g

        x = np.arange(nbnd) / max(nbnd - 1, 1)
        spec = (0.3 + 1.4 * x) if is_sw else (0.6 - 0.4 * x)  # ext larger toward visible
        ssab = (0.55 + 0.4 * x) if is_sw else (0.15 + 0.25 * x)
        dust = np.zeros((3, nbin, nbnd))
        salt = np.zeros((3, nrh, nbin, nbnd))
        for b in range(nbnd):
            for ib in range(nbin):
                e, s, g = props(2.5e2 * spec[b] / (1 + ib), ssab[b] - 0.03 * ib, 0.65 + 0.03 * ib, ())
                dust[:, ib, b] = (e, s, g)
                e, s, g = props(3.0e2 * spec[b] / (1 + ib) * growth, min(ssab[b] + 0.25, 0.99),
                                0.7 + 0.05 * rh, (nrh,))
                salt[:, :, ib, b] = np.stack([e, s, g])

        def rh_tbl(ext0, ssa_off):
            t = np.zeros((3, nrh, nbnd))
            for b in range(nbnd):
                e, s, g = props(ext0 * spec[b] * growth, np.clip(ssab[b] + ssa_off, 0.05, 0.99),
                                0.6 + 0.1 * rh, (nrh,))
                t[:, :, b] = np.stack([e, s, g])
            return t

        def dry_tbl(ext0, ssa_off):
            t = np.zeros((3, nbnd))
            for b in range(nbnd):
                e, s, g = props(ext0 * spec[b], np.clip(ssab[b] + ssa_off, 0.05, 0.99), 0.55, ())
                t[:, b] = (e, s, g)
            return t

        arrays[f"{tag}/dims"] = np.array([nbnd, 3, nbin, nrh, 2], dtype=np.int32)
        arrays[f"{tag}/size_bin_limits"] = lims
        arrays[f"{tag}/rh_levels"] = rh
        arrays[f"{tag}/dust"] = dust
        arrays[f"{tag}/sea_salt"] = salt
        arrays[f"{tag}/sulfate"] = rh_tbl(4.0e2, 0.3)
        arrays[f"{tag}/black_carbon_rh"] = rh_tbl(9.0e2, -0.35)
        arrays[f"{tag}/black_carbon"] = dry_tbl(8.0e2, -0.4)
        arrays[f"{tag}/organic_carbon_rh"] = rh_tbl(5.0e2, 0.2)
        arrays[f"{tag}/organic_carbon"] = dry_tbl(4.0e2, 0.15)
        arrays[f"{tag}/bnd_lims_wn"] = wn
        i550 = 0
        for b in range(nbnd):  # lookup_constructors.jl:41-44
            if 1.0 / (wn[1, b] * 100.0) <= 550e-9 <= 1.0 / (wn[0, b] * 100.0):
                i550 = b + 1
                break
        arrays[f"{tag}/iband_550nm"] = np.array([i550], dtype=np.int32)

    aerosol("aero_lw", d.n_bnd_lw, lw["bnd_lims_wn"], False)
    aerosol("aero_sw", d.n_bnd_sw, sw["bnd_lims_wn"], True)
    return arrays


def make_lut_pack(seed: int = 7, dims: LutDims = REAL_DIMS) -> bytes:
    return pack_luts(make_lut_arrays(seed, dims))


# ------------------------------------------------------------------------------------
# atmospheric columns
# ------------------------------------------------------------------------------------
_ATMOS = [  # atmosphere_profile.jl:44-69 (tropical, midlatitude summer, subarctic winter)
    dict(t_sfc=300.0, z_trop=17.0e3, g_trop=6.5e-3, g_strat=2.2e-3, h2o=2.3e-2, lat=0.0),
    dict(t_sfc=294.0, z_trop=13.0e3, g_trop=6.5e-3, g_strat=2.0e-3, h2o=1.4e-2, lat=45.0),
    dict(t_sfc=257.0, z_trop=9.0e3, g_trop=5.0e-3, g_strat=1.5e-3, h2o=1.6e-3, lat=65.0),
]

# test overrides (test/all_sky_with_aerosols_utils.jl:41-43) on the defaults (standalone.jl:87-97)
PARAMS = dict(grav=9.80665, molmass_dryair=0.028964, molmass_water=0.018016,
              gas_constant=8.314462618, kappa_d=2.0 / 7.0, Stefan=5.670374419e-8,
              avogad=6.02214076e23)


def _std_T(z, a):
    t_trop = a["t_sfc"] - a["g_trop"] * a["z_trop"]
    return np.where(z <= a["z_trop"], a["t_sfc"] - a["g_trop"] * z, t_trop + a["g_strat"] * (z - a["z_trop"]))


def _std_p(z, a, p_sfc, grav, r_d):
    t_trop = a["t_sfc"] - a["g_trop"] * a["z_trop"]
    t = _std_T(z, a)
    e1 = grav / (r_d * a["g_trop"])
    p_low = p_sfc * (np.minimum(t, a["t_sfc"]) / a["t_sfc"]) ** e1
    p_trop = p_sfc * (t_trop / a["t_sfc"]) ** e1
    p_up = p_trop * (np.maximum(t, t_trop) / t_trop) ** (-grav / (r_d * a["g_strat"]))
    return np.where(z <= a["z_trop"], p_low, p_up)


def relative_humidity(p_lay, t_lay, vmr_h2o, params=PARAMS):
    """`compute_relative_humidity_kernel!` (src/optics/gas_optics.jl:58-80); a host duty."""
    dt = p_lay.dtype.type
    mwd = dt(dt(params["molmass_water"]) / dt(params["molmass_dryair"]))
    mmr = vmr_h2o * mwd
    q = mmr / (dt(1) + mmr)
    q = np.maximum(dt(1e-7), q)
    es = np.exp((dt(17.67) * (t_lay - dt(273.16))) / (t_lay - dt(29.65)))
    return np.maximum(dt(0.01) * (dt(0.263) * p_lay * q) / es, dt(0)).astype(p_lay.dtype)


def make_atmosphere(ncol: int, nlay: int = 64, *, seed: int = 20260101, dtype=np.float32,
                    n_bnd_lw: int = 16, n_bnd_sw: int = 14, ngas: int = 19,
                    cloud_bounds=(2.5, 21.5, 5.0, 90.0), solar_src_tot: float = 1360.85,
                    cld_frac: Optional[float] = 1.0, clouds: bool = True, aerosols: bool = True,
                    cos_zenith: Optional[float] = 0.86, z_top: float = 45.0e3,
                    vmr_kind: str = "gm", with_lat: bool = False) -> Dict[str, np.ndarray]:
    """Synthetic all-sky-with-aerosols columns. Arrays are C-ordered `[ncol][vertical]`
    views of the reference's `(vertical, ncol)` Julia arrays.

    `cld_frac=None` draws U(0,1) cloud fractions (McICA variant); `cos_zenith=None`
    draws U(-0.2, 1) (about 17 % night columns)."""
    rng = np.random.default_rng(seed)
    nlev = nlay + 1
    grav, r_d = PARAMS["grav"], PARAMS["gas_constant"] / PARAMS["molmass_dryair"]
    z_lev = np.linspace(0.0, z_top, nlev)
    z_lay = 0.5 * (z_lev[:-1] + z_lev[1:])
    kind = np.arange(ncol) % 3
    p_sfc = 101325.0 * rng.uniform(0.95, 1.02, ncol)
    p_lev = np.zeros((ncol, nlev))
    p_lay = np.zeros((ncol, nlay))
    t_lev = np.zeros((ncol, nlev))
    t_lay = np.zeros((ncol, nlay))
    h2o = np.zeros((ncol, nlay))
    lat = np.zeros(ncol)
    for k, a in enumerate(_ATMOS):
        m = kind == k
        if not m.any():
            continue
        ps = p_sfc[m][:, None]
        p_lev[m] = _std_p(z_lev[None, :], a, ps, grav, r_d)
        p_lay[m] = _std_p(z_lay[None, :], a, ps, grav, r_d)
        t_lev[m] = _std_T(z_lev, a)[None, :]
        t_lay[m] = _std_T(z_lay, a)[None, :]
        h2o[m] = np.maximum(a["h2o"] * np.exp(-z_lay / 2.0e3), 4.0e-6)[None, :]
        lat[m] = a["lat"]
    # per-column perturbations
    dT = rng.normal(0.0, 3.0, ncol)[:, None]
    t_lev += dT
    t_lay += dT
    h2o *= np.exp(rng.normal(0.0, 0.3, ncol))[:, None]
    o3 = 3.0e-8 + 7.5e-6 * np.exp(-np.log(p_lay / 1.2e3) ** 2 / (2 * 1.2 ** 2))
    t_sfc = t_lev[:, 0] + rng.normal(0.0, 2.0, ncol)

    out: Dict[str, np.ndarray] = {}
    f = lambda a: np.ascontiguousarray(a, dtype=dtype)
    out["p_lev"], out["t_lev"], out["t_sfc"] = f(p_lev), f(t_lev), f(t_sfc)
    p_lay_f, t_lay_f, h2o_f = f(p_lay), f(t_lay), f(h2o)
    rh = relative_humidity(p_lay_f, t_lay_f, h2o_f)
    layerdata = np.zeros((ncol, nlay, 4), dtype=dtype)  # (col_dry, p_lay, t_lay, rel_hum)
    layerdata[:, :, 1], layerdata[:, :, 2], layerdata[:, :, 3] = p_lay_f, t_lay_f, rh
    out["layerdata"] = layerdata
    vmr = np.zeros(ngas, dtype=dtype)  # read_all_sky_with_aerosols.jl:76-81
    wm = {2: 348e-6, 4: 306e-9, 5: 0.0, 6: 1650e-9, 7: 0.2095, 8: 0.7808, 9: 1.0e-10,
          10: 2.5e-10, 11: 5.0e-10}
    for ig, v in wm.items():
        if ig <= ngas:
            vmr[ig - 1] = v
    if vmr_kind == "gm":
        out["vmr_h2o"], out["vmr_o3"], out["vmr"] = h2o_f, f(o3), vmr
    else:  # full per-gas storage Vmr (ngas, nlay, ncol) -> [ncol][nlay][ngas]
        full = np.zeros((ncol, nlay, ngas), dtype=dtype)
        full[:, :, :] = vmr[None, None, :]
        full[:, :, 0], full[:, :, 2] = h2o_f, f(o3)
        out["vmr_full"] = full
    if with_lat:
        out["lat"] = f(lat)

    # boundary conditions (read_all_sky_with_aerosols.jl:44-48)
    out["sfc_emis"] = np.full((ncol, n_bnd_lw), 0.98, dtype=dtype)
    out["sfc_alb_direct"] = np.full((ncol, n_bnd_sw), 0.06, dtype=dtype)
    out["sfc_alb_diffuse"] = np.full((ncol, n_bnd_sw), 0.06, dtype=dtype)
    mu0 = np.full(ncol, cos_zenith) if cos_zenith is not None else rng.uniform(-0.2, 1.0, ncol)
    out["cos_zenith"] = f(mu0)
    out["toa_flux"] = np.full(ncol, solar_src_tot, dtype=dtype)

    if clouds:  # read_all_sky_with_aerosols.jl:138-153
        icol1 = np.arange(1, ncol + 1)[:, None]
        cloudy = (p_lay_f > 10000) & (p_lay_f < 90000) & (icol1 % 3 != 0)
        frac = np.full((ncol, nlay), cld_frac) if cld_frac is not None else rng.uniform(0, 1, (ncol, nlay))
        r_liq = 0.5 * (cloud_bounds[0] + cloud_bounds[1])
        r_ice = 0.5 * (cloud_bounds[2] + cloud_bounds[3])
        liq = cloudy & (t_lay_f > 263)
        ice = cloudy & (t_lay_f < 273)
        out["cld_frac"] = f(np.where(cloudy, frac, 0.0))
        out["cld_path_liq"] = f(np.where(liq, 10.0, 0.0))
        out["cld_r_eff_liq"] = f(np.where(liq, r_liq, 0.0))
        out["cld_path_ice"] = f(np.where(ice, 10.0, 0.0))
        out["cld_r_eff_ice"] = f(np.where(ice, r_ice, 0.0))
    if aerosols:  # one species per (layer, column) (read_all_sky_with_aerosols.jl:84-102)
        typ = (np.arange(nlay)[None, :] + np.arange(ncol)[:, None]) % 15
        mass = 10.0 ** rng.uniform(-6.0, -4.0, (ncol, nlay))
        size = rng.uniform(0.1, 10.0, (ncol, nlay))
        am = np.zeros((ncol, nlay, 15), dtype=dtype)
        asz = np.zeros((ncol, nlay, 15), dtype=dtype)
        ii, jj = np.meshgrid(np.arange(ncol), np.arange(nlay), indexing="ij")
        am[ii, jj, typ] = mass
        asz[ii, jj, typ] = size
        out["aero_mass"], out["aero_size"] = am, asz
    return out

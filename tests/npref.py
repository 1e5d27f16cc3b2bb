"""Independent numpy restatement of the spectral hot path (Float64), written from the Julia sources
separately from oracle/rrtmgp_oracle.cpp, to catch transcription errors in the C++ oracle.

TEST INFRASTRUCTURE ONLY.  Vectorised over g-points; loops over columns and layers in Python, so
only for a handful of columns.  cld_frac must be 0 or 1 (deterministic McICA mask).
Arrays are the Julia-shaped tables of `synthetic.make_lut_arrays` (1-based integer tables).
"""
import numpy as np

EPS = np.finfo(np.float64).eps
K_MIN = np.sqrt(EPS)          # src/Numerics.jl:24
RES_WIN = np.sqrt(EPS)        # :49
MU0_MIN = EPS                 # :63


def _fix_key_species(ks):     # ext/lookup_constructors.jl:175-182
    ks = ks.copy()
    z = (ks[0] == 0) & (ks[1] == 0)
    ks[0][z] = 2
    ks[1][z] = 2
    return ks


def _get_vmr(st, ig, lay, col):   # src/optics/VolumeMixingRatios.jl:91-106 (VmrGM)
    if ig == 0:
        return 1.0
    if ig == 1:
        return st["vmr_h2o"][col, lay]
    if ig == 3:
        return st["vmr_o3"][col, lay]
    return st["vmr"][ig - 1]


def gas_optics_column(T, pre, st, col, sw):
    """src/optics/gas_optics.jl:176-320,344-444 for every layer and g-point of one column."""
    kmajor = T[f"{pre}/kmajor"]
    n_eta, n_p, n_t, n_gpt = kmajor.shape
    ks = _fix_key_species(T[f"{pre}/key_species"])
    g2b = T[f"{pre}/major_gpt2bnd"]
    t_ref, ln_p_ref = T[f"{pre}/t_ref"], np.log(T[f"{pre}/p_ref"])
    vmr_ref = T[f"{pre}/vmr_ref"]
    p_trop = T[f"{pre}/params"][0]
    nlay = st["layerdata"].shape[1]
    tau = np.zeros((nlay, n_gpt)); ssa = np.zeros((nlay, n_gpt)); pfrac = np.zeros((nlay, n_gpt))
    dT = t_ref[1] - t_ref[0]
    dlnp = ln_p_ref[0] - ln_p_ref[1]
    for lay in range(nlay):
        col_dry, p, t = st["layerdata"][col, lay, 0], st["layerdata"][col, lay, 1], st["layerdata"][col, lay, 2]
        tropo = 1 if p > p_trop else 2
        h2o = _get_vmr(st, 1, lay, col)
        # :87-93 with optics_utils.jl:7-14
        if t <= t_ref[0]:
            jt = 1
        elif t >= t_ref[-1]:
            jt = n_t - 1
        else:
            jt = min(int((t - t_ref[0]) / dT) + 1, n_t - 1)
        ft = (t - t_ref[jt - 1]) / dT
        # :100-115
        lp = np.log(p)
        jp = min(max(int((ln_p_ref[0] - lp) / dlnp) + 1, 1), ln_p_ref.size - 1) + 1
        fp = (ln_p_ref[jp - 2] - lp) / dlnp
        jpt = jp + tropo - 1
        minor = "minor_lower" if tropo == 1 else "minor_upper"
        bst, gst = T[f"{pre}/{minor}/bnd_st"], T[f"{pre}/{minor}/gpt_st"]
        gd, kmin = T[f"{pre}/{minor}/gasdata"], T[f"{pre}/{minor}/kminor"]
        for g in range(n_gpt):
            b = g2b[g]
            ig1, ig2 = ks[0, tropo - 1, b - 1], ks[1, tropo - 1, b - 1]
            v1, v2 = _get_vmr(st, ig1, lay, col), _get_vmr(st, ig2, lay, col)
            je, fe, cm = [], [], []
            for it in range(2):   # :129-170
                eta_half = vmr_ref[tropo - 1, ig1, jt - 1 + it] / vmr_ref[tropo - 1, ig2, jt - 1 + it]
                col_mix = v1 + eta_half * v2
                eta = v1 / col_mix if col_mix > 0 else 0.5
                loc = eta * (n_eta - 1)
                j = min(int(loc) + 1, n_eta - 1)
                je.append(j); fe.append(loc - (j - 1)); cm.append(col_mix)

            def interp3(tbl, s1=1.0, s2=1.0):   # optics_utils.jl:136-181
                c = tbl[:, :, :, g]
                a = (1 - fp) * ((1 - ft) * ((1 - fe[0]) * c[je[0] - 1, jpt - 2, jt - 1] + fe[0] * c[je[0], jpt - 2, jt - 1])) + \
                    fp * ((1 - ft) * ((1 - fe[0]) * c[je[0] - 1, jpt - 1, jt - 1] + fe[0] * c[je[0], jpt - 1, jt - 1]))
                bb = (1 - fp) * (ft * ((1 - fe[1]) * c[je[1] - 1, jpt - 2, jt] + fe[1] * c[je[1], jpt - 2, jt])) + \
                    fp * (ft * ((1 - fe[1]) * c[je[1] - 1, jpt - 1, jt] + fe[1] * c[je[1], jpt - 1, jt]))
                return s1 * a + s2 * bb

            def interp2(c):                     # optics_utils.jl:85-98
                return (1 - fe[0]) * (1 - ft) * c[je[0] - 1, jt - 1] + fe[0] * (1 - ft) * c[je[0], jt - 1] + \
                    (1 - fe[1]) * ft * c[je[1] - 1, jt] + fe[1] * ft * c[je[1], jt]

            tau_major = interp3(kmajor, cm[0], cm[1]) * col_dry
            tau_minor = 0.0   # :344-412
            n = gst[g + 1] - gst[g]
            for i in range(n):
                idx_gas, idx_sc, swd, sbc = gd[:, bst[b - 1] - 1 + i]
                vm = _get_vmr(st, idx_gas, lay, col)
                if vm > 0:
                    scaling = vm * col_dry
                    if swd == 1:
                        scaling *= 0.01 * p / t
                        if idx_sc > 0:
                            x = _get_vmr(st, idx_sc, lay, col) / (1 + h2o)
                            scaling *= (1 - x) if sbc == 1 else x
                    tau_minor += interp2(kmin[:, :, gst[g] - 1 + i]) * scaling
            if not sw:
                pfrac[lay, g] = interp3(T[f"{pre}/planck_fraction"])
                tau[lay, g] = max(tau_major + tau_minor, 0.0)
            else:
                ray = T[f"{pre}/rayl_lower"] if tropo == 1 else T[f"{pre}/rayl_upper"]
                tau_ray = interp2(ray[:, :, g]) * (h2o + 1) * col_dry
                tt = max(tau_major + tau_minor + tau_ray, 0.0)
                tau[lay, g] = tt
                ssa[lay, g] = tau_ray / tt if tt > 0 else 0.0
    return tau, ssa, pfrac


def _delta_scale(tau, ssa, g):   # optics_utils.jl:208-223
    s1g2 = ssa * (1 - g) * (1 + g)
    omwf = (1 - ssa) + s1g2
    return omwf * tau, s1g2 / max(EPS, omwf), g / max(EPS, 1 + g)


def _increment(t1, s1, g1, t2, s2, g2):   # optics_utils.jl:189-202 (vector t1.. over g-points of a band)
    tau = t1 + t2
    ssa = t1 * s1 + t2 * s2
    ssag = (t1 * s1 * g1 + t2 * s2 * g2) / np.maximum(EPS, ssa)
    return tau, ssa / np.maximum(EPS, tau), ssag


def _cld(n, lwr, upr, tbl, re, path):   # cloud_optics.jl:154-244; tbl = (3n,) ext|ssa|asy
    if not path > EPS:
        return 0.0, 0.0, 0.0
    dr = (upr - lwr) / (n - 1)
    re = max(min(re, upr), lwr)
    loc = max(min(int((re - lwr) / dr) + 1, n - 1), 1)
    fac = (re - lwr - (loc - 1) * dr) / dr
    lerp = lambda o: (1 - fac) * tbl[o + loc - 1] + fac * tbl[o + loc]
    t = max(lerp(0) * path, 0.0)
    ts = lerp(n) * t
    return t, ts, lerp(2 * n) * ts


def _aerosol(T, pre, ibnd, mass, size, rh):   # aerosol_optics.jl:141-451; ibnd 1-based
    lims, rhl = T[f"{pre}/size_bin_limits"], T[f"{pre}/rh_levels"]
    nbin = lims.shape[1]

    def sbin(sz):
        for ib in range(nbin):
            if lims[0, ib] <= sz <= lims[1, ib]:
                return ib
        return nbin - 1

    def rhw():   # optics_utils.jl:51-62
        if rh < rhl[0]:
            return 0, 0.0
        if rh > rhl[-1]:
            return rhl.size - 2, 1.0
        loc = rhl.size - 2
        if rh <= rhl[0]:
            loc = 0
        else:
            for i in range(rhl.size):
                if rh < rhl[i]:
                    loc = i - 1
                    break
        return loc, (rh - rhl[loc]) / (rhl[loc + 1] - rhl[loc])

    acc = np.zeros(3)

    def add(e, s, g, m):
        t = m * e
        acc[:] += (t, t * s, t * s * g)

    def add_rh(tbl3, m):   # tbl3 (3, nrh)
        loc, f = rhw()
        v = tbl3[:, loc] * (1 - f) + tbl3[:, loc + 1] * f
        add(v[0], v[1], v[2], m)

    for i in (1, 8, 9, 10, 11):
        if mass[i - 1] > 0:
            add(*T[f"{pre}/dust"][:, sbin(size[i - 1]), ibnd - 1], mass[i - 1])
    for i in (2, 12, 13, 14, 15):
        if mass[i - 1] > 0:
            add_rh(T[f"{pre}/sea_salt"][:, :, sbin(size[i - 1]), ibnd - 1], mass[i - 1])
    if mass[2] > 0: add_rh(T[f"{pre}/sulfate"][:, :, ibnd - 1], mass[2])
    if mass[3] > 0: add_rh(T[f"{pre}/black_carbon_rh"][:, :, ibnd - 1], mass[3])
    if mass[4] > 0: add(*T[f"{pre}/black_carbon"][:, ibnd - 1], mass[4])
    if mass[5] > 0: add_rh(T[f"{pre}/organic_carbon_rh"][:, :, ibnd - 1], mass[5])
    if mass[6] > 0: add(*T[f"{pre}/organic_carbon"][:, ibnd - 1], mass[6])
    return acc


def add_clouds_aerosols(T, st, col, sw, tau, ssa, g, ice_rgh=2, clouds=True, aerosols=True):
    """compute_optical_props.jl:197-245 / :346-385: cloud then aerosol increments; returns AOD (SW)."""
    pre_c, pre_a, pre = ("cld_sw", "aero_sw", "sw") if sw else ("cld_lw", "aero_lw", "lw")
    g2b = T[f"{pre}/major_gpt2bnd"]
    nlay = tau.shape[0]
    nbnd = int(g2b.max())
    aod = np.zeros(2)
    for lay in range(nlay):
        for b in range(1, nbnd + 1):
            sel = g2b == b
            if clouds and st["cld_frac"][col, lay] > 0:
                d = T[f"{pre_c}/dims"]; bo = T[f"{pre_c}/bounds"]
                tl = _cld(d[2], bo[0], bo[1], T[f"{pre_c}/liqdata"][:, b - 1], st["cld_r_eff_liq"][col, lay], st["cld_path_liq"][col, lay])
                ti = _cld(d[3], bo[2], bo[3], T[f"{pre_c}/icedata"][:, b - 1, ice_rgh - 1], st["cld_r_eff_ice"][col, lay], st["cld_path_ice"][col, lay])
                tc = tl[0] + ti[0]; sc = tl[1] + ti[1]
                gc = (tl[2] + ti[2]) / max(EPS, sc); sc = sc / max(EPS, tc)
                if sw:
                    tc, sc, gc = _delta_scale(tc, sc, gc)
                tau[lay, sel], ssa[lay, sel], g[lay, sel] = _increment(tau[lay, sel], ssa[lay, sel], g[lay, sel], tc, sc, gc)
            if aerosols and (st["aero_mass"][col, lay] > 0).any():
                ta, tsa, tsga = _aerosol(T, pre_a, b, st["aero_mass"][col, lay], st["aero_size"][col, lay], st["layerdata"][col, lay, 3])
                if sw and b == T[f"{pre_a}/iband_550nm"][0]:
                    aod += (ta, tsa)
                ga = tsga / max(EPS, tsa); sa = tsa / max(EPS, ta)
                if sw:
                    ta, sa, ga = _delta_scale(ta, sa, ga)
                tau[lay, sel], ssa[lay, sel], g[lay, sel] = _increment(tau[lay, sel], ssa[lay, sel], g[lay, sel], ta, sa, ga)
    return aod


def lw_two_stream(T, st, col, tau, ssa, g, pfrac):
    """compute_optical_props.jl:157-195 sources + longwave_2stream.jl:149-334; returns broadband up, dn."""
    g2b = T["lw/major_gpt2bnd"]
    t_pl, tot = T["lw/t_planck"], T["lw/tot_planck"]
    nlay, n_gpt = tau.shape
    nlev = nlay + 1

    def B(t):   # optics_utils.jl:34-44 per band -> per g-point
        return np.array([np.interp(t, t_pl, tot[:, b - 1]) for b in g2b])

    t_lev, t_sfc = st["t_lev"][col], st["t_sfc"][col]
    lev_source = np.zeros((nlev, n_gpt))
    inc_prev = None
    for k in range(nlay):
        inc = B(t_lev[k + 1]) * pfrac[k]
        dec = B(t_lev[k]) * pfrac[k]
        lev_source[k] = dec if k == 0 else np.sqrt(inc_prev * dec)
        inc_prev = inc
    lev_source[nlay] = inc_prev
    sfc_source = B(t_sfc) * pfrac[0]
    emis = st["sfc_emis"][col][g2b - 1]
    R = np.zeros((nlay, n_gpt)); Tt = np.zeros_like(R); su = np.zeros_like(R); sd = np.zeros_like(R)
    D = 1.66
    for k in range(nlay):
        t, w, gg = tau[k], ssa[k], g[k]
        g1 = D * (1 - 0.5 * w * (1 + gg)); g2 = D * 0.5 * w * (1 - gg)
        kk = np.sqrt(np.maximum(D * (1 - w) * (g1 + g2), K_MIN))
        e1 = np.exp(-t * kk); om1 = -np.expm1(-t * kk)
        om2 = om1 * (1 + e1)
        RT = 1 / (kk * (1 + e1 * e1) + g1 * om2)
        R[k] = RT * g2 * om2; Tt[k] = RT * 2 * kk * e1
        bot, top = lev_source[k], lev_source[k + 1]
        dB = bot - top; gs = g1 + g2
        emis_fac = om1 * (kk * om1 + D * (1 - w) * (1 + e1)) * RT
        with np.errstate(divide="ignore", invalid="ignore"):
            dBz = dB * (om1 / t) * (kk * om1 + gs * (1 + e1)) * RT / np.maximum(gs, EPS)
        up_ = np.pi * (top * emis_fac - Tt[k] * dB + dBz); dn_ = np.pi * (bot * emis_fac + Tt[k] * dB - dBz)
        su[k] = np.where(t > 0, up_, 0.0); sd[k] = np.where(t > 0, dn_, 0.0)
    alb = np.zeros((nlev, n_gpt)); src = np.zeros_like(alb)
    alb[0] = 1 - emis; src[0] = np.pi * emis * sfc_source
    for k in range(nlay):
        den = 1 / (1 - R[k] * alb[k])
        alb[k + 1] = R[k] + Tt[k] ** 2 * alb[k] * den
        src[k + 1] = su[k] + Tt[k] * den * (src[k] + alb[k] * sd[k])
    dn = np.zeros((nlev, n_gpt)); up = np.zeros_like(dn)
    up[nlay] = dn[nlay] * alb[nlay] + src[nlay]
    for k in range(nlay - 1, -1, -1):
        den = 1 / (1 - R[k] * alb[k])
        dn[k] = (Tt[k] * dn[k + 1] + R[k] * src[k] + sd[k]) * den
        up[k] = dn[k] * alb[k] + src[k]
    return up.sum(1), dn.sum(1)


def sw_two_stream(T, st, col, tau, ssa, g):
    """shortwave_2stream.jl:189-392; returns broadband up, dn, dir."""
    g2b = T["sw/major_gpt2bnd"]
    nlay, n_gpt = tau.shape
    nlev = nlay + 1
    mu0 = st["cos_zenith"][col]
    if mu0 <= 0:
        z = np.zeros(nlev)
        return z, z.copy(), z.copy()
    top = st["toa_flux"][col] * T["sw/solar_src_scaled"] * mu0
    inv = 1 / max(mu0, MU0_MIN)
    dirf = np.zeros((nlev, n_gpt)); dirf[nlay] = top
    cum = np.zeros(n_gpt)
    for k in range(nlay - 1, -1, -1):
        cum = cum + tau[k]
        dirf[k] = top * np.exp(-cum * inv)
    Rdir = np.zeros((nlay, n_gpt)); Tdir = np.zeros_like(Rdir); R = np.zeros_like(Rdir); Tt = np.zeros_like(Rdir)
    for k in range(nlay):
        t, w, gg = tau[k], ssa[k], g[k]
        g1 = (8 - w * (5 + 3 * gg)) * 0.25; g2 = 3 * (w * (1 - gg)) * 0.25
        g3 = (2 - (3 * mu0) * gg) * 0.25; g4 = 1 - g3
        a1 = g1 * g4 + g2 * g3; a2 = g1 * g3 + g2 * g4
        kk = np.sqrt(np.maximum(2 * (1 - w) * (g1 + g2), K_MIN))
        e = np.exp(-t * kk); e2 = e * e; om1 = -np.expm1(-t * kk); om2 = om1 * (1 + e)
        RT = 1 / (kk * (1 + e2) + g1 * om2)
        R[k] = RT * g2 * om2; Tt[k] = RT * 2 * kk * e
        T0 = np.exp(-t / max(mu0, MU0_MIN))
        kmu = kk * mu0; kmu2 = kmu * kmu; diff = 1 - kmu2
        res = np.abs(diff) < RES_WIN
        kmu2 = np.where(res, np.where(diff >= 0, 1 - RES_WIN, 1 + RES_WIN), kmu2)
        kmu = np.where(res, np.sqrt(kmu2), kmu)
        kg3, kg4 = kk * g3, kk * g4
        RT2 = w * RT / (1 - kmu2)
        Ru = RT2 * ((1 - kmu) * (a2 + kg3) - (1 + kmu) * (a2 - kg3) * e2 - 2 * (kg3 - a2 * kmu) * e * T0)
        Tu = -RT2 * ((1 + kmu) * (a1 + kg4) * T0 - (1 - kmu) * (a1 - kg4) * e2 * T0 - 2 * (kg4 + a1 * kmu) * e)
        Rd, Td = np.maximum(0, Ru), np.maximum(0, Tu)
        av = np.maximum(0, 1 - T0); totd = Rd + Td
        sc = np.where(totd > av, av / np.maximum(EPS, totd), 1.0)
        Rdir[k], Tdir[k] = Rd * sc, Td * sc
    adir = st["sfc_alb_direct"][col][g2b - 1]; adif = st["sfc_alb_diffuse"][col][g2b - 1]
    alb = np.zeros((nlev, n_gpt)); src = np.zeros_like(alb)
    alb[0] = adif; src[0] = dirf[0] * adir
    for k in range(nlay):
        den = 1 / (1 - R[k] * alb[k])
        alb[k + 1] = R[k] + Tt[k] ** 2 * alb[k] * den
        src[k + 1] = Rdir[k] * dirf[k + 1] + Tt[k] * den * (src[k] + alb[k] * Tdir[k] * dirf[k + 1])
    dn = np.zeros((nlev, n_gpt)); up = np.zeros_like(dn)
    up[nlay] = src[nlay]
    for k in range(nlay - 1, -1, -1):
        den = 1 / (1 - R[k] * alb[k])
        dn[k] = (Tt[k] * dn[k + 1] + R[k] * src[k] + Tdir[k] * dirf[k + 1]) * den
        up[k] = dn[k] * alb[k] + src[k]
    return up.sum(1), (dn + dirf).sum(1), dirf.sum(1)


def solve_column(T, st, col, clouds=True, aerosols=True):
    """LW + SW two-stream fluxes of one column (state must already be prepared: col_dry filled)."""
    tau, ssa, pf = gas_optics_column(T, "lw", st, col, sw=False)
    g = np.zeros_like(tau)
    add_clouds_aerosols(T, st, col, False, tau, ssa, g, clouds=clouds, aerosols=aerosols)
    lw_up, lw_dn = lw_two_stream(T, st, col, tau, ssa, g, pf)
    tau, ssa, _ = gas_optics_column(T, "sw", st, col, sw=True)
    g = np.zeros_like(tau)
    aod = add_clouds_aerosols(T, st, col, True, tau, ssa, g, clouds=clouds, aerosols=aerosols)
    sw_up, sw_dn, sw_dir = sw_two_stream(T, st, col, tau, ssa, g)
    return dict(lw_up=lw_up, lw_dn=lw_dn, sw_up=sw_up, sw_dn=sw_dn, sw_dir=sw_dir, aod=aod)

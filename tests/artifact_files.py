"""Writes NetCDF files in the RAW variable layout of the rrtmgp-data lookup artifacts (the names, dimension
orders and contributor ordering `ext/lookup_constructors.jl` reads) from post-load arrays, i.e. the inverse
of the reference's constructors, written independently of `rrtmgp.jl_b200/tables.py` so the round trip tests it.

NetCDF classic (scipy) stands in for the artifact's NetCDF-4 container; the variable content is what matters.
"""
from __future__ import annotations

import os

import numpy as np
from scipy.io import netcdf_file

STRLEN = 32


def _chars(names):
    a = np.full((len(names), STRLEN), b" ", dtype="S1")
    for i, n in enumerate(names):
        a[i, :len(n)] = np.frombuffer(n.encode(), dtype="S1")
    return a


class _Writer:
    def __init__(self, path):
        self.nc = netcdf_file(path, "w", version=2)
        self.dims = {}

    def dim(self, name, n):
        if name not in self.dims:
            self.nc.createDimension(name, int(n))
            self.dims[name] = int(n)
        assert self.dims[name] == int(n), name

    def put(self, name, julia_array, julia_dims, dtype="d"):
        """`julia_array` has the dimension order NCDatasets shows (first fastest); the file gets it reversed."""
        a = np.asarray(julia_array)
        assert a.shape == tuple(self.dims[d] for d in julia_dims), (name, a.shape, julia_dims)
        v = self.nc.createVariable(name, dtype, tuple(reversed(julia_dims)))
        v[...] = np.ascontiguousarray(a.T)

    def put_strings(self, name, names, dim):
        v = self.nc.createVariable(name, "c", (dim, "string_len"))
        v[...] = _chars(names)

    def scalar(self, name, value, dtype="d"):
        v = self.nc.createVariable(name, dtype, ())
        v[()] = value   # (assignValue indexes a 0-d array with [:] and fails on numpy 2)

    def close(self):
        self.nc.close()


def _raw_minor(w, arrays, pre, tag, gas_names, lims_gpt, rng):
    """post-load CSR (g-point major) -> raw intervals (interval major, one interval per (band, absorber))."""
    bnd_st = arrays[f"{pre}/minor_{tag}/bnd_st"].astype(int)
    gpt_st = arrays[f"{pre}/minor_{tag}/gpt_st"].astype(int)
    gasdata = arrays[f"{pre}/minor_{tag}/gasdata"].astype(int)
    kmin = arrays[f"{pre}/minor_{tag}/kminor"]
    n_int = gasdata.shape[1]
    n_bnd = lims_gpt.shape[1]
    n_eta, n_t = kmin.shape[:2]
    limits = np.zeros((2, n_int), dtype=np.int32)
    raw_slices = []
    for b in range(n_bnd):
        lo, hi = int(lims_gpt[0, b]), int(lims_gpt[1, b])
        for i_in_b, i in enumerate(range(bnd_st[b] - 1, bnd_st[b + 1] - 1)):
            limits[:, i] = (lo, hi)
            for g in range(lo, hi + 1):     # the interval's contributor slices, one per g-point of its band
                raw_slices.append(kmin[:, :, gpt_st[g - 1] - 1 + i_in_b])
    n_contrib = len(raw_slices)
    w.dim(f"minor_absorber_intervals_{tag}", n_int)
    w.dim(f"contributors_{tag}", max(n_contrib, 1) if n_contrib == 0 else n_contrib)
    raw = np.stack(raw_slices, axis=0) if raw_slices else np.zeros((1, n_eta, n_t))   # (contrib, eta, T)
    w.put(f"kminor_{tag}", raw, (f"contributors_{tag}", "mixing_fraction", "temperature"))
    w.put_strings(f"minor_gases_{tag}", [gas_names[g - 1] if g > 0 else "" for g in gasdata[0]],
                  f"minor_absorber_intervals_{tag}")
    w.put_strings(f"scaling_gas_{tag}", [gas_names[g - 1] if g > 0 else "" for g in gasdata[1]],
                  f"minor_absorber_intervals_{tag}")
    w.dim("pair", 2)
    w.put(f"minor_limits_gpt_{tag}", limits, ("pair", f"minor_absorber_intervals_{tag}"), "i")
    w.put(f"minor_scales_with_density_{tag}", gasdata[2], (f"minor_absorber_intervals_{tag}",), "i")
    w.put(f"scale_by_complement_{tag}", gasdata[3], (f"minor_absorber_intervals_{tag}",), "i")
    starts = np.concatenate([[1], 1 + np.cumsum(limits[1] - limits[0] + 1)[:-1]]) if n_int else np.zeros(0)
    w.put(f"kminor_start_{tag}", starts.astype(np.int32), (f"minor_absorber_intervals_{tag}",), "i")


def write_gas_file(path, arrays, pre, gas_names, seed=0, zero_key_species=True):
    """`pre` = "lw" or "sw"; `arrays` = `make_lut_arrays` output (post-load layouts)."""
    rng = np.random.default_rng(seed)
    A = lambda k: arrays[f"{pre}/{k}"]
    kmajor = A("kmajor")                                   # (eta, p+1, T, gpt)
    n_eta, n_pi, n_t, n_gpt = kmajor.shape
    n_bnd = A("bnd_lims_gpt").shape[1]
    w = _Writer(path)
    w.dim("string_len", STRLEN)
    for name, n in (("bnd", n_bnd), ("gpt", n_gpt), ("atmos_layer", 2), ("temperature", n_t),
                    ("pressure", A("p_ref").size), ("pressure_interp", n_pi), ("mixing_fraction", n_eta),
                    ("absorber", len(gas_names)), ("absorber_ext", len(gas_names) + 1), ("pair", 2),
                    ("minor_absorber", 4)):
        w.dim(name, n)
    w.put_strings("gas_names", gas_names, "absorber")
    w.put_strings("gas_minor", ["co2", "o3", "n2o", "ch4"], "minor_absorber")
    w.put_strings("identifier_minor", ["co2", "o3", "n2o", "ch4"], "minor_absorber")
    ks = A("key_species").copy()
    if zero_key_species:                                   # the artifact marks "no key species" as (0, 0)
        both2 = (ks[0] == 2) & (ks[1] == 2)
        ks[:, both2] = 0
    w.put("key_species", ks, ("pair", "atmos_layer", "bnd"), "i")
    w.put("kmajor", kmajor.transpose(3, 0, 1, 2), ("gpt", "mixing_fraction", "pressure_interp", "temperature"))
    w.put("bnd_limits_gpt", A("bnd_lims_gpt"), ("pair", "bnd"), "i")
    w.put("bnd_limits_wavenumber", A("bnd_lims_wn"), ("pair", "bnd"))
    for tag in ("lower", "upper"):
        _raw_minor(w, arrays, pre, tag, gas_names, A("bnd_lims_gpt").astype(int), rng)
    w.put("press_ref", A("p_ref"), ("pressure",))
    w.put("temp_ref", A("t_ref"), ("temperature",))
    w.put("vmr_ref", A("vmr_ref"), ("atmos_layer", "absorber_ext", "temperature"))
    w.scalar("press_ref_trop", A("params")[0])
    w.scalar("absorption_coefficient_ref_T", 296.0)
    w.scalar("absorption_coefficient_ref_P", 101325.0)
    if pre == "lw":
        w.dim("temperature_Planck", A("t_planck").size)
        w.put("plank_fraction", A("planck_fraction").transpose(3, 0, 1, 2),
              ("gpt", "mixing_fraction", "pressure_interp", "temperature"))
        w.put("temperature_Planck", A("t_planck"), ("temperature_Planck",))
        w.put("totplnk", A("tot_planck"), ("temperature_Planck", "bnd"))
    else:
        w.put("rayl_lower", A("rayl_lower").transpose(2, 0, 1), ("gpt", "mixing_fraction", "temperature"))
        w.put("rayl_upper", A("rayl_upper").transpose(2, 0, 1), ("gpt", "mixing_fraction", "temperature"))
        # solar_src = quiet + (mg - a) facular + (sb - b) sunspot, normalised at load (lookup_constructors.jl:656-665)
        mg, sb = 0.1567652, 902.71260
        total = A("params")[4]
        fac = rng.uniform(0.0, 0.2, n_gpt) * total / n_gpt
        spot = -rng.uniform(0.0, 1e-4, n_gpt) * total / n_gpt
        quiet = A("solar_src_scaled") * total - (mg - 0.1495954) * fac - (sb - 0.00066696) * spot
        w.put("solar_source_quiet", quiet, ("gpt",))
        w.put("solar_source_facular", fac, ("gpt",))
        w.put("solar_source_sunspot", spot, ("gpt",))
        w.scalar("mg_default", mg)
        w.scalar("sb_default", sb)
    w.close()


def write_cloud_file(path, arrays, tag):
    A = lambda k: arrays[f"{tag}/{k}"]
    nband, nrgh, nliq, nice, pair = (int(x) for x in A("dims"))
    w = _Writer(path)
    for name, n in (("nband", nband), ("nrghice", nrgh), ("nsize_liq", nliq), ("nsize_ice", nice), ("pair", pair)):
        w.dim(name, n)
    b = A("bounds")
    w.scalar("radliq_lwr", b[0]); w.scalar("radliq_upr", b[1])
    w.scalar("diamice_lwr", 2.0 * b[2]); w.scalar("diamice_upr", 2.0 * b[3])   # the file holds DIAMETERS
    liq, ice = A("liqdata"), A("icedata")
    for i, q in enumerate(("ext", "ssa", "asy")):
        w.put(f"{q}liq", liq[i * nliq:(i + 1) * nliq], ("nsize_liq", "nband"))
        w.put(f"{q}ice", ice[i * nice:(i + 1) * nice], ("nsize_ice", "nband", "nrghice"))
    w.put("bnd_limits_wavenumber", A("bnd_lims_wn"), ("pair", "nband"))
    w.close()


def write_aerosol_file(path, arrays, tag):
    A = lambda k: arrays[f"{tag}/{k}"]
    nband, nval, nbin, nrh, pair = (int(x) for x in A("dims"))
    w = _Writer(path)
    for name, n in (("nband", nband), ("nval", nval), ("nbin", nbin), ("nrh", nrh), ("pair", pair)):
        w.dim(name, n)
    w.put("merra_aero_bin_lims", A("size_bin_limits"), ("pair", "nbin"))
    w.put("aero_rh", A("rh_levels"), ("nrh",))
    w.put("aero_dust_tbl", A("dust"), ("nval", "nbin", "nband"))
    w.put("aero_salt_tbl", A("sea_salt"), ("nval", "nrh", "nbin", "nband"))
    w.put("aero_sulf_tbl", A("sulfate"), ("nval", "nrh", "nband"))
    w.put("aero_bcar_rh_tbl", A("black_carbon_rh"), ("nval", "nrh", "nband"))
    w.put("aero_bcar_tbl", A("black_carbon"), ("nval", "nband"))
    w.put("aero_ocar_rh_tbl", A("organic_carbon_rh"), ("nval", "nrh", "nband"))
    w.put("aero_ocar_tbl", A("organic_carbon"), ("nval", "nband"))
    w.put("bnd_limits_wavenumber", A("bnd_lims_wn"), ("pair", "nband"))
    w.close()


def write_artifact(directory, arrays, gas_names, seed=0):
    """All six lookup files under the artifact's names (src/ArtifactPaths.jl:34-45)."""
    names = {("gas", "lw"): "rrtmgp-gas-lw-g256.nc", ("gas", "sw"): "rrtmgp-gas-sw-g224.nc",
             ("cloud", "lw"): "rrtmgp-clouds-lw-bnd.nc", ("cloud", "sw"): "rrtmgp-clouds-sw-bnd.nc",
             ("aerosol", "lw"): "rrtmgp-aerosols-merra-lw.nc", ("aerosol", "sw"): "rrtmgp-aerosols-merra-sw.nc"}
    os.makedirs(directory, exist_ok=True)
    write_gas_file(os.path.join(directory, names[("gas", "lw")]), arrays, "lw", gas_names, seed)
    write_gas_file(os.path.join(directory, names[("gas", "sw")]), arrays, "sw", gas_names, seed + 1)
    write_cloud_file(os.path.join(directory, names[("cloud", "lw")]), arrays, "cld_lw")
    write_cloud_file(os.path.join(directory, names[("cloud", "sw")]), arrays, "cld_sw")
    write_aerosol_file(os.path.join(directory, names[("aerosol", "lw")]), arrays, "aero_lw")
    write_aerosol_file(os.path.join(directory, names[("aerosol", "sw")]), arrays, "aero_sw")
    return names

"""Real-table ingestion: rrtmgp-data NetCDF files -> the engine's LUT pack.

The reference builds its lookup structs from six NetCDF files of the rrtmgp-data v1.9 artifact
(`src/ArtifactPaths.jl:28-46`) through `ext/RRTMGPNCDatasetsExt.jl:26-133` and the constructors of
`ext/lookup_constructors.jl`.  This module restates those constructors on numpy arrays and emits the
post-load arrays under the names `rrtmgp_b200_load_luts` reads (`csrc/lut.cu`), so a host with the artifact
can run the engine -- and re-pin parity -- on the real k-distributions without Julia.

File access is kept apart from the table logic:

* `Dataset` is the minimal view the constructors need: `dims[name] -> int` and `var(name) -> ndarray` in
  the FILE's (C, slowest-first) dimension order.  NCDatasets presents the same variable with the dimension
  order reversed (Julia is column-major), which `_jl` reproduces so every index expression below can be
  checked against the Julia source it cites.
* `open_dataset(path)` reads NetCDF classic / 64-bit-offset files with `scipy.io.netcdf_file`.  The
  artifact ships NetCDF-4 (HDF5) files; those are read through `netCDF4` or `h5py` when one of them is
  importable, otherwise through the built-in minimal HDF5 reader `hdf5min.py` (pure Python; its docstring says
  what it covers and that it was validated only against this repo's own test writer); if that refuses the file
  the error says how to convert (`nccopy -k nc6 in.nc out.nc`).  Nothing here falls back to synthetic data.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Tuple

import numpy as np

from .lutpack import pack_luts

ARTIFACT_FILES = {  # src/ArtifactPaths.jl:34-45
    ("gas", "lw"): "rrtmgp-gas-lw-g256.nc",
    ("gas", "sw"): "rrtmgp-gas-sw-g224.nc",
    ("cloud", "lw"): "rrtmgp-clouds-lw-bnd.nc",
    ("cloud", "sw"): "rrtmgp-clouds-sw-bnd.nc",
    ("aerosol", "lw"): "rrtmgp-aerosols-merra-lw.nc",
    ("aerosol", "sw"): "rrtmgp-aerosols-merra-sw.nc",
}

# canonical aerosol slots (lookup_constructors.jl:47-58; src/api/aerosols.jl)
AEROSOL_INDEX = {"dust1": 1, "sea_salt1": 2, "sulfate": 3, "black_carbon_rh": 4, "black_carbon": 5,
                 "organic_carbon_rh": 6, "organic_carbon": 7,
                 **{f"dust{i}": i + 6 for i in range(2, 6)}, **{f"sea_salt{i}": i + 10 for i in range(2, 6)}}


class TableError(ValueError):
    """A lookup file violates an assumption the kernels rely on (the reference `error`s / `@assert`s)."""


class Dataset:
    """`dims` and variables of one file; variables in file (C) dimension order."""

    def __init__(self, dims: Mapping[str, int], variables: Mapping[str, np.ndarray]):
        self.dims = dict(dims)
        self._vars = variables

    def var(self, name: str) -> np.ndarray:
        if name not in self._vars:
            raise TableError(f"lookup file has no variable '{name}'")
        return np.asarray(self._vars[name])

    def dim(self, name: str) -> int:
        if name not in self.dims:
            raise TableError(f"lookup file has no dimension '{name}'")
        return int(self.dims[name])


def open_dataset(path: str) -> Dataset:
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic[:3] == b"CDF" and magic[3] in (1, 2):
        from scipy.io import netcdf_file
        with netcdf_file(path, "r", mmap=False, maskandscale=False) as nc:
            dims = {k: int(v) if v is not None else 0 for k, v in nc.dimensions.items()}
            variables = {k: np.array(v.data) for k, v in nc.variables.items()}
        return Dataset(dims, variables)
    if magic == b"\x89HDF\r\n\x1a\n":
        try:
            import netCDF4  # type: ignore
        except ImportError:
            netCDF4 = None
        if netCDF4 is not None:
            with netCDF4.Dataset(path) as nc:
                nc.set_auto_maskandscale(False)
                dims = {k: len(v) for k, v in nc.dimensions.items()}
                variables = {k: np.array(v[...]) for k, v in nc.variables.items()}
            return Dataset(dims, variables)
        try:
            import h5py  # type: ignore
        except ImportError:
            h5py = None
        if h5py is not None:
            with h5py.File(path, "r") as h5:
                variables = {k: np.array(v[...]) for k, v in h5.items() if isinstance(v, h5py.Dataset)}
                dims = {k: int(v.shape[0]) for k, v in h5.items()
                        if isinstance(v, h5py.Dataset) and v.attrs.get("CLASS", b"") == b"DIMENSION_SCALE"}
            return Dataset(dims, variables)
        # neither library: the built-in minimal reader (hdf5min.py states what it covers and how it was validated)
        from . import hdf5min
        try:
            dims, variables = hdf5min.read_netcdf4(path)
        except hdf5min.HDF5Error as e:
            raise TableError(
                f"{path} is a NetCDF-4/HDF5 file that the built-in reader cannot read ({e}); install netCDF4 or "
                f"h5py, or convert it once with `nccopy -k nc6 {path} out.nc` (64-bit-offset classic format) and "
                f"pass the copy") from None
        return Dataset(dims, variables)
    raise TableError(f"{path}: not a NetCDF file (magic {magic!r})")


# ------------------------------------------------------------------------------------------------------
# helpers mirroring what NCDatasets hands the Julia constructors
# ------------------------------------------------------------------------------------------------------
def _jl(ds: Dataset, name: str) -> np.ndarray:
    """The variable with Julia's dimension order (first index fastest) = file order reversed."""
    return ds.var(name).T


def _scalar(ds: Dataset, name: str) -> float:
    return float(np.asarray(ds.var(name)).reshape(-1)[0])


def _strings(ds: Dataset, name: str) -> list:
    """`strip(String(ds[name][:, i]))` for every i (lookup_constructors.jl:134-141)."""
    a = ds.var(name)  # file order: (n, string_len)
    if a.dtype.kind == "U":
        return [str(s).strip() for s in a.reshape(-1)]
    if a.dtype.kind == "S" and a.ndim == 1:
        return [s.decode("ascii", "ignore").strip("\0 ").strip() for s in a]
    out = []
    for row in a.reshape(a.shape[0], -1):
        raw = b"".join(bytes(c) if isinstance(c, (bytes, np.bytes_)) else bytes([int(c)]) for c in row)
        out.append(raw.decode("ascii", "ignore").replace("\0", " ").strip())
    return out


def _gas_index(ds: Dataset) -> Dict[str, int]:
    idx = {name: i + 1 for i, name in enumerate(_strings(ds, "gas_names"))}   # :134-137
    if "h2o" not in idx or "o3" not in idx or idx["h2o"] != 1 or idx["o3"] != 3:   # :9-16
        raise TableError("unexpected gas ordering in lookup table: the solver kernels require h2o -> 1 and "
                         f"o3 -> 3, got h2o -> {idx.get('h2o')}, o3 -> {idx.get('o3')}.")
    idx["h2o_frgn"] = idx["h2o"]   # :139-141
    idx["h2o_self"] = idx["h2o"]
    return idx


def _minor(ds: Dataset, tag: str, idx_gases: Dict[str, int], gpt2bnd: np.ndarray, bnd_lims_gpt: np.ndarray):
    """`LookUpMinor` of one atmosphere (lookup_constructors.jl:145-171, 218-311, 313-371).

    The file stores, for each minor-absorber interval i (a gas active over the g-points of one band), a
    contiguous run of contributor slices starting at `gpt_sh[i]`; the kernels want them g-point-major
    (for a g-point: the slices of all intervals of its band), addressed by two CSR arrays.
    """
    n_bnd = bnd_lims_gpt.shape[1]
    n_gpt = gpt2bnd.size
    gases = _strings(ds, f"minor_gases_{tag}")
    scaling = _strings(ds, f"scaling_gas_{tag}")
    n_int = len(gases)
    try:
        idx_minor = np.array([idx_gases[g] if g else 0 for g in gases], dtype=np.int64)
        idx_scal = np.array([idx_gases[g] if g else 0 for g in scaling], dtype=np.int64)
    except KeyError as e:
        raise TableError(f"minor absorber / scaling gas {e} is not in gas_names") from None
    lims = _jl(ds, f"minor_limits_gpt_{tag}").astype(np.int64)            # (2, n_int), 1-based, inclusive
    width = lims[1] - lims[0] + 1
    gpt_sh = np.concatenate([[0], np.cumsum(width)[:-1]]) if n_int else np.zeros(0, np.int64)   # :233-252
    bnd_of = gpt2bnd[lims[0] - 1] if n_int else np.zeros(0, np.int64)     # band of each interval, 1-based

    # bnd_st[b] .. bnd_st[b+1]-1 = intervals of band b; `findlast` semantics of :256-271 (bands without an
    # interval inherit the previous start)
    bnd_st = np.ones(n_bnd + 1, dtype=np.int64)
    for b in range(1, n_bnd + 1):
        hits = np.nonzero(bnd_of == b)[0]
        bnd_st[b] = hits[-1] + 2 if hits.size else bnd_st[b - 1]

    gpt_st = np.ones(n_gpt + 1, dtype=np.int64)
    order = []
    for b in range(n_bnd):
        lo, hi = int(bnd_lims_gpt[0, b]), int(bnd_lims_gpt[1, b])
        members = np.arange(bnd_st[b], bnd_st[b + 1]) - 1                  # 0-based interval ids of band b
        for loc_in_bnd, igpt in enumerate(range(lo, hi + 1), start=1):    # :277-297
            gpt_st[igpt] = gpt_st[igpt - 1] + members.size
            order.extend((gpt_sh[members] + loc_in_bnd).tolist())
    n_contrib = ds.dim(f"contributors_{tag}")
    if len(order) != n_contrib:
        raise TableError(f"minor_{tag}: {len(order)} (interval, g-point) pairs but {n_contrib} contributors in the file")
    kminor = _jl(ds, f"kminor_{tag}").transpose(1, 2, 0)                  # (eta, T, contributor)  :299-311
    kminor = kminor[:, :, np.asarray(order, dtype=np.int64) - 1]
    gasdata = np.stack([idx_minor, idx_scal,                             # vcat of four 1 x n rows, :344-353
                        _jl(ds, f"minor_scales_with_density_{tag}").reshape(-1).astype(np.int64),
                        _jl(ds, f"scale_by_complement_{tag}").reshape(-1).astype(np.int64)])
    return {"bnd_st": bnd_st.astype(np.int32), "gpt_st": gpt_st.astype(np.int32),
            "gasdata": gasdata.astype(np.int32).reshape(4, n_int), "kminor": np.asarray(kminor, dtype=np.float64)}


def _gas_common(ds: Dataset) -> Tuple[Dict[str, np.ndarray], Dict[str, int]]:
    """The part `LookUpLW` and `LookUpSW` share (lookup_constructors.jl:95-342 / :408-640)."""
    n_bnd, n_gpt = ds.dim("bnd"), ds.dim("gpt")
    idx_gases = _gas_index(ds)
    out: Dict[str, np.ndarray] = {}

    key_species = _jl(ds, "key_species").astype(np.int32).copy()           # (2, atmos_layer, bnd)
    both_zero = (key_species[0] == 0) & (key_species[1] == 0)             # :175-182
    key_species[:, both_zero] = 2
    out["key_species"] = key_species

    out["kmajor"] = _jl(ds, "kmajor").transpose(1, 2, 3, 0).astype(np.float64)   # (eta, p, T, gpt)  :186
    lims = _jl(ds, "bnd_limits_gpt").astype(np.int32)                     # (2, n_bnd)
    gpt2bnd = np.zeros(n_gpt, dtype=np.int32)
    for b in range(n_bnd):                                                 # :209-212
        gpt2bnd[lims[0, b] - 1:lims[1, b]] = b + 1
    if (gpt2bnd == 0).any():
        raise TableError("bnd_limits_gpt does not cover every g-point")
    out["bnd_lims_gpt"], out["major_gpt2bnd"] = lims, gpt2bnd
    out["bnd_lims_wn"] = _jl(ds, "bnd_limits_wavenumber").astype(np.float64)

    for tag in ("lower", "upper"):
        for k, v in _minor(ds, tag, idx_gases, gpt2bnd.astype(np.int64), lims.astype(np.int64)).items():
            out[f"minor_{tag}/{k}"] = v

    p_ref = _jl(ds, "press_ref").astype(np.float64)
    t_ref = _jl(ds, "temp_ref").astype(np.float64)
    out["p_ref"], out["t_ref"] = p_ref, t_ref
    out["vmr_ref"] = _jl(ds, "vmr_ref").astype(np.float64)                # (atmos_layer, absorber_ext, T)
    out["idx_h2o"] = np.array([idx_gases["h2o"]], dtype=np.int32)
    # p_ref_tropo, p_ref_min, t_ref_min, t_ref_max (:111, :333-339), solar total (SW fills it in)
    out["params"] = np.array([_scalar(ds, "press_ref_trop"), p_ref.min(), t_ref.min(), t_ref.max(), 0.0])
    return out, idx_gases


def lookup_lw(ds: Dataset) -> Tuple[Dict[str, np.ndarray], Dict[str, int]]:
    """`LookUpLW(ds, FT, DA)` (lookup_constructors.jl:83-398) -> (pack arrays without prefix, idx_gases)."""
    out, idx_gases = _gas_common(ds)
    out["planck_fraction"] = _jl(ds, "plank_fraction").transpose(1, 2, 3, 0).astype(np.float64)   # :190 (sic)
    t_planck = _jl(ds, "temperature_Planck").astype(np.float64)
    if not (100.0 <= t_planck[0] and t_planck[-1] <= 500.0):             # :192-201
        raise TableError("`temperature_Planck` in the longwave lookup file does not look like Kelvin "
                         f"(range {t_planck[0]}...{t_planck[-1]}); this file is not usable with the Planck interpolation.")
    out["t_planck"] = t_planck
    out["tot_planck"] = _jl(ds, "totplnk").astype(np.float64)             # (n_t_plnk, n_bnd)
    return out, idx_gases


def lookup_sw(ds: Dataset) -> Tuple[Dict[str, np.ndarray], Dict[str, int]]:
    """`LookUpSW(ds, FT, DA)` (lookup_constructors.jl:400-725)."""
    out, idx_gases = _gas_common(ds)
    out["rayl_lower"] = _jl(ds, "rayl_lower").transpose(1, 2, 0).astype(np.float64)   # (eta, T, gpt)  :653-654
    out["rayl_upper"] = _jl(ds, "rayl_upper").transpose(1, 2, 0).astype(np.float64)
    a_offset, b_offset = 0.1495954, 0.00066696                             # :656-665
    mg = max(_scalar(ds, "mg_default"), 0.0)
    sb = max(_scalar(ds, "sb_default"), 0.0)
    solar = (_jl(ds, "solar_source_quiet").astype(np.float64)
             + (mg - a_offset) * _jl(ds, "solar_source_facular").astype(np.float64)
             + (sb - b_offset) * _jl(ds, "solar_source_sunspot").astype(np.float64))
    total = float(solar.sum())
    out["solar_src_scaled"] = solar / total
    out["params"][4] = total
    return out, idx_gases


def lookup_cld(ds: Dataset) -> Dict[str, np.ndarray]:
    """`LookUpCld(ds, FT, DA)` (lookup_constructors.jl:727-751): ice diameters -> radii, ext/ssa/asy stacked."""
    out = {
        "dims": np.array([ds.dim("nband"), ds.dim("nrghice"), ds.dim("nsize_liq"), ds.dim("nsize_ice"),
                          ds.dim("pair")], dtype=np.int32),
        "bounds": np.array([_scalar(ds, "radliq_lwr"), _scalar(ds, "radliq_upr"),
                            _scalar(ds, "diamice_lwr") / 2.0, _scalar(ds, "diamice_upr") / 2.0]),
        "liqdata": np.concatenate([_jl(ds, n) for n in ("extliq", "ssaliq", "asyliq")], axis=0).astype(np.float64),
        "icedata": np.concatenate([_jl(ds, n) for n in ("extice", "ssaice", "asyice")], axis=0).astype(np.float64),
        "bnd_lims_wn": _jl(ds, "bnd_limits_wavenumber").astype(np.float64),
    }
    return out


def lookup_aerosol(ds: Dataset) -> Dict[str, np.ndarray]:
    """`LookUpAerosolMerra(ds, FT, DA)` (lookup_constructors.jl:18-81)."""
    wn = _jl(ds, "bnd_limits_wavenumber").astype(np.float64)               # (2, nband), cm^-1
    i550 = 0
    for b in range(wn.shape[1]):                                           # findfirst, :41-44
        if 1.0 / (wn[1, b] * 100.0) <= 550e-9 <= 1.0 / (wn[0, b] * 100.0):
            i550 = b + 1
            break
    f = lambda n: _jl(ds, n).astype(np.float64)
    return {
        "dims": np.array([ds.dim("nband"), ds.dim("nval"), ds.dim("nbin"), ds.dim("nrh"), ds.dim("pair")],
                         dtype=np.int32),
        "size_bin_limits": f("merra_aero_bin_lims"), "rh_levels": f("aero_rh"),
        "dust": f("aero_dust_tbl"), "sea_salt": f("aero_salt_tbl"), "sulfate": f("aero_sulf_tbl"),
        "black_carbon_rh": f("aero_bcar_rh_tbl"), "black_carbon": f("aero_bcar_tbl"),
        "organic_carbon_rh": f("aero_ocar_rh_tbl"), "organic_carbon": f("aero_ocar_tbl"),
        "bnd_lims_wn": wn, "iband_550nm": np.array([i550], dtype=np.int32),
    }


def lookup_tables(gas_lw: Dataset, gas_sw: Dataset, cloud_lw: Optional[Dataset] = None,
                  cloud_sw: Optional[Dataset] = None, aerosol_lw: Optional[Dataset] = None,
                  aerosol_sw: Optional[Dataset] = None) -> Tuple[Dict[str, np.ndarray], Dict[str, Dict[str, int]]]:
    """`lookup_tables(method, device, FT)` (ext/RRTMGPNCDatasetsExt.jl:26-133): all arrays of the pack plus
    the name -> slot maps the host needs to fill `vmr` / `aero_mass` (`idx_gases_*`, `idx_aerosol`)."""
    arrays: Dict[str, np.ndarray] = {}
    lw, idx_lw = lookup_lw(gas_lw)
    sw, idx_sw = lookup_sw(gas_sw)
    # the @asserts of RRTMGPNCDatasetsExt.jl:71-75
    if lw["vmr_ref"].shape[1] != sw["vmr_ref"].shape[1]:
        raise TableError("longwave and shortwave lookups disagree on the number of gases")
    for i, what in ((1, "p_ref_min"), (2, "t_ref_min"), (3, "t_ref_max")):
        if lw["params"][i] != sw["params"][i]:
            raise TableError(f"longwave and shortwave lookups disagree on {what}")
    for k, v in lw.items():
        arrays[f"lw/{k}"] = v
    for k, v in sw.items():
        arrays[f"sw/{k}"] = v
    for tag, ds, fn in (("cld_lw", cloud_lw, lookup_cld), ("cld_sw", cloud_sw, lookup_cld),
                        ("aero_lw", aerosol_lw, lookup_aerosol), ("aero_sw", aerosol_sw, lookup_aerosol)):
        if ds is not None:
            for k, v in fn(ds).items():
                arrays[f"{tag}/{k}"] = v
    return arrays, {"idx_gases_lw": idx_lw, "idx_gases_sw": idx_sw, "idx_aerosol": dict(AEROSOL_INDEX)}


def lut_pack_from_artifact(directory: str, *, clouds: bool = True, aerosols: bool = True) -> Tuple[bytes, dict]:
    """LUT pack from a directory holding the rrtmgp-data lookup files under their artifact names."""
    import os
    op = lambda kind, band: open_dataset(os.path.join(directory, ARTIFACT_FILES[(kind, band)]))
    arrays, maps = lookup_tables(
        op("gas", "lw"), op("gas", "sw"),
        op("cloud", "lw") if clouds else None, op("cloud", "sw") if clouds else None,
        op("aerosol", "lw") if aerosols else None, op("aerosol", "sw") if aerosols else None)
    return pack_luts(arrays), maps

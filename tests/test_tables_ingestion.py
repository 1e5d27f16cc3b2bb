"""Real-table ingestion: NetCDF files in the rrtmgp-data artifact's raw layout ->
`rrtmgp.jl_b200/tables.py` (restated `ext/lookup_constructors.jl`) -> the engine's LUT pack.

The raw files are produced by `tests/artifact_files.py`, an independently written INVERSE of the reference's
constructors, so the round trip pins the permutes (`lookup_constructors.jl:186,190,299-311,653-654`), the
minor-contributor reorder (`:218-311`), the key-species rewrite (`:175-182`), diameters -> radii (`:741-743`),
the `vcat(ext, ssa, asy)` stacking (`:745-748`) and the solar-source normalisation (`:656-665`)."""
import os

import numpy as np
import pytest

import rrtmgp_b200 as R
from artifact_files import write_artifact, write_gas_file

T = R.tables
SYN = R.synthetic


@pytest.fixture(scope="module")
def small_arrays():
    return SYN.make_lut_arrays(seed=11, dims=SYN.SMALL_DIMS)


@pytest.fixture(scope="module")
def artifact_dir(tmp_path_factory, small_arrays):
    d = tmp_path_factory.mktemp("rrtmgp_data")
    write_artifact(str(d), small_arrays, SYN.GAS_NAMES)
    return str(d)


def test_round_trip_reproduces_every_post_load_array(artifact_dir, small_arrays):
    pack, maps = T.lut_pack_from_artifact(artifact_dir)
    got = R.lutpack.unpack_luts(pack)
    assert set(got) == set(small_arrays)
    for k, want in small_arrays.items():
        assert got[k].shape == np.asarray(want).shape, k
        if k == "sw/solar_src_scaled" or k == "sw/params":
            np.testing.assert_allclose(got[k], want, rtol=1e-12, err_msg=k)   # quiet + facular + sunspot, renormalised
        else:
            np.testing.assert_array_equal(got[k], np.asarray(want, dtype=got[k].dtype), err_msg=k)
    # name -> slot maps the host fills `vmr` / `aero_mass` with (lookup_constructors.jl:134-141, 47-58)
    assert maps["idx_gases_lw"]["h2o"] == 1 and maps["idx_gases_lw"]["o3"] == 3
    assert maps["idx_gases_lw"]["h2o_self"] == maps["idx_gases_lw"]["h2o_frgn"] == 1
    assert maps["idx_aerosol"]["dust5"] == 11 and maps["idx_aerosol"]["sea_salt2"] == 12


def test_key_species_zero_pairs_are_rewritten(artifact_dir, small_arrays):
    raw = T.open_dataset(os.path.join(artifact_dir, T.ARTIFACT_FILES[("gas", "lw")]))
    ks_raw = raw.var("key_species").T                       # (pair, atmos_layer, bnd)
    assert ((ks_raw[0] == 0) & (ks_raw[1] == 0)).any(), "fixture must contain a (0, 0) pair"
    out, _ = T.lookup_lw(raw)
    assert not ((out["key_species"][0] == 0) & (out["key_species"][1] == 0)).any()
    # a half-zero pair (dry-air member) is NOT rewritten
    assert ((out["key_species"] == 0).sum(axis=0) == 1).any()


def test_minor_reorder_hand_worked_example():
    """2 bands (g-points 1-2 and 3-5); lower intervals i1 (band 1), i2 and i3 (band 2).  The file stores the
    contributor slices interval by interval: i1 -> 1,2; i2 -> 3,4,5; i3 -> 6,7,8.  The reference's loop
    (`lookup_constructors.jl:273-297`) orders them g-point-major: [1,2 | 3,6 | 4,7 | 5,8]."""
    n_eta, n_t = 2, 3
    k = np.arange(1, 9, dtype=np.float64)[None, None, :] * np.ones((n_t, n_eta, 1))   # file order (T, eta, contrib)
    chars = lambda names: np.array([list(n.ljust(8)) for n in names], dtype="S1")
    ds = T.Dataset(
        {"contributors_lower": 8},
        {"minor_gases_lower": chars(["co2", "o3", "n2o"]), "scaling_gas_lower": chars(["", "h2o", ""]),
         "minor_limits_gpt_lower": np.array([[1, 2], [3, 5], [3, 5]], dtype=np.int32),   # file order (interval, pair)
         "kminor_lower": k,
         "minor_scales_with_density_lower": np.array([0, 1, 0], dtype=np.int32),
         "scale_by_complement_lower": np.array([0, 1, 0], dtype=np.int32)})
    idx = {"h2o": 1, "co2": 2, "o3": 3, "n2o": 4}
    gpt2bnd = np.array([1, 1, 2, 2, 2])
    lims = np.array([[1, 3], [2, 5]])
    m = T._minor(ds, "lower", idx, gpt2bnd, lims)
    np.testing.assert_array_equal(m["kminor"][0, 0], [1, 2, 3, 6, 4, 7, 5, 8])
    np.testing.assert_array_equal(m["bnd_st"], [1, 2, 4])
    np.testing.assert_array_equal(m["gpt_st"], [1, 2, 3, 5, 7, 9])
    np.testing.assert_array_equal(m["gasdata"], [[2, 3, 4], [0, 1, 0], [0, 1, 0], [0, 1, 0]])


def test_band_without_minor_absorbers_inherits_the_start():
    """`findlast(...) === nothing` branch (`lookup_constructors.jl:258-264`): band 2 of 3 has no interval."""
    chars = lambda names: np.array([list(n.ljust(8)) for n in names], dtype="S1")
    ds = T.Dataset(
        {"contributors_upper": 3},
        {"minor_gases_upper": chars(["co2", "o3"]), "scaling_gas_upper": chars(["", ""]),
         "minor_limits_gpt_upper": np.array([[1, 1], [3, 4]], dtype=np.int32),
         "kminor_upper": np.arange(1, 4, dtype=np.float64)[None, None, :] * np.ones((2, 2, 1)),
         "minor_scales_with_density_upper": np.zeros(2, np.int32), "scale_by_complement_upper": np.zeros(2, np.int32)})
    m = T._minor(ds, "upper", {"co2": 2, "o3": 3}, np.array([1, 2, 3, 3]), np.array([[1, 2, 3], [1, 2, 4]]))
    np.testing.assert_array_equal(m["bnd_st"], [1, 2, 2, 3])
    np.testing.assert_array_equal(m["gpt_st"], [1, 2, 2, 3, 4])
    np.testing.assert_array_equal(m["kminor"][0, 0], [1, 2, 3])


def test_rejects_noncanonical_gas_order(tmp_path, small_arrays):
    """`_assert_canonical_gas_slots` (`lookup_constructors.jl:9-16`): h2o -> 1 and o3 -> 3 are hard-coded in the kernels."""
    names = list(SYN.GAS_NAMES)
    names[1], names[2] = names[2], names[1]   # o3 in slot 2
    p = str(tmp_path / "bad.nc")
    write_gas_file(p, small_arrays, "lw", names)
    with pytest.raises(T.TableError, match="h2o -> 1 and o3 -> 3"):
        T.lookup_lw(T.open_dataset(p))


def test_rejects_index_valued_planck_temperatures(artifact_dir):
    """The g128 files store `temperature_Planck` as an index, not Kelvin (`lookup_constructors.jl:192-201`)."""
    ds = T.open_dataset(os.path.join(artifact_dir, T.ARTIFACT_FILES[("gas", "lw")]))
    n = ds.var("temperature_Planck").size
    variables = dict(ds._vars)
    variables["temperature_Planck"] = np.arange(n, dtype=np.float64)
    with pytest.raises(T.TableError, match="does not look like Kelvin"):
        T.lookup_lw(T.Dataset(ds.dims, variables))


def test_lw_sw_consistency_asserts(artifact_dir):
    """The `@assert`s of `ext/RRTMGPNCDatasetsExt.jl:71-75`."""
    op = lambda kind, band: T.open_dataset(os.path.join(artifact_dir, T.ARTIFACT_FILES[(kind, band)]))
    sw = op("gas", "sw")
    variables = dict(sw._vars)
    variables["temp_ref"] = np.asarray(variables["temp_ref"]) + 1.0
    with pytest.raises(T.TableError, match="t_ref_min"):
        T.lookup_tables(op("gas", "lw"), T.Dataset(sw.dims, variables))


def test_unreadable_hdf5_container_says_how_to_convert(tmp_path):
    # (readable HDF5 containers: tests/test_hdf5min.py)
    p = tmp_path / "x.nc"
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\x07" + b"\0" * 64)   # superblock version 7 does not exist
    try:
        import netCDF4  # noqa: F401
        pytest.skip("netCDF4 is installed")
    except ImportError:
        pass
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py is installed")
    except ImportError:
        pass
    with pytest.raises(T.TableError, match="nccopy"):
        T.open_dataset(str(p))
    (tmp_path / "y.nc").write_bytes(b"not a netcdf file")
    with pytest.raises(T.TableError, match="not a NetCDF file"):
        T.open_dataset(str(tmp_path / "y.nc"))


def test_clear_sky_pack_has_no_cloud_or_aerosol_sections(artifact_dir):
    pack, _ = T.lut_pack_from_artifact(artifact_dir, clouds=False, aerosols=False)
    names = set(R.lutpack.unpack_luts(pack))
    assert not any(n.startswith(("cld_", "aero_")) for n in names)
    assert {"lw/kmajor", "sw/kmajor", "lw/planck_fraction", "sw/rayl_lower"} <= names


def test_oracle_on_ingested_tables_matches_direct_tables(artifact_dir, small_arrays):
    """End to end on the CPU: fluxes from the ingested pack == fluxes from the directly built pack."""
    from oracle import Oracle
    pack_i, _ = T.lut_pack_from_artifact(artifact_dir)
    pack_d = R.lutpack.pack_luts(small_arrays)
    st = SYN.make_atmosphere(4, 12, dtype=np.float64, seed=3)
    a = Oracle(pack_i, np.float64).update_fluxes(st, seed=5, method="all_sky", aerosols=True)
    b = Oracle(pack_d, np.float64).update_fluxes(st, seed=5, method="all_sky", aerosols=True)
    for k in ("lw_up", "lw_dn", "sw_up", "sw_dn", "sw_dir", "net"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-11, atol=1e-11, err_msg=k)


def test_clear_sky_pack_and_ln_p_ref_run_through_the_oracle(artifact_dir, small_arrays):
    """A gas-only pack serves the clear-sky method; `ln_p_ref` (what a dump of the loaded Julia struct holds,
    LookUpTables.jl:70-74) is accepted in place of `p_ref`."""
    from oracle import Oracle
    pack_c, _ = T.lut_pack_from_artifact(artifact_dir, clouds=False, aerosols=False)
    st = SYN.make_atmosphere(3, 10, dtype=np.float64, seed=4, clouds=False, aerosols=False)
    a = Oracle(pack_c, np.float64).update_fluxes(st, method="clear_sky", aerosols=False)
    arrays = dict(small_arrays)
    for pre in ("lw", "sw"):
        arrays[f"{pre}/ln_p_ref"] = np.log(arrays.pop(f"{pre}/p_ref"))
    b = Oracle(R.lutpack.pack_luts(arrays), np.float64).update_fluxes(st, method="clear_sky", aerosols=False)
    for k in ("lw_up", "lw_dn", "sw_up", "sw_dn", "net"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-11, atol=1e-11, err_msg=k)


# ---- the contributor reorder against a loop-for-loop restatement on random band layouts -------------------------
def _reorder_by_the_book(n_bnd, bnd_lims_gpt, lims):
    """`lookup_constructors.jl:218-297` restated index for index with 1-based arrays (slot 0 unused), kept apart
    from the vectorised `tables._minor` so the two can disagree."""
    n_gpt = int(bnd_lims_gpt[1][-1])
    n_int = len(lims[0])
    gpt2bnd = [0] * (n_gpt + 1)
    for i in range(1, n_bnd + 1):
        for g in range(bnd_lims_gpt[0][i - 1], bnd_lims_gpt[1][i - 1] + 1):
            gpt2bnd[g] = i
    bnd = [0] * (n_int + 1)
    sh = [0] * (n_int + 1)
    for i in range(1, n_int + 1):
        bnd[i] = gpt2bnd[lims[0][i - 1]]
        if i > 1:
            sh[i] = sh[i - 1] + lims[1][i - 2] - lims[0][i - 2] + 1
    bnd_st = [0] * (n_bnd + 2)
    bnd_st[1] = 1
    for ib in range(2, n_bnd + 2):
        loc = None
        for i in range(1, n_int + 1):
            if bnd[i] == ib - 1:
                loc = i                      # findlast
        bnd_st[ib] = bnd_st[ib - 1] if loc is None else loc + 1
    gpt_st = [1] * (n_gpt + 2)
    reorder = []
    for ib in range(1, n_bnd + 1):
        n = bnd_st[ib + 1] - bnd_st[ib]
        for loc_in_bnd, igpt in enumerate(range(bnd_lims_gpt[0][ib - 1], bnd_lims_gpt[1][ib - 1] + 1), start=1):
            gpt_st[igpt + 1] = gpt_st[igpt] + n
            for i in range(bnd_st[ib], bnd_st[ib + 1]):
                reorder.append(sh[i] + loc_in_bnd)
    return bnd_st[1:], gpt_st[1:], reorder


def test_minor_reorder_matches_the_loops_on_random_layouts():
    from hypothesis import given, settings, strategies as hs

    @settings(max_examples=60, deadline=None)
    @given(hs.lists(hs.integers(1, 6), min_size=1, max_size=7), hs.data())
    def run(gpts, data):
        n_bnd = len(gpts)
        hi = np.cumsum(gpts)
        lo = hi - np.array(gpts) + 1
        counts = [data.draw(hs.integers(0, 3)) for _ in range(n_bnd)]      # intervals per band, bands may have none
        lims = [[], []]
        for b in range(n_bnd):
            for _ in range(counts[b]):
                lims[0].append(int(lo[b])); lims[1].append(int(hi[b]))
        n_int = len(lims[0])
        n_contrib = int(sum(c * g for c, g in zip(counts, gpts)))
        if n_int == 0:
            return
        chars = lambda names: np.array([list(n.ljust(8)) for n in names], dtype="S1")
        k = np.arange(1, n_contrib + 1, dtype=np.float64)[None, None, :] * np.ones((2, 3, 1))   # file order (T, eta, contrib)
        ds = T.Dataset({"contributors_lower": n_contrib},
                       {"minor_gases_lower": chars(["co2"] * n_int), "scaling_gas_lower": chars([""] * n_int),
                        "minor_limits_gpt_lower": np.array(lims, dtype=np.int32).T.copy(),
                        "kminor_lower": k, "minor_scales_with_density_lower": np.zeros(n_int, np.int32),
                        "scale_by_complement_lower": np.zeros(n_int, np.int32)})
        gpt2bnd = np.concatenate([[b + 1] * g for b, g in enumerate(gpts)])
        m = T._minor(ds, "lower", {"co2": 2}, gpt2bnd, np.array([lo, hi]))
        bnd_st, gpt_st, reorder = _reorder_by_the_book(n_bnd, [lo.tolist(), hi.tolist()], lims)
        np.testing.assert_array_equal(m["bnd_st"], bnd_st)
        np.testing.assert_array_equal(m["gpt_st"], gpt_st)
        np.testing.assert_array_equal(m["kminor"][0, 0], reorder)      # slice c carries the value c

    run()

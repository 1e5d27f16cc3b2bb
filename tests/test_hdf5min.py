"""`rrtmgp.jl_b200/hdf5min.py`: the pure-Python reader for the HDF5 container of the rrtmgp-data NetCDF-4 files
(`src/ArtifactPaths.jl:28-46` names the files, `ext/RRTMGPNCDatasetsExt.jl:26-133` reads them).

No HDF5 library or file exists in this image, so the files come from `tests/hdf5_files.py`, a writer of the same
on-disk structures written separately from the reader (both follow the published format specification).  The
round trips cover both container generations and every layout / chunk index / filter the reader claims; the last
tests push the whole rrtmgp-data layout through the HDF5 path into `tables.lookup_tables`."""
import itertools
import os

import numpy as np
import pytest

import rrtmgp_b200 as R
from hdf5_files import write_hdf5

H = R.hdf5min
T = R.tables


def _variables(seed=0, many=0):
    rng = np.random.default_rng(seed)
    v = {
        "kmajor": rng.standard_normal((5, 7, 9, 6)),                                   # float64, ragged edge chunks
        "key_species": rng.integers(0, 7, (4, 2, 2)).astype("<i4"),
        "gas_names": np.frombuffer(b"h2o     co2     o3      ", dtype="S1").reshape(3, 8).copy(),
        "press_ref_trop": np.array(9948.43),                                          # scalar
        "temp_ref": np.linspace(160.0, 355.0, 7),                                     # coordinate variable of a dimension
        "flags": rng.integers(0, 2, (5,)).astype("i1"),
        "single": rng.standard_normal((3, 4)).astype("<f4"),
        "be": rng.standard_normal((2, 3)).astype(">f8"),                             # big-endian on disk
        "pair": rng.integers(0, 100, (6, 2)).astype("<u2"),                           # shares its name with a dimension
    }
    for i in range(many):
        v[f"extra_variable_with_a_long_name_{i:03d}"] = rng.standard_normal((2, 3))
    dims = {"temperature": 7, "temp_ref": 7, "pair": 2, "gpt": 6, "mixing_fraction": 9, "absorber": 3, "string_len": 8}
    return v, dims


def _check(path, v, dims):
    got_dims, got = H.read_netcdf4(path)
    assert got_dims == dims
    assert set(got) == set(v)
    for k, a in v.items():
        assert got[k].shape == a.shape, k
        assert got[k].dtype.kind == a.dtype.kind and got[k].dtype.itemsize == a.dtype.itemsize, k
        np.testing.assert_array_equal(got[k], a, err_msg=k)


DEFLATE, SHUFFLE, FLETCHER = (1, [4]), (2, [8]), (3, [])


@pytest.mark.parametrize("style", ["v0", "v2"])
@pytest.mark.parametrize("layout", ["compact", "contiguous", "chunked"])
@pytest.mark.parametrize("split", [False, True])
def test_round_trip_layouts(tmp_path, style, layout, split):
    v, dims = _variables(1)
    p = str(tmp_path / "a.nc")
    write_hdf5(p, v, dims, style=style, layout=layout, split_headers=split)
    _check(p, v, dims)


@pytest.mark.parametrize("style", ["v0", "v2"])
@pytest.mark.parametrize("pipeline", [[DEFLATE], [(2, []), DEFLATE], [(2, []), DEFLATE, FLETCHER], [FLETCHER]])
def test_round_trip_filters(tmp_path, style, pipeline):
    v, dims = _variables(2)
    # shuffle's element size comes from its client data (libhdf5 always writes it) or, absent, from the datatype
    p = str(tmp_path / "f.nc")
    write_hdf5(p, v, dims, style=style, layout="chunked", filters=pipeline, chunk={"kmajor": (2, 3, 4, 6)})
    _check(p, v, dims)


def test_shuffle_client_data_element_size(tmp_path):
    v = {"x": np.arange(24, dtype="<f8").reshape(4, 6)}
    p = str(tmp_path / "s.nc")
    write_hdf5(p, v, {}, style="v0", layout="chunked", filters=[SHUFFLE, DEFLATE], chunk={"x": (3, 4)})
    _check(p, v, {})


def test_two_level_trees_of_the_classic_format(tmp_path):
    # 75 links -> 10 symbol-table nodes under 3 leaf B-tree nodes under one root; 3 x 4 x 5 x 1 = 60 chunks under 8 leaves
    v, dims = _variables(3, many=60)
    p = str(tmp_path / "t.nc")
    write_hdf5(p, v, dims, style="v0", layout="chunked", group_fanout=4, chunk_fanout=8, chunk={"kmajor": (2, 2, 2, 6)},
               filters=[DEFLATE])
    _check(p, v, dims)


@pytest.mark.parametrize("many,heap", [(0, {}), (60, {}), (200, {"max_direct": 1024}), (60, {"checksum_blocks": False})])
def test_dense_links_fractal_heap_and_btree_v2(tmp_path, many, heap):
    # 16 links fit the root direct block and one leaf; 76 need the root indirect block and a depth-1 B-tree;
    # 216 spill into nested indirect blocks (rows beyond the largest direct block size)
    v, dims = _variables(4, many=many)
    p = str(tmp_path / "d.nc")
    write_hdf5(p, v, dims, style="v2", dense=True, heap_kwargs=heap)
    _check(p, v, dims)


def test_dense_links_without_a_usable_name_index(tmp_path):
    v, dims = _variables(5, many=60)
    p = str(tmp_path / "d.nc")
    write_hdf5(p, v, dims, style="v2", dense=True)
    raw = bytearray(open(p, "rb").read())
    at = raw.index(b"BTHD")
    raw[at:at + 4] = b"XXXX"                       # the reader falls back to scanning the heap's direct blocks
    open(p, "wb").write(bytes(raw))
    _check(p, v, dims)


@pytest.mark.parametrize("filters", [[], [(2, []), DEFLATE]])
@pytest.mark.parametrize("page_bits", [10, 2])
def test_latest_format_chunk_indexes(tmp_path, filters, page_bits):
    v, dims = _variables(6)
    p = str(tmp_path / "c.nc")
    # kmajor: 3 x 4 x 3 x 1 = 36 chunks -> fixed array (paged when page_bits = 2); the others one chunk each
    write_hdf5(p, v, dims, style="v2", layout="chunked", filters=filters, page_bits=page_bits,
               chunk={"kmajor": (2, 2, 4, 6), "key_species": (4, 2, 2), "single": (3, 4), "gas_names": (3, 8),
                      "temp_ref": (7,), "flags": (5,), "be": (2, 3), "pair": (6, 2)},
               hflags=0x34)
    _check(p, v, dims)
    write_hdf5(p, v, dims, style="v2", layout="implicit", chunk={"kmajor": (2, 2, 4, 6)}, hflags=0x00)
    _check(p, v, dims)


def test_unsupported_content_is_skipped_or_refused(tmp_path):
    v, dims = _variables(7)
    p = str(tmp_path / "u.nc")
    write_hdf5(p, v, dims, style="v0", extra_vlen="names_as_nc_string")       # NC_STRING: skipped, the rest is read
    _check(p, v, dims)
    write_hdf5(p, v, dims, style="v0", layout="chunked", filters=[(32015, [3])])   # e.g. zstd
    with pytest.raises(H.HDF5Error, match="filter 32015"):
        H.read_netcdf4(p)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[: len(raw) // 2])
    with pytest.raises(H.HDF5Error):
        H.read_netcdf4(p)
    open(p, "wb").write(b"\x89HDF\r\n\x1a\n" + b"\x05" + b"\0" * 100)
    with pytest.raises(H.HDF5Error, match="superblock version"):
        H.read_netcdf4(p)


def test_structure_sizes_match_a_default_libhdf5_file():
    """An empty file written by libhdf5 with default settings has its root object header at byte 96, the root
    group's B-tree node at 136 and its local heap at 680: 96-byte version-0 superblock, 40-byte object header
    (16 prefix + 8 message header + 16 symbol-table message), 544-byte group B-tree node (K = 16)."""
    import hdf5_files as W
    buf = W._Buf()
    buf.add(b"\0" * 96)
    root = W._ohdr_v1(buf, [(0x11, b"\0" * 16)], False)
    assert (root, buf.tell()) == (96, 136)
    assert 24 + (2 * 16 + 1) * 8 + 2 * 16 * 8 == 680 - 136
    assert 8 + 2 * 4 * 40 == 328                                           # symbol-table node, leaf K = 4


# ---- the rrtmgp-data layout through the HDF5 path ---------------------------------------------------------
@pytest.fixture(scope="module")
def classic_dir(tmp_path_factory):
    from artifact_files import write_artifact
    arrays = R.synthetic.make_lut_arrays(seed=11, dims=R.synthetic.SMALL_DIMS)
    d = tmp_path_factory.mktemp("classic")
    write_artifact(str(d), arrays, R.synthetic.GAS_NAMES)
    return str(d)


@pytest.mark.parametrize("style,kw", [("v0", dict(layout="chunked", filters=[(2, []), DEFLATE])),
                                       ("v2", dict(dense=True, layout="contiguous")),
                                       ("v2", dict(layout="chunked", filters=[DEFLATE]))])
def test_artifact_layout_through_the_hdf5_container(tmp_path, classic_dir, style, kw):
    from scipy.io import netcdf_file
    for fname in T.ARTIFACT_FILES.values():
        with netcdf_file(os.path.join(classic_dir, fname), "r", mmap=False, maskandscale=False) as nc:
            dims = {k: int(n) for k, n in nc.dimensions.items()}
            variables = {k: np.array(x.data) for k, x in nc.variables.items()}
        write_hdf5(str(tmp_path / fname), variables, dims, style=style, **kw)
    try:
        import netCDF4  # noqa: F401
        pytest.skip("netCDF4 is installed: open_dataset would not use the built-in reader")
    except ImportError:
        pass
    want, maps_w = T.lookup_tables(*(T.open_dataset(os.path.join(classic_dir, T.ARTIFACT_FILES[k]))
                                     for k in itertools.product(("gas", "cloud", "aerosol"), ("lw", "sw"))))
    got, maps_g = T.lookup_tables(*(T.open_dataset(str(tmp_path / T.ARTIFACT_FILES[k]))
                                    for k in itertools.product(("gas", "cloud", "aerosol"), ("lw", "sw"))))
    assert maps_w == maps_g and set(want) == set(got)
    for k in want:
        np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(want[k]), err_msg=k)
    pack_w, _ = T.lut_pack_from_artifact(classic_dir)
    pack_g, _ = T.lut_pack_from_artifact(str(tmp_path))
    assert pack_w == pack_g


# ---- property test: random shapes / chunkings / filters / container styles -------------------------------
from hypothesis import HealthCheck, given, settings, strategies as st_


@st_.composite
def _random_file(draw):
    rank = draw(st_.integers(0, 4))
    shape = tuple(draw(st_.integers(1, 7)) for _ in range(rank))
    dtype = draw(st_.sampled_from(["<f8", "<f4", ">f8", "<i4", "<i2", "<u1", ">i4", "S1", "S5"]))
    style = draw(st_.sampled_from(["v0", "v2"]))
    layout = draw(st_.sampled_from(["compact", "contiguous", "chunked", "chunked", "implicit"]))
    if style == "v0" and layout == "implicit":
        layout = "chunked"
    chunk = tuple(draw(st_.integers(1, s)) for s in shape)
    filters = draw(st_.sampled_from([[], [DEFLATE], [(2, []), DEFLATE], [(2, []), DEFLATE, FLETCHER]])) if layout == "chunked" else []
    extra = draw(st_.integers(0, 12))
    seed = draw(st_.integers(0, 2 ** 16))
    return shape, dtype, style, layout, chunk, filters, extra, seed, draw(st_.booleans()), draw(st_.sampled_from([1, 2, 10]))


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(_random_file())
def test_random_files_round_trip(tmp_path, case):
    shape, dtype, style, layout, chunk, filters, extra, seed, dense, page_bits = case
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64))
    if dt.kind == "S":
        a = np.array([bytes(rng.integers(65, 91, dt.itemsize).astype(np.uint8)) for _ in range(n)], dtype=dt).reshape(shape)
    elif dt.kind == "f":
        a = rng.standard_normal(shape).astype(dt)
    else:
        a = rng.integers(0, 100, shape).astype(dt)
    v = {"x": a}
    for i in range(extra):
        v[f"v{i}"] = rng.standard_normal((2,)).astype("<f4")
    p = str(tmp_path / "r.nc")
    write_hdf5(p, v, {"d0": 3}, style=style, layout=layout, chunk={"x": chunk}, filters=filters, dense=dense and style == "v2",
               page_bits=page_bits, split_headers=bool(seed & 1))
    dims, got = H.read_netcdf4(p)
    assert dims == {"d0": 3} and set(got) == set(v)
    for k in v:
        assert got[k].shape == v[k].shape and got[k].dtype.itemsize == v[k].dtype.itemsize
        np.testing.assert_array_equal(got[k], v[k])

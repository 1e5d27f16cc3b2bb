"""Flat little-endian binary "LUT pack": the container the engine loads its
k-distribution / Planck / cloud / aerosol lookup tables from.

The reference builds its tables from NetCDF (`ext/RRTMGPNCDatasetsExt.jl:18-133`,
`ext/lookup_constructors.jl`), which is out of scope here (no NetCDF in this image).
The pack stores every array in exactly the *post-load* layout of the reference
structs (`src/optics/LookUpTables.jl:36-41,70-74,87-91,104-108,130-143,185-201,
239-245,312-325`): Julia column-major, first listed dimension fastest, 1-based
integer tables. A Julia host can therefore dump `Adapt.adapt(Array, bundle)`
array by array and the engine re-lays the arrays out for
its kernels at `rrtmgp_b200_load_luts` time.

File layout (all little endian):
    0   char[16]  magic  "RRTMGPB200LUT\\0\\0\\0"
    16  u32       version (=1)
    20  u32       n_entries
    24  u64       total_bytes
    32  entry[n_entries], 80 bytes each:
            char[32] name, u32 dtype (0=f64, 1=i32), u32 ndim, u32 dims[6],
            u64 offset, u64 nbytes
    ... data blocks, each 64-byte aligned, column-major
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

MAGIC = b"RRTMGPB200LUT\0\0\0"
VERSION = 1
_ENTRY = struct.Struct("<32sII6IQQ")
_HEADER = struct.Struct("<16sIIQ")
DT_F64, DT_I32 = 0, 1


def pack_luts(arrays: Dict[str, np.ndarray]) -> bytes:
    """Serialise `name -> ndarray` (shape in reference/Julia dimension order)."""
    names = list(arrays)
    n = len(names)
    offset = _HEADER.size + n * _ENTRY.size
    offset = (offset + 63) // 64 * 64
    entries, blobs = [], []
    for name in names:
        a = np.asarray(arrays[name])
        if a.dtype.kind == "f":
            a = a.astype("<f8")
            dt = DT_F64
        elif a.dtype.kind in "iub":
            a = a.astype("<i4")
            dt = DT_I32
        else:
            raise TypeError(f"{name}: unsupported dtype {a.dtype}")
        if a.ndim > 6:
            raise ValueError(f"{name}: ndim {a.ndim} > 6")
        raw = np.asfortranarray(a).tobytes(order="F")
        dims = list(a.shape) + [1] * (6 - a.ndim)
        bname = name.encode()
        if len(bname) > 31:
            raise ValueError(f"name too long: {name}")
        entries.append(_ENTRY.pack(bname, dt, a.ndim, *dims, offset, len(raw)))
        pad = (-len(raw)) % 64
        blobs.append(raw + b"\0" * pad)
        offset += len(raw) + pad
    head = _HEADER.pack(MAGIC, VERSION, n, offset)
    body = head + b"".join(entries)
    body += b"\0" * ((-len(body)) % 64)
    return body + b"".join(blobs)


def unpack_luts(buf: bytes) -> Dict[str, np.ndarray]:
    """Inverse of `pack_luts`; arrays come back Fortran-ordered with reference shapes."""
    magic, version, n, total = _HEADER.unpack_from(buf, 0)
    if magic != MAGIC:
        raise ValueError("not an RRTMGP-B200 LUT pack (bad magic)")
    if version != VERSION:
        raise ValueError(f"unsupported LUT pack version {version}")
    if total != len(buf):
        raise ValueError(f"LUT pack truncated: header says {total}, got {len(buf)}")
    out: Dict[str, np.ndarray] = {}
    for i in range(n):
        rec = _ENTRY.unpack_from(buf, _HEADER.size + i * _ENTRY.size)
        name = rec[0].split(b"\0", 1)[0].decode()
        dt, ndim = rec[1], rec[2]
        dims = rec[3:9][:ndim]
        offset, nbytes = rec[9], rec[10]
        dtype = "<f8" if dt == DT_F64 else "<i4"
        a = np.frombuffer(buf, dtype=dtype, count=nbytes // np.dtype(dtype).itemsize, offset=offset)
        out[name] = a.reshape(dims, order="F")
    return out

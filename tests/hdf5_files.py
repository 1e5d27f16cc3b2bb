"""Writes small HDF5 files the way NetCDF-4 lays them out, for the tests of `rrtmgp.jl_b200/hdf5min.py`.

Test infrastructure: an independent WRITER of the structures the reader parses, following the "HDF5 File Format
Specification Version 3.0" (no HDF5 library exists in this image).  Two container generations:

* `style="v0"`: what libhdf5 writes by default (NetCDF-4 files made with netcdf-c < 4.? and most tools): superblock
  version 0, version-1 object headers (optionally with a continuation block), root group as symbol table = version-1
  B-tree (one or two levels) + local heap + symbol-table nodes, version-1 dataspace / attribute / filter messages,
  layout message version 3: compact, contiguous, or chunked with a version-1 chunk B-tree (one or two levels).
* `style="v2"`: the "latest" format: superblock version 2, version-2 object headers (`OHDR` / `OCHK`, optional time
  and creation-order fields), links as compact link messages or dense in a fractal heap (root direct block, or
  root indirect block with direct and nested indirect children) + version-2 B-tree name index (depth 0 or 1),
  version-2 dataspace / version-3 attribute / version-2 filter messages, layout version 4 with the single-chunk,
  implicit and fixed-array (plain and paged) chunk indexes.

Checksums of version-2 structures are written as zero (the reader does not verify them).
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


def _p8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def tell(self) -> int:
        return len(self.b)

    def add(self, data: bytes, align: int = 8) -> int:
        self.b += b"\0" * (-len(self.b) % align)
        pos = len(self.b)
        self.b += data
        return pos

    def patch(self, pos: int, data: bytes) -> None:
        self.b[pos:pos + len(data)] = data


# ---- messages -------------------------------------------------------------------------------------------
def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    be = 1 if dt.byteorder == ">" else 0
    if dt.kind in "iu":
        bits0 = be | (0x08 if dt.kind == "i" else 0)
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f":
        n = dt.itemsize
        exp_loc, exp_sz, man_sz, bias = {4: (23, 8, 23, 127), 8: (52, 11, 52, 1023), 2: (10, 5, 10, 15)}[n]
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20 | be, 8 * n - 1, 0, n, 0, 8 * n, exp_loc, exp_sz, 0, man_sz, bias)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)
    raise TypeError(dt)


def _vlen_string_datatype() -> bytes:   # class 9, type = string: the reader must skip such datasets, not fail
    return struct.pack("<BBBBI", 0x19, 0x01, 0, 0, 16) + struct.pack("<BBBBI", 0x13, 0x00, 0, 0, 1)


def _dataspace(shape: Sequence[int], version: int) -> bytes:
    rank = len(shape)
    if version == 1:
        return struct.pack("<BBBBI", 1, rank, 1 if rank else 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape) * (2 if rank else 1)
    return struct.pack("<BBBB", 2, rank, 0, 1 if rank else 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _attribute(name: str, value: bytes, version: int) -> bytes:
    nm = name.encode() + b"\0"
    dt = _datatype(np.dtype(f"S{len(value)}"))
    ds = _dataspace((), 1 if version == 1 else 2)
    if version == 1:
        return struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _p8(nm) + _p8(dt) + _p8(ds) + value
    return struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(ds), 0) + nm + dt + ds + value


def _filters(pipeline: Sequence[Tuple[int, Sequence[int]]], version: int) -> bytes:
    if version == 1:
        out = struct.pack("<BBHI", 1, len(pipeline), 0, 0)
        for fid, cd in pipeline:
            out += struct.pack("<HHHH", fid, 0, 1, len(cd)) + b"".join(struct.pack("<I", c) for c in cd)
            if len(cd) % 2: out += b"\0" * 4
        return out
    out = struct.pack("<BB", 2, len(pipeline))
    for fid, cd in pipeline:
        out += struct.pack("<HHH", fid, 1, len(cd)) + b"".join(struct.pack("<I", c) for c in cd)
    return out


def _apply_filters(raw: bytes, pipeline: Sequence[Tuple[int, Sequence[int]]], itemsize: int) -> bytes:
    for fid, cd in pipeline:
        if fid == 2:
            n = len(raw) // itemsize
            raw = np.frombuffer(raw, np.uint8, n * itemsize).reshape(n, itemsize).T.tobytes() + raw[n * itemsize:]
        elif fid == 1:
            raw = zlib.compress(raw, cd[0] if cd else 4)
        elif fid == 3:
            raw = raw + b"\xde\xad\xbe\xef"   # (a real file holds the Fletcher-32 sum here; the reader strips it)
        else:
            raw = raw[::-1]                    # an "unknown" filter for the refusal test
    return raw


def _link(name: str, addr: int, order: Optional[int] = None) -> bytes:
    nm = name.encode()
    flags = 0x00 | (0x04 if order is not None else 0) | 0x10       # 1-byte name length, charset present
    out = struct.pack("<BB", 1, flags)
    if order is not None: out += struct.pack("<Q", order)
    return out + struct.pack("<BB", 0, len(nm)) + nm + struct.pack("<Q", addr)


# ---- object headers -------------------------------------------------------------------------------------
def _ohdr_v1(buf: _Buf, msgs: List[Tuple[int, bytes]], split: bool) -> int:
    enc = lambda t, d: struct.pack("<HHBBBB", t, len(_p8(d)), 0, 0, 0, 0) + _p8(d)
    first, second = (msgs, []) if not split or len(msgs) < 3 else (msgs[:2], msgs[2:])
    cont_pos = None
    if second:
        block = b"".join(enc(t, d) for t, d in second)
        cont_pos = buf.add(block)
        first = first + [(0x10, struct.pack("<QQ", cont_pos, len(block)))]
    body = b"".join(enc(t, d) for t, d in first)
    nmsg = len(first) + len(second)
    return buf.add(struct.pack("<BBHII", 1, 0, nmsg, 1, len(body)) + b"\0" * 4 + body)


def _ohdr_v2(buf: _Buf, msgs: List[Tuple[int, bytes]], split: bool, hflags: int) -> int:
    track = bool(hflags & 0x04)
    enc = lambda i, t, d: struct.pack("<BHB", t, len(d), 0) + (struct.pack("<H", i) if track else b"") + d
    first, second = (msgs, []) if not split or len(msgs) < 3 else (msgs[:2], msgs[2:])
    if second:
        block = b"OCHK" + b"".join(enc(i, t, d) for i, (t, d) in enumerate(second)) + b"\0" * 4
        cont_pos = buf.add(block)
        first = first + [(0x10, struct.pack("<QQ", cont_pos, len(block)))]
    body = b"".join(enc(i, t, d) for i, (t, d) in enumerate(first)) + b"\0" * 3   # a gap too small for a message
    head = b"OHDR" + struct.pack("<BB", 2, (hflags & ~3) | 1)
    if hflags & 0x20: head += struct.pack("<IIII", 1, 2, 3, 4)
    if hflags & 0x10: head += struct.pack("<HH", 8, 6)
    head += struct.pack("<H", len(body))
    return buf.add(head + body + b"\0" * 4)


# ---- chunk storage ------------------------------------------------------------------------------------------
def _chunk_list(a: np.ndarray, cdims: Tuple[int, ...]):
    grid = tuple((s + c - 1) // c for s, c in zip(a.shape, cdims))
    for idx in np.ndindex(*grid):
        offs = tuple(i * c for i, c in zip(idx, cdims))
        chunk = np.zeros(cdims, dtype=a.dtype)
        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, a.shape))
        chunk[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
        yield offs, chunk.tobytes()


def _btree1_chunks(buf: _Buf, entries: List[Tuple[Tuple[int, ...], int, int]], fanout: int) -> int:
    """entries: (offsets, address, stored size); leaves of <= fanout entries, one more level above if needed."""
    nd = len(entries[0][0]) + 1
    key = lambda offs, size: struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)

    def node(level: int, items: List[Tuple[Tuple[int, ...], int, int]]) -> int:
        out = b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), UNDEF, UNDEF)
        for offs, addr, size in items:
            out += key(offs, size) + struct.pack("<Q", addr)
        out += key(tuple(0 for _ in range(nd - 1)), 0)
        return buf.add(out)

    if len(entries) <= fanout:
        return node(0, entries)
    leaves = []
    for i in range(0, len(entries), fanout):
        part = entries[i:i + fanout]
        leaves.append((part[0][0], node(0, part), part[0][2]))
    return node(1, leaves)


def _farray(buf: _Buf, chunks: List[Tuple[int, int]], filtered: bool, page_bits: int) -> int:
    """Fixed-array index over (address, stored size) in chunk order."""
    esize = 8 + (4 + 4 if filtered else 0)
    elem = lambda a, n: struct.pack("<Q", a) + (struct.pack("<II", n, 0) if filtered else b"")
    n = len(chunks)
    per_page = 1 << page_bits
    body = b""
    if n > per_page:
        npages = (n + per_page - 1) // per_page
        body += b"\xff" * ((npages + 7) // 8) + b"\0" * 4
        for pg in range(npages):
            body += b"".join(elem(a, s) for a, s in chunks[pg * per_page:(pg + 1) * per_page]) + b"\0" * 4
    else:
        body += b"".join(elem(a, s) for a, s in chunks) + b"\0" * 4
    hdr_pos = buf.add(b"\0" * (8 + 8 + 8 + 4))
    db_pos = buf.add(b"FADB" + struct.pack("<BBQ", 0, 1 if filtered else 0, hdr_pos) + body)
    buf.patch(hdr_pos, b"FAHD" + struct.pack("<BBBBQQI", 0, 1 if filtered else 0, esize, page_bits, n, db_pos, 0))
    return hdr_pos


# ---- dense link storage ---------------------------------------------------------------------------------
def _fractal_heap(buf: _Buf, objects: List[bytes], start: int = 512, width: int = 4, max_direct: int = 1024,
                  heap_bits: int = 32, checksum_blocks: bool = True) -> Tuple[int, List[bytes]]:
    """Stores the objects in a fractal heap; returns (heap header address, 7-byte heap IDs)."""
    off_bytes = (heap_bits + 7) // 8
    head = 5 + 8 + off_bytes + (4 if checksum_blocks else 0)
    log2 = lambda x: x.bit_length() - 1
    max_direct_rows = log2(max_direct) - log2(start) + 2
    row_size = lambda row: start << max(row - 1, 0)
    hdr_pos = buf.add(b"\0" * 160)
    # lay the objects into direct blocks in heap-offset order: rows of the root indirect block, then nested ones
    blocks: List[Tuple[int, int]] = []   # (heap offset, size) of direct blocks in allocation order

    def direct_sizes():
        offset = 0
        row = 0
        while True:
            for _ in range(width):
                if row < max_direct_rows:
                    yield offset, row_size(row), (row,)
                    offset += row_size(row)
                else:   # an indirect child covering row_size(row) bytes with blocks of the starting size
                    sub_rows = log2(row_size(row)) - log2(start * width) + 1
                    sub_off = offset
                    for sr in range(sub_rows):
                        for sc in range(width):
                            yield sub_off, row_size(sr), (row, sr, sc)
                            sub_off += row_size(sr)
                    offset += row_size(row)
            row += 1

    gen = direct_sizes()
    ids: List[bytes] = []
    contents: Dict[int, bytearray] = {}
    where: Dict[int, Tuple] = {}
    cur_off, cur_size, cur_path = next(gen)
    fill = head
    contents[cur_off] = bytearray(cur_size); where[cur_off] = cur_path
    for obj in objects:
        while fill + len(obj) > cur_size:
            cur_off, cur_size, cur_path = next(gen)
            contents[cur_off] = bytearray(cur_size); where[cur_off] = cur_path
            fill = head
        contents[cur_off][fill:fill + len(obj)] = obj
        ids.append(bytes([0]) + (cur_off + fill).to_bytes(off_bytes, "little") + len(obj).to_bytes(2, "little"))
        fill += len(obj)
    # write the direct blocks
    addr: Dict[int, int] = {}
    for off, data in contents.items():
        data[0:head] = b"FHDB" + struct.pack("<BQ", 0, hdr_pos) + off.to_bytes(off_bytes, "little") + (b"\0" * 4 if checksum_blocks else b"")
        addr[off] = buf.add(bytes(data))
    single = len(contents) == 1
    root_rows = 0
    root_addr = addr[0]
    if not single:
        root_rows = max(p[0] for p in where.values()) + 1
        # nested indirect blocks first
        nested: Dict[Tuple[int, int], int] = {}   # (row, column in row) -> address
        by_row: Dict[int, List[Tuple[int, Tuple]]] = {}
        for off, path in where.items():
            by_row.setdefault(path[0], []).append((off, path))
        entries = b""
        offset = 0
        for row in range(root_rows):
            for col in range(width):
                if row < max_direct_rows:
                    entries += struct.pack("<Q", addr.get(offset, UNDEF))
                else:
                    sub_rows = log2(row_size(row)) - log2(start * width) + 1
                    sub = b""
                    sub_off = offset
                    any_child = False
                    for sr in range(sub_rows):
                        for sc in range(width):
                            a = addr.get(sub_off, UNDEF)
                            any_child |= a != UNDEF
                            sub += struct.pack("<Q", a)
                            sub_off += row_size(sr)
                    if any_child:
                        a = buf.add(b"FHIB" + struct.pack("<BQ", 0, hdr_pos) + offset.to_bytes(off_bytes, "little") + sub + b"\0" * 4)
                    else:
                        a = UNDEF
                    entries += struct.pack("<Q", a)
                offset += row_size(row)
        root_addr = buf.add(b"FHIB" + struct.pack("<BQ", 0, hdr_pos) + (0).to_bytes(off_bytes, "little") + entries + b"\0" * 4)
    hdr = b"FRHP" + struct.pack("<BHHB", 0, 7, 0, 0x02 if checksum_blocks else 0) + struct.pack("<I", 4096)
    hdr += struct.pack("<QQQQQQQQQQQQ", 0, UNDEF, 0, UNDEF, 0, 0, 0, len(objects), 0, 0, 0, 0)
    hdr += struct.pack("<HQQHHQH", width, start, max_direct, heap_bits, 1, root_addr, root_rows) + b"\0" * 4
    assert len(hdr) <= 160
    buf.patch(hdr_pos, hdr)
    return hdr_pos, ids


def _btree2_names(buf: _Buf, names: List[str], ids: List[bytes], node_size: int = 512, per_leaf: int = 40) -> int:
    recs = sorted((zlib.crc32(n.encode()) & 0xFFFFFFFF, i) for n, i in zip(names, ids))
    recs = [struct.pack("<I", h) + i for h, i in recs]
    leaf = lambda rs: buf.add((b"BTLF" + struct.pack("<BB", 0, 5) + b"".join(rs) + b"\0" * 4).ljust(node_size, b"\0"))
    if len(recs) <= per_leaf:
        root, root_n, depth = leaf(recs), len(recs), 0
    else:
        # leaves of per_leaf records with one separator record between neighbours in the internal root
        leaves, seps = [], []
        i = 0
        while i < len(recs):
            part = recs[i:i + per_leaf]
            leaves.append((leaf(part), len(part)))
            i += per_leaf
            if i < len(recs):
                seps.append(recs[i]); i += 1
        if len(leaves) > len(seps) + 1: raise AssertionError
        if len(leaves) == len(seps):   # the last separator has no right neighbour: give it an empty leaf
            leaves.append((leaf([]), 0))
        body = b"BTIN" + struct.pack("<BB", 0, 5) + b"".join(seps)
        for a, n in leaves:
            body += struct.pack("<QB", a, n)           # 1 byte for the record count (<= 45 per 512-byte leaf)
        root, root_n, depth = buf.add((body + b"\0" * 4).ljust(node_size, b"\0")), len(seps), 1
    return buf.add(b"BTHD" + struct.pack("<BBIHHBBQHQI", 0, 5, node_size, 11, depth, 100, 40, root, root_n, len(recs), 0))


# ---- the file -------------------------------------------------------------------------------------------
def write_hdf5(path: str, variables: Dict[str, np.ndarray], dims: Dict[str, int], *, style: str = "v0",
               layout: str = "contiguous", chunk: Optional[Dict[str, Tuple[int, ...]]] = None,
               filters: Sequence[Tuple[int, Sequence[int]]] = (), split_headers: bool = False, dense: bool = False,
               group_fanout: int = 32, chunk_fanout: int = 64, page_bits: int = 10, hflags: int = 0x24,
               extra_vlen: Optional[str] = None, heap_kwargs: Optional[dict] = None) -> None:
    """`variables`: name -> array (file order); `dims`: NetCDF dimensions; those without a variable of the same
    name become dimension-only scales.  `layout`: compact | contiguous | chunked (v0: version-1 B-tree; v2: single
    chunk when one chunk covers the dataset, else fixed array, or implicit with layout="implicit")."""
    v0 = style == "v0"
    buf = _Buf()
    if v0:
        buf.add(b"\0" * 96)
    else:
        buf.add(b"\0" * 48)
    targets: List[Tuple[str, int]] = []

    def dataset(name: str, a: Optional[np.ndarray], shape: Tuple[int, ...], dt: np.dtype, attrs: List[Tuple[str, bytes]], lay: str):
        msgs: List[Tuple[int, bytes]] = [(0x01, _dataspace(shape, 1 if v0 else 2)),
                                         (0x03, _datatype(dt) if dt is not None else _vlen_string_datatype())]
        msgs.append((0x05, struct.pack("<BBBB", 2, 2, 0, 0) if v0 else struct.pack("<BB", 3, 0x09)))
        if a is None:
            msgs.append((0x08, struct.pack("<BBQQ", 3, 1, UNDEF, 0)))   # never written (dimension-only scale)
        elif lay == "compact":
            raw = a.tobytes()
            msgs.append((0x08, struct.pack("<BBH", 3, 0, len(raw)) + raw))
        elif lay == "contiguous" or a.ndim == 0:
            raw = a.tobytes()
            msgs.append((0x08, struct.pack("<BBQQ", 3, 1, buf.add(raw) if raw else UNDEF, len(raw))))
        else:
            cd = (chunk or {}).get(name) or tuple(max(1, (s + 1) // 2) for s in a.shape)
            pipeline = list(filters) if lay != "implicit" else []
            stored = []
            for offs, raw in _chunk_list(a, cd):
                enc = _apply_filters(raw, pipeline, a.dtype.itemsize)
                stored.append((offs, buf.add(enc), len(enc)))
            if v0:
                root = _btree1_chunks(buf, stored, chunk_fanout)
                msgs.append((0x08, struct.pack("<BBBQ", 3, 2, a.ndim + 1, root) +
                             b"".join(struct.pack("<I", c) for c in cd) + struct.pack("<I", a.dtype.itemsize)))
            else:
                dimsb = b"".join(struct.pack("<I", c) for c in cd) + struct.pack("<I", a.dtype.itemsize)
                pre = lambda flags: struct.pack("<BBBBB", 4, 2, flags, a.ndim + 1, 4) + dimsb
                if lay == "implicit":
                    raws = b"".join(raw for _, raw in _chunk_list(a, cd))
                    msgs.append((0x08, pre(0) + struct.pack("<BQ", 2, buf.add(raws))))
                elif len(stored) == 1:
                    if pipeline:
                        msgs.append((0x08, pre(0x02) + struct.pack("<BQIQ", 1, stored[0][2], 0, stored[0][1])))
                    else:
                        msgs.append((0x08, pre(0) + struct.pack("<BQ", 1, stored[0][1])))
                else:
                    fa = _farray(buf, [(p, n) for _, p, n in stored], bool(pipeline), page_bits)
                    msgs.append((0x08, pre(0) + struct.pack("<BBQ", 3, page_bits, fa)))
            if pipeline:
                msgs.append((0x0B, _filters(pipeline, 1 if v0 else 2)))
        for an, av in attrs:
            msgs.append((0x0C, _attribute(an, av, 1 if v0 else 3)))
        pos = _ohdr_v1(buf, msgs, split_headers) if v0 else _ohdr_v2(buf, msgs, split_headers, hflags)
        targets.append((name, pos))

    for dname, n in dims.items():
        if dname in variables and variables[dname].shape == (n,):
            continue
        note = ("This is a netCDF dimension but not a netCDF variable.%10d" % n).encode()
        dataset(dname, None, (n,), np.dtype("<f4"), [("CLASS", b"DIMENSION_SCALE\0"), ("NAME", note)], "contiguous")
    for vname, a in variables.items():
        a = np.asarray(a)
        attrs = []
        name = vname
        if vname in dims:
            if a.shape == (dims[vname],):
                attrs = [("CLASS", b"DIMENSION_SCALE\0"), ("NAME", vname.encode() + b"\0")]
            else:
                name = "_nc4_non_coord_" + vname
        dataset(name, a, a.shape, a.dtype, attrs, layout)
    if extra_vlen:
        dataset(extra_vlen, np.zeros(3, "V16"), (3,), None, [], "contiguous")

    targets.sort()
    if v0:
        # local heap: offset 0 is the empty string (the B-tree's first key)
        heap = bytearray(b"\0" * 8)
        name_off = {}
        for name, _ in targets:
            name_off[name] = len(heap)
            heap += _p8(name.encode() + b"\0")
        heap_data = buf.add(bytes(heap) + b"\0" * 16)
        heap_pos = buf.add(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap) + 16, len(heap), heap_data))
        snods = []
        for i in range(0, len(targets), 8):                            # group leaf node K = 4 -> 8 symbols per node
            part = targets[i:i + 8]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for name, pos in part:
                body += struct.pack("<QQII", name_off[name], pos, 0, 0) + b"\0" * 16
            snods.append((name_off[part[-1][0]], buf.add(body.ljust(8 + 8 * 40, b"\0"))))

        def tree(level: int, items: List[Tuple[int, int]]) -> int:
            out = b"TREE" + struct.pack("<BBHQQ", 0, level, len(items), UNDEF, UNDEF) + struct.pack("<Q", 0)
            for key, child in items:
                out += struct.pack("<QQ", child, key)
            return buf.add(out.ljust(24 + (2 * 16 + 1) * 8 + 2 * 16 * 8, b"\0"))

        if len(snods) <= group_fanout:
            root_tree = tree(0, snods)
        else:
            mids = []
            for i in range(0, len(snods), group_fanout):
                part = snods[i:i + group_fanout]
                mids.append((part[-1][0], tree(0, part)))
            root_tree = tree(1, mids)
        root = _ohdr_v1(buf, [(0x11, struct.pack("<QQ", root_tree, heap_pos))], False)
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, 0, UNDEF) + struct.pack("<QQII", 0, root, 0, 0) + b"\0" * 16
        assert len(sb) == 96
        buf.patch(0, sb)
        buf.patch(24 + 16, struct.pack("<Q", buf.tell()))
    else:
        msgs: List[Tuple[int, bytes]] = []
        if dense:
            links = [_link(n, p, i) for i, (n, p) in enumerate(targets)]
            heap_pos, ids = _fractal_heap(buf, links, **(heap_kwargs or {}))
            bt = _btree2_names(buf, [n for n, _ in targets], ids)
            msgs.append((0x02, struct.pack("<BBQQQ", 0, 0x01, len(targets), heap_pos, bt)))
        else:
            msgs.append((0x02, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)))
            msgs += [(0x06, _link(n, p)) for n, p in targets]
        msgs.insert(1, (0x0A, struct.pack("<BB", 0, 0)))                 # group info
        root = _ohdr_v2(buf, msgs, split_headers, hflags)
        buf.patch(0, SIG + struct.pack("<BBBBQQQQI", 2, 8, 8, 0, 0, UNDEF, buf.tell(), root, 0))
    with open(path, "wb") as f:
        f.write(bytes(buf.b))

"""Minimal HDF5 container for the explanation packs (``{data}_{MODE}_cat.h5``, reference processed/data_preprocess.py:139-143,
393-419; loader utils/batch_loader.py:120-201 via ``h5py.File(path)[name][:]``).

The image has no HDF5 library (no h5py, no libhdf5), so the container is written directly from the HDF5 File Format
Specification (version 0 superblock, version 1 object headers, "old style" group = v1 B-tree + local heap + one symbol-table
node, contiguous data layout v3, little-endian fixed/floating-point datatypes) -- the subset h5py itself emits with default
settings for plain numeric arrays.  ``read`` is an independent parser of the same subset (it also follows object-header
continuation blocks, which libhdf5-written headers use), used by the tests and by callers that load a pack without h5py.
NOT verified against libhdf5 in this container; when ``h5py`` is importable ``pack.save_pack`` uses it instead.
"""
from __future__ import annotations

import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 16, 16            # symbol-table node holds 2 * LEAF_K entries: every pack (10 datasets) fits one node


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise ValueError("h5min writes little-endian data")
    if dt.kind == "f" and dt.itemsize in (4, 8):
        exp_bits, man_bits = (8, 23) if dt.itemsize == 4 else (11, 52)
        head = struct.pack("<BBBBI", 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize)       # class 1 (float), version 1; msb of mantissa implied
        return head + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)   # class 0 (fixed point), version 1
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise ValueError(f"h5min: unsupported dtype {dt}")


def _object_header(msgs: list) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


def write(path, arrays: dict):
    """Writes every ``name -> ndarray`` of ``arrays`` as a contiguous dataset of the root group."""
    names = sorted(arrays)                                  # symbol-table entries are ordered by name (strcmp)
    if len(names) > 2 * LEAF_K:
        raise ValueError(f"h5min: at most {2 * LEAF_K} datasets per file")
    # local heap data: "" at offset 0, then the names, each null-terminated and padded to 8 bytes
    heap, name_off = b"\0" * 8, {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode() + b"\0")
    pos = 96                                                # superblock
    root_hdr = pos; pos += 16 + 8 + 16                      # root object header: prefix + one symbol-table message
    btree = pos; pos += 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    heap_hdr = pos; pos += 32
    heap_data = pos; pos += len(heap)
    snod = pos; pos += 8 + 2 * LEAF_K * 40
    hdr_at, data_at, hdrs = {}, {}, {}
    for n in names:
        a = np.ascontiguousarray(arrays[n])
        a = a.astype(a.dtype.newbyteorder("<")) if a.dtype.byteorder == ">" else a
        arrays[n] = a
        space = struct.pack("<BBBB4x", 1, a.ndim, 0, 0) + b"".join(struct.pack("<Q", d) for d in a.shape)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)             # version 2, late allocation, written if set, no value defined
        hdrs[n] = (space, _datatype(a.dtype), fill)
        size = 16 + sum(len(_msg(0, x)) for x in hdrs[n]) + len(_msg(8, b"\0" * 18))
        hdr_at[n] = pos; pos += size
    for n in names:
        pos += -pos % 8
        data_at[n] = pos; pos += arrays[n].nbytes
    eof = pos
    with open(path, "wb") as f:
        f.write(SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0))
        f.write(struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF))
        f.write(struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree, heap_hdr))       # root symbol-table entry, cache type 1
        assert f.tell() == 96
        f.write(_object_header([_msg(0x11, struct.pack("<QQ", btree, heap_hdr))]))
        # group B-tree, one leaf-level entry: key 0 = "" (heap offset 0), child 0 = the symbol-table node, key 1 = the largest name
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
        node += struct.pack("<QQQ", 0, snod, name_off[names[-1]] if names else 0)
        f.write(node.ljust(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8, b"\0"))
        f.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data))                         # free-list head 1 = no free block
        f.write(heap)
        sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n in names:
            sn += struct.pack("<QQII16x", name_off[n], hdr_at[n], 0, 0)
        f.write(sn.ljust(8 + 2 * LEAF_K * 40, b"\0"))
        for n in names:
            assert f.tell() == hdr_at[n]
            space, dtype, fill = hdrs[n]
            layout = struct.pack("<BBQQ", 3, 1, data_at[n], arrays[n].nbytes)                         # version 3, contiguous
            f.write(_object_header([_msg(1, space), _msg(3, dtype), _msg(5, fill), _msg(8, layout)]))
        for n in names:
            f.write(b"\0" * (data_at[n] - f.tell()))
            f.write(arrays[n].tobytes())
        assert f.tell() == eof


# ----------------------------------------------------------------------------------------------- reader
def _messages(buf, addr):
    """(type, data) of every message of the version-1 object header at `addr`, following continuation blocks."""
    ver, _, n_msgs, _, size = struct.unpack_from("<BBHII", buf, addr)
    if ver != 1:
        raise ValueError("h5min.read: only version-1 object headers")
    blocks, out = [(addr + 16, size)], []
    while blocks and len(out) < n_msgs:
        p, left = blocks.pop(0)
        while left >= 8 and len(out) < n_msgs:
            mtype, msize, _ = struct.unpack_from("<HHB", buf, p)
            data = bytes(buf[p + 8:p + 8 + msize])
            if mtype == 0x10:                                  # continuation: (address, length) of the next block
                blocks.append(struct.unpack_from("<QQ", data))
            out.append((mtype, data))
            p += 8 + msize; left -= 8 + msize
    return out


def _dtype_of(data):
    cls, ver = data[0] & 15, data[0] >> 4
    size = struct.unpack_from("<I", data, 4)[0]
    order = ">" if data[1] & 1 else "<"
    if cls == 1:
        return np.dtype(f"{order}f{size}")
    if cls == 0:
        return np.dtype(f"{order}{'i' if data[1] & 8 else 'u'}{size}")
    raise ValueError(f"h5min.read: datatype class {cls} (version {ver}) not supported")


def _group_entries(buf, btree, heap_data):
    """(name, object header address) of every link below the v1 group B-tree node at `btree`."""
    if buf[btree:btree + 4] != b"TREE":
        raise ValueError("h5min.read: bad B-tree signature")
    _, level, used = struct.unpack_from("<BBH", buf, btree + 4)
    out = []
    for i in range(used):
        child = struct.unpack_from("<Q", buf, btree + 24 + 8 + 16 * i)[0]
        if level > 0:
            out += _group_entries(buf, child, heap_data)
            continue
        if buf[child:child + 4] != b"SNOD":
            raise ValueError("h5min.read: bad symbol-table node signature")
        n = struct.unpack_from("<H", buf, child + 6)[0]
        for k in range(n):
            off, hdr = struct.unpack_from("<QQ", buf, child + 8 + 40 * k)
            end = buf.index(b"\0", heap_data + off)
            out.append((bytes(buf[heap_data + off:end]).decode(), hdr))
    return out


def read(path) -> dict:
    """name -> ndarray for every contiguous numeric dataset of the root group."""
    with open(path, "rb") as fh:
        b = fh.read()
    if b[:8] != SIG or b[8] != 0 or b[13] != 8 or b[14] != 8:
        raise ValueError("h5min.read: need a version-0 superblock with 8-byte offsets and lengths")
    base = struct.unpack_from("<Q", b, 24)[0]
    if base != 0:
        raise ValueError("h5min.read: non-zero base address")
    root_hdr = struct.unpack_from("<Q", b, 56 + 8)[0]
    sym = [d for t, d in _messages(b, root_hdr) if t == 0x11]
    if not sym:
        raise ValueError("h5min.read: root group has no symbol-table message")
    btree, heap_hdr = struct.unpack_from("<QQ", sym[0])
    if b[heap_hdr:heap_hdr + 4] != b"HEAP":
        raise ValueError("h5min.read: bad local-heap signature")
    heap_data = struct.unpack_from("<Q", b, heap_hdr + 24)[0]
    out = {}
    for name, hdr in _group_entries(b, btree, heap_data):
        shape = dt = addr = nbytes = None
        for t, d in _messages(b, hdr):
            if t == 1:
                rank = d[1]
                shape = struct.unpack_from(f"<{rank}Q", d, 8 if d[0] == 1 else 4)
            elif t == 3:
                dt = _dtype_of(d)
            elif t == 8:
                if d[0] != 3 or d[1] != 1:
                    raise ValueError(f"h5min.read: dataset {name} is not contiguous (layout version {d[0]}, class {d[1]})")
                addr, nbytes = struct.unpack_from("<QQ", d, 2)
        if shape is None or dt is None or addr is None:
            continue                                           # not a dataset (sub-group)
        count = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        out[name] = np.frombuffer(b, dtype=dt, count=count, offset=addr).reshape(shape).copy() if addr != UNDEF else np.zeros(shape, dt)
    return out

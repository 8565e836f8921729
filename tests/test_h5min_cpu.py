"""The minimal HDF5 container (tempme_b200/h5min.py): structural checks against the HDF5 file-format specification and a
write -> read round trip of a pack-shaped set of arrays.  No HDF5 library exists in this image; when h5py is importable the file
is also opened with it."""
import struct

import numpy as np
import pytest

from tempme_b200 import h5min


def pack_like(rng):
    a = {f"subgraph_{r}_{l}": rng.standard_normal((7, 3 * 5 ** (l + 1))) for r in ("src", "tgt", "bgd") for l in (0, 1)}
    a.update({f"walks_{r}_new": rng.standard_normal((7, 15, 14)) for r in ("src", "tgt", "bgd")})
    a["dst_fake"] = rng.integers(0, 1000, 7).astype(np.int64)
    return a


def test_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    arrays = pack_like(rng)
    arrays["f32"] = rng.standard_normal((3, 4)).astype(np.float32)
    arrays["i32"] = np.arange(-5, 5, dtype=np.int32)
    arrays["u8"] = np.arange(7, dtype=np.uint8)
    arrays["scalar_like"] = np.array([3.5])
    arrays["empty"] = np.zeros((0, 4))
    path = tmp_path / "x_cat.h5"
    h5min.write(str(path), dict(arrays))
    raw = path.read_bytes()
    # superblock v0 (spec III.A): signature, versions 0, 8-byte offsets / lengths, end-of-file address = file size
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8:13] == b"\0\0\0\0\0" and raw[13] == 8 and raw[14] == 8
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)
    root_hdr, cache = struct.unpack_from("<Q", raw, 64)[0], struct.unpack_from("<I", raw, 72)[0]
    assert cache == 1 and root_hdr % 8 == 0 and raw[root_hdr] == 1
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert raw[btree:btree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP"
    snod = struct.unpack_from("<Q", raw, btree + 32)[0]
    assert raw[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", raw, snod + 6)[0] == len(arrays)
    got = h5min.read(str(path))
    assert sorted(got) == sorted(arrays)                              # names come back in strcmp order
    for k, v in arrays.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    with pytest.raises(ValueError):
        h5min.write(str(tmp_path / "y.h5"), {f"d{i}": np.zeros(1) for i in range(40)})
    with pytest.raises(ValueError):
        h5min.write(str(tmp_path / "z.h5"), {"c": np.zeros(2, np.complex64)})


def test_pack_files_load_like_the_reference_loader(tmp_path):
    """save_pack -> the slicing utils/batch_loader.load_subgraph_margin does on the opened file (:120-201): file[name][:]."""
    from tempme_b200 import pack as pk
    rng = np.random.default_rng(1)
    arrays = pack_like(rng)
    edge = rng.standard_normal((3, 7, 15, 3, 3))
    cat_path, edge_path = pk.save_pack(arrays, edge, str(tmp_path), "unit", "test")
    assert cat_path.endswith("unit_test_cat.h5") and edge_path.endswith("unit_test_edge.npy")
    f = pk.load_pack(cat_path)
    for k in pk.PACK_KEYS:
        assert np.array_equal(f[k][:], arrays[k])
    assert np.array_equal(np.load(edge_path), edge)


def test_h5py_opens_the_file_when_available(tmp_path):
    h5py = pytest.importorskip("h5py")                                # not in this image; runs wherever h5py exists
    arrays = pack_like(np.random.default_rng(2))
    h5min.write(str(tmp_path / "m.h5"), dict(arrays))
    with h5py.File(str(tmp_path / "m.h5"), "r") as hf:
        for k in arrays:
            assert np.array_equal(hf[k][:], arrays[k])

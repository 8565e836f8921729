"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see tempme_oracle.c header).

ctypes front-end for ``libtempme_oracle.so`` (plain-C restatement of the reference's
``utils/graph.py`` / ``utils/null_model.py`` / ``processed/data_preprocess.py`` hot path)
plus ``oracle.encoder`` (numpy fp32 restatement of ``models/explainer.py`` ``TempME.forward``).

Nothing under ``tempme_b200/`` may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtempme_oracle.so")

ORC_ERR_EIDX_NOT_FOUND = -2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tempme_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libtempme_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        p = C.c_void_p
        L.orc_graph_build.argtypes = [C.c_int64, C.c_int64, p, p, p, p, C.POINTER(p)]
        L.orc_graph_free.argtypes = [p]
        L.orc_graph_sizes.argtypes = [p, p, p, p]
        L.orc_graph_export.argtypes = [p, p, p, p, p]
        L.orc_dict_get.argtypes = [p, C.c_int64, C.c_int32, p]
        L.orc_find_before.argtypes = [p, C.c_int64, C.c_double, C.c_int, C.c_int32, p, p]
        L.orc_find_before_batch.argtypes = [p, C.c_int64, p, p, p, p, p]
        L.orc_sample_hop.argtypes = [p, C.c_int64, p, p, p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint64, p, p, p]
        L.orc_sample_walks.argtypes = [p, C.c_int64, C.c_int, C.c_int, p, p, p, p, C.c_uint64, C.c_uint64,
                                       p, p, p, p, p]
        L.orc_class_hist_null.argtypes = [C.c_int64, p, p]
        L.orc_class_ids_prep.argtypes = [C.c_int64, p, p, p]
        L.orc_edge_identity.argtypes = [C.c_int64, C.c_int64, p, p]
        L.orc_edge_identity.restype = None
        L.orc_philox4x32_10.argtypes = [C.c_uint32, C.c_uint32, p, p]
        L.orc_philox4x32_10.restype = None
        L.orc_draw_index.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64]
        L.orc_draw_index.restype = C.c_uint64
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def philox4x32_10(key, ctr):
    out = np.zeros(4, np.uint32)
    c = np.asarray(ctr, np.uint32)
    lib().orc_philox4x32_10(int(key[0]), int(key[1]), _ptr(c), _ptr(out))
    return out


def draw_index(seed, stage, row, slot, L):
    return int(lib().orc_draw_index(seed, stage, row, slot, L))


def entries_from_events(src, dst, eidx, ts):
    """Flattened adj_list exactly as the callers build it (temp_exp_main.py:135-144):
    every event is appended to adj[src] as (dst, e, t) and then to adj[dst] as (src, e, t)."""
    src = np.asarray(src); dst = np.asarray(dst)
    n = len(src)
    node = np.empty(2 * n, np.int32); nbr = np.empty(2 * n, np.int32)
    node[0::2] = src; node[1::2] = dst
    nbr[0::2] = dst; nbr[1::2] = src
    e = np.repeat(np.asarray(eidx, np.int32), 2)
    t = np.repeat(np.asarray(ts, np.float64), 2)
    return node, nbr, e, t


class OracleGraph:
    """NeighborFinder restated (utils/graph.py:12-476)."""

    def __init__(self, n_nodes, entry_node, entry_nbr, entry_eidx, entry_ts):
        en, eb, ee, et = _c(entry_node, np.int32), _c(entry_nbr, np.int32), _c(entry_eidx, np.int32), _c(entry_ts, np.float64)
        h = C.c_void_p()
        rc = lib().orc_graph_build(int(n_nodes), len(en), _ptr(en), _ptr(eb), _ptr(ee), _ptr(et), C.byref(h))
        if rc:
            raise RuntimeError(f"orc_graph_build failed: {rc}")
        self._h = h
        self.n_nodes = int(n_nodes)
        self.n_entries = len(en)

    @classmethod
    def from_events(cls, n_nodes, src, dst, eidx, ts):
        return cls(n_nodes, *entries_from_events(src, dst, eidx, ts))

    def __del__(self):
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.orc_graph_free(self._h)
            self._h = None

    def export(self):
        off = np.zeros(self.n_nodes + 1, np.int64)
        nbr = np.zeros(self.n_entries, np.int32); eidx = np.zeros(self.n_entries, np.int32)
        ts = np.zeros(self.n_entries, np.float64)
        lib().orc_graph_export(self._h, _ptr(off), _ptr(nbr), _ptr(eidx), _ptr(ts))
        return off, nbr, eidx, ts

    def dict_get(self, node, e):
        v = C.c_int64()
        ok = lib().orc_dict_get(self._h, int(node), int(e), C.byref(v))
        return v.value if ok else None

    def find_before_batch(self, node, cut_time=None, eidx=None):
        node = _c(node, np.int32); R = len(node)
        ct = _c(cut_time, np.float64); e = _c(eidx, np.int32)
        start = np.zeros(R, np.int64); cut = np.zeros(R, np.int64)
        rc = lib().orc_find_before_batch(self._h, R, _ptr(node), _ptr(ct), _ptr(e), _ptr(start), _ptr(cut))
        if rc == ORC_ERR_EIDX_NOT_FOUND:
            raise IndexError("e_idx not found in edge list")
        if rc:
            raise RuntimeError(f"orc_find_before_batch failed: {rc}")
        return start, cut

    def sample_hop(self, node, cut_time, n, eidx=None, seed=0, stage=0, row_offset=0):
        node = _c(node, np.int32); R = len(node)
        ct = _c(cut_time, np.float64); e = _c(eidx, np.int32)
        o_node = np.zeros((R, n), np.int32); o_eidx = np.zeros((R, n), np.int32); o_ts = np.zeros((R, n), np.float32)
        rc = lib().orc_sample_hop(self._h, R, _ptr(node), _ptr(ct), _ptr(e), n, seed, stage, row_offset,
                                  _ptr(o_node), _ptr(o_eidx), _ptr(o_ts))
        if rc == ORC_ERR_EIDX_NOT_FOUND:
            raise IndexError("e_idx not found in edge list")
        if rc:
            raise RuntimeError(f"orc_sample_hop failed: {rc}")
        return o_node, o_eidx, o_ts

    def find_k_hop(self, k, node, cut_time, n, eidx=None, seed=0, row_offset=0):
        """find_k_hop, utils/graph.py:233-262."""
        recs = ([], [], [])
        B = len(node)
        x, y, z = self.sample_hop(node, cut_time, n, eidx, seed, 0, row_offset)
        for r, v in zip(recs, (x, y, z)):
            r.append(v)
        for layer in range(1, k):
            pn, pe, pt = recs[0][-1].reshape(-1), recs[1][-1].reshape(-1), recs[2][-1].reshape(-1)
            x, y, z = self.sample_hop(pn, pt.astype(np.float64), n, pe, seed, layer, row_offset * (n ** layer))
            for r, v in zip(recs, (x, y, z)):
                r.append(v.reshape(B, -1))
        return recs

    def sample_walks(self, root, h1_node, h1_eidx, h1_ts, N2, seed=0, row_offset=0, want_scanned=False):
        root = _c(root, np.int32); B = len(root)
        h1n, h1e, h1t = _c(h1_node, np.int32), _c(h1_eidx, np.int32), _c(h1_ts, np.float32)
        n = h1n.shape[1]; W = n * N2
        nodes = np.zeros((B, W, 6), np.int32); eidx = np.zeros((B, W, 3), np.int32)
        t = np.zeros((B, W, 3), np.float32); anony = np.zeros((B, W, 3), np.int32)
        scanned = np.zeros(B * W, np.int64) if want_scanned else None
        rc = lib().orc_sample_walks(self._h, B, n, N2, _ptr(root), _ptr(h1n), _ptr(h1e), _ptr(h1t), seed, row_offset,
                                    _ptr(nodes), _ptr(eidx), _ptr(t), _ptr(anony), _ptr(scanned))
        if rc:
            raise RuntimeError(f"orc_sample_walks failed: {rc}")
        return (nodes, eidx, t, anony, scanned) if want_scanned else (nodes, eidx, t, anony)


def class_hist_null(anony):
    a = _c(anony, np.int32).reshape(-1, 3)
    h = np.zeros(12, np.int64)
    if lib().orc_class_hist_null(len(a), _ptr(a), _ptr(h)):
        raise KeyError("anonymized row is not one of the 12 motif classes")
    return h


def class_ids_prep(anony):
    a = _c(anony, np.int32)
    flat = a.reshape(-1, 3)
    cat = np.zeros(len(flat), np.int32); h = np.zeros(12, np.int64)
    if lib().orc_class_ids_prep(len(flat), _ptr(flat), _ptr(cat), _ptr(h)):
        raise KeyError("anonymized row is not one of the 12 motif classes")
    return cat.reshape(a.shape[:-1]), h


def edge_identity(eidx):
    e = _c(eidx, np.int32)
    B, W, _ = e.shape
    out = np.zeros((B, W, 3, 3), np.float64)
    lib().orc_edge_identity(B, W, _ptr(e), _ptr(out))
    return out

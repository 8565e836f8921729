"""Harness that runs the UNMODIFIED reference (imported from /root/reference) under the
deterministic-draw contract.  Used only in the build container (the reference does not travel to
the GPU box): by tests/golden/make_golden.py to produce the committed fixtures, and by the
``live reference`` CPU tests, which skip when /root/reference is absent.

No reference file is modified or copied: ``numpy.random.randint`` is swapped for a function that
reads the caller's loop index from the calling frame and returns the contract's Philox draws
(SURVEY.md App. B).  The Philox here is a third, pure-Python implementation (the others are in
oracle/tempme_oracle.c and tempme_b200/csrc), so the three can be checked against each other.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np

REF = os.environ.get("TEMPME_REFERENCE", "/root/reference")
M32 = 0xFFFFFFFF


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "utils"))


def philox4x32_10(key, ctr):
    k0, k1 = key
    c0, c1, c2, c3 = ctr
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def draw_index(seed, stage, row, slot, L):
    o = philox4x32_10((seed & M32, (seed >> 32) & M32), (slot >> 1, row & M32, (row >> 32) & M32, stage))
    r = (o[3] << 32 | o[2]) if (slot & 1) else (o[1] << 32 | o[0])
    return (r * L) >> 64


def import_reference():
    """Import the reference's ``utils`` (and ``models.explainer`` with a torch_scatter stand-in)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "torch_scatter" not in sys.modules:
        import torch

        def scatter(src, index, dim=-1, dim_size=None, reduce="sum"):
            red = {"max": "amax", "mean": "mean", "sum": "sum", "add": "sum", "min": "amin"}[reduce]
            shape = list(src.shape)
            shape[dim] = dim_size if dim_size is not None else int(index.max()) + 1
            out = torch.zeros(shape, dtype=src.dtype, device=src.device)
            return out.scatter_reduce(dim, index, src, reduce=red, include_self=False)

        m = types.ModuleType("torch_scatter")
        m.scatter = scatter
        sys.modules["torch_scatter"] = m
    import utils.graph as rgraph  # noqa: E402  (the reference's module)
    return rgraph


class DrawShim:
    """Replaces numpy.random.randint while the reference's sampling loops run.

    seed for top-level call number c (find_k_hop / get_temporal_neighbor / find_k_walks, counted by
    the wrappers installed in ``patched``) is ``base_seed + c``; ``row_offset`` is the global index
    of the first root of the batch (0 unless a shard is being emulated).
    """

    def __init__(self, base_seed=0, row_offset=0):
        self.base_seed = base_seed
        self.row_offset = row_offset
        self.calls = 0
        self.seed = base_seed
        self._real = np.random.randint
        self.n_draw_calls = 0

    def randint(self, low, high=None, size=None, dtype=int):
        f = sys._getframe(1)
        name = f.f_code.co_name
        if name not in ("get_temporal_neighbor", "get_next_step", "get_final_step"):
            return self._real(low, high, size, dtype)
        assert low == 0 and high is not None
        i = f.f_locals["i"]
        if name == "get_temporal_neighbor":
            level, caller = 0, f.f_back
            for _ in range(2):  # skip the call-counting wrapper installed by ``patched``
                if caller is not None and caller.f_code.co_name == "find_k_hop":
                    level = caller.f_locals.get("layer_i", 0)
                    break
                caller = caller.f_back if caller is not None else None
            n = f.f_locals["num_neighbor"]
            stage, row = level, self.row_offset * (n ** level) + i
        elif name == "get_next_step":
            stage, row = 16, self.row_offset * f.f_locals["degree"] + i
        else:
            kw = f.f_back.f_locals  # find_k_walks frame: degree * num_neighbors walks per root
            W = kw["degree"] * kw["num_neighbors"]
            stage, row = 17, self.row_offset * W + i
        self.n_draw_calls += 1
        return np.array([draw_index(self.seed, stage, row, s, int(high)) for s in range(int(size))], dtype=np.int64)

    @contextlib.contextmanager
    def patched(self, rgraph):
        NF = rgraph.NeighborFinder
        orig = {k: getattr(NF, k) for k in ("find_k_hop", "find_k_walks", "get_temporal_neighbor")}
        shim = self
        depth = {"d": 0}

        def wrap(fn):
            def inner(this, *a, **k):
                top = depth["d"] == 0
                if top:
                    shim.seed = shim.base_seed + shim.calls
                    shim.calls += 1
                depth["d"] += 1
                try:
                    return fn(this, *a, **k)
                finally:
                    depth["d"] -= 1
            return inner

        for k, fn in orig.items():
            setattr(NF, k, wrap(fn))
        np.random.randint = self.randint
        try:
            yield self
        finally:
            np.random.randint = self._real
            for k, fn in orig.items():
                setattr(NF, k, fn)


def adj_list_from_events(n_nodes, src, dst, eidx, ts):
    """temp_exp_main.py:135-144 (python scalars, as pandas .values iteration yields numpy scalars)."""
    adj = [[] for _ in range(n_nodes)]
    for s, d, e, t in zip(src, dst, eidx, ts):
        adj[int(s)].append((int(d), int(e), float(t)))
        adj[int(d)].append((int(s), int(e), float(t)))
    return adj


class DrawRecorder:
    """Leaves the reference's own MT19937 stream in place and only RECORDS what numpy.random.randint returned to
    the sampling loops: {(function, hop level): {row i: raw draws}}.  Feeding these arrays back through the CUDA
    path's INJECTED mode must reproduce the reference's outputs bit for bit."""

    def __init__(self):
        self._real = np.random.randint
        self.log = {}

    def randint(self, low, high=None, size=None, dtype=int):
        out = self._real(low, high, size, dtype)
        f = sys._getframe(1)
        name = f.f_code.co_name
        if name in ("get_temporal_neighbor", "get_next_step", "get_final_step"):
            level = 0
            if name == "get_temporal_neighbor" and f.f_back is not None and f.f_back.f_code.co_name == "find_k_hop":
                level = f.f_back.f_locals.get("layer_i", 0)
            self.log.setdefault((name, level), {})[int(f.f_locals["i"])] = np.array(out, dtype=np.int64).reshape(-1)
        return out

    @contextlib.contextmanager
    def patched(self):
        np.random.randint = self.randint
        try:
            yield self
        finally:
            np.random.randint = self._real

    def dense(self, key, rows, width):
        a = np.zeros((rows, width), np.int64)
        for i, v in self.log.get(key, {}).items():
            a[i, :len(v)] = v
        return a

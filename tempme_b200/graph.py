"""``NeighborFinder`` with the reference's Python surface (reference utils/graph.py:12-476) on top of
the device-resident CSR of libtempme_b200.

Host API (numpy in, numpy out, reference dtypes/shapes): ``find_before``, ``get_temporal_neighbor``,
``find_k_hop``, ``find_k_walks``, ``find_before_walk``, ``get_next_step`` and ``get_final_step`` (the last two are
views of the one fused walk kernel: step 2 alone / step 3 on given second events).  Device API (torch CUDA tensors in/out, no
host round trip): ``sample_hop_device``, ``find_k_hop_device``, ``find_k_walks_device``.

Randomness: the reference draws from numpy's global MT19937 stream.  Here every top-level sampling
call number ``c`` uses the counter-based Philox stream keyed by ``seed + c`` (DESIGN.md "RNG"), so
results do not depend on how queries are split across GPUs; ``inject=`` replays recorded indices.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import TM_EIDX_NONE, check, lib, ptr

PRECISION = 5


def _flatten_adj_list(adj_list):
    """adj_list: List[List[(nbr, eidx, ts)]] -> entry arrays in insertion order."""
    lens = np.fromiter((len(a) for a in adj_list), dtype=np.int64, count=len(adj_list))
    total = int(lens.sum())
    node = np.repeat(np.arange(len(adj_list), dtype=np.int32), lens)
    nbr = np.empty(total, np.int32); eidx = np.empty(total, np.int32); ts = np.empty(total, np.float64)
    p = 0
    for a in adj_list:
        if a:
            arr = np.asarray(a, dtype=np.float64)          # ids are exact in float64 up to 2^53
            k = len(a)
            nbr[p:p + k] = arr[:, 0]; eidx[p:p + k] = arr[:, 1]; ts[p:p + k] = arr[:, 2]
            p += k
    return node, nbr, eidx, ts


class NeighborFinder:
    def __init__(self, adj_list, bias=0, ts_precision=PRECISION, use_cache=False, sample_method='multinomial',
                 device=None, seed=None, _entries=None, _events=None):
        if not math.isclose(bias, 0) or sample_method != 'multinomial':
            # the reference's callers never leave the defaults; those branches are not built
            raise NotImplementedError("tempme_b200.NeighborFinder supports bias=0, sample_method='multinomial' only")
        if not torch.cuda.is_available():
            raise RuntimeError("tempme_b200.NeighborFinder needs a CUDA device (no CPU fallback)")
        self.bias = bias
        self.ts_precision = ts_precision
        self.use_cache = use_cache
        self.cache = {}
        self.sample_method = sample_method
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        h = C.c_void_p()
        if _events is not None:        # event list: expanded to the two adjacency entries per event on the device
            n_nodes, (src, dst, eidx, ts) = _events
            check(lib().tm_graph_create_from_events(n_nodes, len(src), ptr(src), ptr(dst), ptr(eidx), ptr(ts), dev.index, C.byref(h)),
                  "tm_graph_create_from_events")
            self.n_entries = 2 * len(src)
        else:
            n_nodes, (node, nbr, eidx, ts) = (len(adj_list), _flatten_adj_list(adj_list)) if _entries is None else _entries
            check(lib().tm_graph_create(n_nodes, len(node), ptr(node), ptr(nbr), ptr(eidx), ptr(ts), dev.index, C.byref(h)),
                  "tm_graph_create")
            self.n_entries = len(node)
        self._h = h
        self.n_nodes = n_nodes
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.calls = 0
        self._host = None
        self._dict = None
        self._err = torch.zeros(1, dtype=torch.int32, device=dev)

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_events(cls, n_nodes, src, dst, eidx, ts, device=None, seed=None):
        """Graph of an event list; equals NeighborFinder(adj_list) for the adj_list the reference's
        callers build (each event appended to both endpoints, temp_exp_main.py:135-144)."""
        c = lambda a, dt: np.ascontiguousarray(np.asarray(a), dtype=dt)
        return cls(None, device=device, seed=seed, _events=(int(n_nodes), (c(src, np.int32), c(dst, np.int32), c(eidx, np.int32), c(ts, np.float64))))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and _lib._lib is not None:
            _lib._lib.tm_graph_destroy(h)
            self._h = None

    def device_bytes(self):
        v = [C.c_int64() for _ in range(4)]
        check(lib().tm_graph_sizes(self._h, *[C.byref(x) for x in v]), "tm_graph_sizes")
        return v[3].value

    # ------------------------------------------------------------------ reference attribute surface
    def _export(self):
        if self._host is None:
            off = np.zeros(self.n_nodes + 1, np.int64)
            nbr = np.zeros(self.n_entries, np.int32); e = np.zeros(self.n_entries, np.int32)
            ts = np.zeros(self.n_entries, np.float64)
            check(lib().tm_graph_export(self._h, ptr(off), ptr(nbr), ptr(e), ptr(ts)), "tm_graph_export")
            self._host = (off, nbr.astype(np.int64), e.astype(np.int64), ts)
        return self._host

    @property
    def off_set_l(self):
        return self._export()[0]

    @property
    def node_idx_l(self):
        return self._export()[1]

    @property
    def edge_idx_l(self):
        return self._export()[2]

    @property
    def node_ts_l(self):
        return self._export()[3]

    @property
    def binary_prob_l(self):
        # compute_binary_prob with bias == 0 (graph.py:68-75): exp(0)/cumsum(1) = 1/(position+1)
        off = self.off_set_l
        pos = np.arange(self.n_entries) - np.repeat(off[:-1], np.diff(off))
        return 1.0 / (pos + 1.0)

    def edge_table(self):
        v = [C.c_int64() for _ in range(4)]
        check(lib().tm_graph_sizes(self._h, *[C.byref(x) for x in v]), "tm_graph_sizes")
        tab = np.full((v[2].value + 1, 4), -1, np.int32)
        check(lib().tm_graph_export_edge_table(self._h, ptr(tab)), "tm_graph_export_edge_table")
        return tab

    def secondary_index(self):
        """Per node the sorted keys (neighbour << 32 | position) of the id filter of step 3 (uint64 [n_entries]); for tests of the build."""
        k = np.zeros(self.n_entries, np.uint64)
        check(lib().tm_graph_export_skey(self._h, ptr(k)), "tm_graph_export_skey")
        return k

    @property
    def nodeedge2idx(self):
        """{node: {e_idx: cut}} rebuilt from the device edge table (graph.py:56,77-101).  Values are the
        effective prefix lengths, i.e. already passed through python's ``[:cut]`` slice semantics."""
        if self._dict is None:
            d = {v: {} for v in range(self.n_nodes)}
            tab = self.edge_table()
            for e in np.nonzero(tab[:, 0] >= 0)[0]:
                a, b, ca, cb = tab[e]
                d[int(a)][int(e)] = int(ca)
                if b >= 0:
                    d[int(b)][int(e)] = int(cb)
            self._dict = d
        return self._dict

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, a, dtype):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(np.asarray(a)), device="cpu").to(dtype).to(self.device, non_blocking=True)

    def _next_seed(self, seed):
        if seed is not None:
            return int(seed) & (2 ** 64 - 1)
        s = (self.seed + self.calls) & (2 ** 64 - 1)
        self.calls += 1
        return s

    def _raise_if_err(self, what):
        row = int(self._err.item())
        if row:
            self._err.zero_()
            raise IndexError(f"{what}: e_idx not found in edge list (or node id out of range) at row {row - 1}")

    def check_errors(self, what="device call"):
        """Raises the IndexError the reference raises (graph.py:134-135) if any device-API call since the last check saw an
        e_idx that is not in its node's list or a node id out of range; clears the flag.  Synchronises the current stream.
        The host wrappers call this themselves; callers of the ``*_device`` methods call it where they synchronise anyway."""
        self._raise_if_err(what)

    def _clear_err(self):
        """Start of a host wrapper: a flag left by an earlier, unchecked device call must not be blamed on this one."""
        self._err.zero_()

    # ------------------------------------------------------------------ find_before (graph.py:103-146)
    def find_before_batch_device(self, node, cut_time=None, e_idx=None):
        node = self._dev(node, torch.int32)
        R = node.numel()
        ct = self._dev(cut_time, torch.float64)
        e = self._dev(e_idx, torch.int32)
        start = torch.empty(R, dtype=torch.int64, device=self.device)
        cut = torch.empty(R, dtype=torch.int32, device=self.device)
        check(lib().tm_find_before_batch(self._h, R, ptr(node), ptr(ct), ptr(e), ptr(start), ptr(cut), ptr(self._err),
                                         self._stream()), "tm_find_before_batch")
        return start, cut

    def find_before(self, src_idx, cut_time, e_idx=None, return_binary_prob=False):
        """Returns views into the exported CSR arrays, like the reference.  One packed 32-byte H2D, the lookup kernel, one packed
        D2H and a single stream synchronisation per call (scalar callers: TGN / GraphMixer style per-event lookups)."""
        if getattr(self, "_fb_pin", None) is None:
            self._fb_pin = torch.zeros(4, dtype=torch.int64).pin_memory()      # f64 cut_time | i32 node, i32 e_idx | i64 start | i32 cut, i32 err
            self._fb_dev = torch.zeros(4, dtype=torch.int64, device=self.device)
        h = self._fb_pin.numpy()
        h.view(np.float64)[0] = float(cut_time)
        h32 = h.view(np.int32)
        h32[2] = int(src_idx)
        h32[3] = TM_EIDX_NONE if e_idx is None else int(e_idx)
        h[2] = 0; h[3] = 0
        with torch.cuda.device(self.device):
            self._fb_dev.copy_(self._fb_pin, non_blocking=True)
            base = self._fb_dev.data_ptr()
            check(lib().tm_find_before_batch(self._h, 1, C.c_void_p(base + 8), C.c_void_p(base), C.c_void_p(base + 12), C.c_void_p(base + 16),
                                             C.c_void_p(base + 24), C.c_void_p(base + 28), self._stream()), "tm_find_before_batch")
            self._fb_pin.copy_(self._fb_dev, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        s, c, row = int(h[2]), int(h32[6]), int(h32[7])
        if row:
            raise IndexError('e_idx {} not found in edge list of {}'.format(e_idx, src_idx))
        off, nbr, e, ts = self._export()
        prob = self.binary_prob_l[s:s + c] if return_binary_prob else None
        return nbr[s:s + c], e[s:s + c], ts[s:s + c], prob

    # ------------------------------------------------------------------ get_temporal_neighbor (graph.py:197-231)
    def sample_hop_device(self, node, cut_time, num_neighbor, e_idx=None, seed=0, stage=0, row_offset=0, inject=None, out=None):
        """out: optional preallocated (node i32 [R,n], eidx i32 [R,n], ts f32 [R,n]) to write into (no allocation on the call)."""
        node = self._dev(node, torch.int32)
        R = node.numel()
        ct = self._dev(cut_time, torch.float64) if cut_time is not None else None
        e = self._dev(e_idx, torch.int32)
        inj = self._dev(inject, torch.int64).to(torch.int32) if inject is not None else None
        if out is not None:
            o_node, o_eidx, o_ts = out
            assert o_node.shape == (R, num_neighbor) and o_node.dtype == torch.int32 and o_ts.dtype == torch.float32 and o_node.is_contiguous()
        else:
            o_node = torch.empty((R, num_neighbor), dtype=torch.int32, device=self.device)
            o_eidx = torch.empty_like(o_node)
            o_ts = torch.empty((R, num_neighbor), dtype=torch.float32, device=self.device)
        check(lib().tm_sample_hop(self._h, R, ptr(node), ptr(ct), ptr(e), int(num_neighbor), seed, stage, row_offset,
                                  ptr(inj), ptr(o_node), ptr(o_eidx), ptr(o_ts), ptr(self._err), self._stream()),
              "tm_sample_hop")
        return o_node, o_eidx, o_ts

    def get_temporal_neighbor(self, src_idx_l, cut_time_l, num_neighbor, e_idx_l=None, seed=None, row_offset=0, inject=None):
        assert (len(src_idx_l) == len(cut_time_l))
        self._clear_err()
        s = self._next_seed(seed)
        ct = None if e_idx_l is not None else cut_time_l
        o = self.sample_hop_device(src_idx_l, ct, num_neighbor, e_idx_l, s, 0, row_offset, inject)
        out = tuple(x.cpu().numpy() for x in o)
        self._raise_if_err("get_temporal_neighbor")
        return out

    # ------------------------------------------------------------------ find_k_hop (graph.py:233-262)
    def find_k_hop_device(self, k, src_idx_l, cut_time_l, num_neighbors, e_idx_l=None, seed=None, row_offset=0, inject=None):
        """inject: optional list (one per hop) of recorded index arrays."""
        if k == 0:
            return ([], [], [])
        s = self._next_seed(seed)
        n = int(num_neighbors)
        B = len(src_idx_l)
        ct = None if e_idx_l is not None else cut_time_l
        if inject:                      # replaying recorded indices: hop by hop
            x, y, z = self.sample_hop_device(src_idx_l, ct, n, e_idx_l, s, 0, row_offset, inject[0])
            recs = ([x], [y], [z])
            for layer in range(1, k):
                pn, pe = recs[0][-1].reshape(-1), recs[1][-1].reshape(-1)
                # deeper hops look the window up by e_idx; the (float32) time is unused (graph.py:247-250)
                x, y, z = self.sample_hop_device(pn, None, n, pe, s, layer, row_offset * (n ** layer), inject[layer])
                for r, v in zip(recs, (x, y, z)):
                    r.append(v.view(B, n ** (layer + 1)))
            return recs
        root = self._dev(src_idx_l, torch.int32)
        ctd = self._dev(ct, torch.float64) if ct is not None else None
        ed = self._dev(e_idx_l, torch.int32)
        recs = ([], [], [])
        for layer in range(k):
            w = n ** (layer + 1)
            recs[0].append(torch.empty((B, w), dtype=torch.int32, device=self.device))
            recs[1].append(torch.empty((B, w), dtype=torch.int32, device=self.device))
            recs[2].append(torch.empty((B, w), dtype=torch.float32, device=self.device))
        arr = lambda ts: (C.c_void_p * k)(*[t.data_ptr() for t in ts])
        check(lib().tm_sample_khop(self._h, B, int(k), n, ptr(root), ptr(ctd), ptr(ed), s, row_offset, arr(recs[0]), arr(recs[1]), arr(recs[2]),
                                   ptr(self._err), self._stream()), "tm_sample_khop")
        return recs

    def find_k_hop(self, k, src_idx_l, cut_time_l, num_neighbors, e_idx_l=None, seed=None, row_offset=0, inject=None):
        self._clear_err()
        recs = self.find_k_hop_device(k, src_idx_l, cut_time_l, num_neighbors, e_idx_l, seed, row_offset, inject)
        out = tuple([t.cpu().numpy() for t in r] for r in recs)
        self._raise_if_err("find_k_hop")
        return out

    # ------------------------------------------------------------------ find_k_walks (graph.py:265-476)
    def find_k_walks_device(self, degree, src_idx_l, num_neighbors, subgraph_src, seed=None, row_offset=0,
                            inject2=None, inject3=None, want_anony=True, want_cat=True, hist_null=None, hist_prep=None,
                            scanned=None, out=None):
        """out: optional preallocated (nodes i32 [B,W,6], eidx i32 [B,W,3], t f32 [B,W,3], cat u8 [B,W]) to write into."""
        s = self._next_seed(seed)
        root = self._dev(src_idx_l, torch.int32)
        B = root.numel()
        n, N2 = int(degree), int(num_neighbors)
        h1n = self._dev(subgraph_src[0][0], torch.int32); h1e = self._dev(subgraph_src[1][0], torch.int32)
        h1t = self._dev(subgraph_src[2][0], torch.float32)
        assert h1n.shape == (B, n)
        W = n * N2
        dev = self.device
        if out is not None:
            nodes, eidx, t, cat = out
            assert nodes.shape == (B, W, 6) and eidx.shape == (B, W, 3) and t.shape == (B, W, 3) and nodes.is_contiguous()
        else:
            nodes = torch.empty((B, W, 6), dtype=torch.int32, device=dev)
            eidx = torch.empty((B, W, 3), dtype=torch.int32, device=dev)
            t = torch.empty((B, W, 3), dtype=torch.float32, device=dev)
            cat = torch.empty((B, W), dtype=torch.uint8, device=dev) if want_cat else None
        anony = torch.empty((B, W, 3), dtype=torch.int32, device=dev) if want_anony else None
        i2 = self._dev(inject2, torch.int64).to(torch.int32) if inject2 is not None else None
        i3 = self._dev(inject3, torch.int64).to(torch.int32) if inject3 is not None else None
        check(lib().tm_sample_walks(self._h, B, n, N2, ptr(root), ptr(h1n), ptr(h1e), ptr(h1t), s, row_offset,
                                    ptr(i2), ptr(i3), ptr(nodes), ptr(eidx), ptr(t), ptr(anony), ptr(cat),
                                    ptr(hist_null), ptr(hist_prep), ptr(scanned), self._stream()), "tm_sample_walks")
        return nodes, eidx, t, anony, cat

    def find_k_walks(self, degree, src_idx_l, num_neighbors, subgraph_src, seed=None, row_offset=0, inject2=None, inject3=None):
        nodes, eidx, t, anony, _ = self.find_k_walks_device(degree, src_idx_l, num_neighbors, subgraph_src, seed, row_offset,
                                                            inject2, inject3, want_cat=False)
        node_dtype = np.result_type(np.asarray(src_idx_l).dtype, np.int32)   # np.stack of the int64 roots with int32 hops, graph.py:303
        return (nodes.cpu().numpy().astype(node_dtype), eidx.cpu().numpy(), t.cpu().numpy(), anony.cpu().numpy())


    # ------------------------------------------------------------------ the pieces of find_k_walks, reference signatures
    def find_before_walk(self, src_idx_list, cut_time, e_idx=None, return_binary_prob=False):
        """graph.py:149-194: concatenated prefixes of the given nodes; with e_idx a missing key counts as 0 (:174-176)."""
        nodes = [int(v) for v in src_idx_list]
        start, cut = self.find_before_batch_device(nodes, [float(cut_time)] * len(nodes) if e_idx is None else None,
                                                   None if e_idx is None else [int(e_idx)] * len(nodes))
        st, ct = start.cpu().numpy(), cut.cpu().numpy()
        self._err.zero_()                                   # "not found" is not an error here
        off, nbr, e, ts = self._export()
        sl = [slice(int(a), int(a) + int(c)) for a, c in zip(st, ct)]
        source = np.concatenate([np.full(int(c), v, dtype=np.int64) for v, c in zip(nodes, ct)]) if nodes else np.zeros(0, np.int64)
        prob = np.concatenate([self.binary_prob_l[x] for x in sl]) if return_binary_prob else None
        return (source, np.concatenate([nbr[x] for x in sl]), np.concatenate([e[x] for x in sl]), np.concatenate([ts[x] for x in sl]), prob)

    def get_next_step(self, src_idx_l, cut_time_l, num_neighbor, degree, e_idx_l=None, source_id=None, seed=None, row_offset=0):
        """graph.py:308-333 -> (src2, tgt2, e2, t2), each [B*degree, num_neighbor]."""
        assert len(src_idx_l) == len(cut_time_l) == len(source_id) * degree
        if e_idx_l is None:        # prefixes cut by time (find_before_walk's bisect branch, :170-171)
            s = self._next_seed(seed)
            R = len(src_idx_l)
            source = self._dev(np.repeat(np.asarray(source_id), degree), torch.int32)
            nbr = self._dev(src_idx_l, torch.int32); ct = self._dev(np.asarray(cut_time_l, dtype=np.float64), torch.float64)
            o = [torch.empty((R, num_neighbor), dtype=torch.int32, device=self.device) for _ in range(3)]
            o_t = torch.empty((R, num_neighbor), dtype=torch.float32, device=self.device)
            self._clear_err()
            check(lib().tm_walk_next_step_time(self._h, R, int(num_neighbor), ptr(source), ptr(nbr), ptr(ct), s, row_offset * degree, None,
                                               ptr(o[0]), ptr(o[1]), ptr(o[2]), ptr(o_t), ptr(self._err), self._stream()), "tm_walk_next_step_time")
            out = (o[0].cpu().numpy(), o[1].cpu().numpy(), o[2].cpu().numpy(), o_t.cpu().numpy())
            self._raise_if_err("get_next_step")
            return out
        B = len(source_id)
        sub = ([np.asarray(src_idx_l).reshape(B, degree)], [np.asarray(e_idx_l).reshape(B, degree)],
               [np.asarray(cut_time_l, dtype=np.float32).reshape(B, degree)])
        nodes, eidx, t, _, _ = self.find_k_walks_device(degree, source_id, num_neighbor, sub, seed, row_offset, want_anony=False, want_cat=False)
        R = B * degree
        return (nodes[..., 2].reshape(R, -1).cpu().numpy(), nodes[..., 3].reshape(R, -1).cpu().numpy(),
                eidx[..., 1].reshape(R, -1).cpu().numpy(), t[..., 1].reshape(R, -1).cpu().numpy())

    def get_final_step(self, n_id_src_1, n_id_tgt_1, n_id_src_2, n_id_tgt_2, e_id_1, e_id_2, t_id_1, t_id_2, seed=None, row_offset=0, inject=None):
        """graph.py:335-476 -> (src3, tgt3, e3, t3 [R], anony [R, 3]) for walks whose first two events are given."""
        s = self._next_seed(seed)
        dev = self.device
        s1 = self._dev(np.asarray(n_id_src_1).reshape(-1), torch.int32); R = s1.numel()
        t1n = self._dev(np.asarray(n_id_tgt_1).reshape(-1), torch.int32); e1 = self._dev(np.asarray(e_id_1).reshape(-1), torch.int32)
        t1 = self._dev(np.asarray(t_id_1).reshape(-1), torch.float32); t2 = self._dev(np.asarray(t_id_2).reshape(-1), torch.float32)
        step2 = torch.stack([self._dev(np.asarray(a).reshape(-1), torch.int32) for a in (n_id_src_2, n_id_tgt_2, e_id_2)], dim=1).contiguous()
        nodes = torch.empty((R, 6), dtype=torch.int32, device=dev); eidx = torch.empty((R, 3), dtype=torch.int32, device=dev)
        t = torch.empty((R, 3), dtype=torch.float32, device=dev); anony = torch.empty((R, 3), dtype=torch.int32, device=dev)
        inj = self._dev(inject, torch.int64).to(torch.int32) if inject is not None else None
        check(lib().tm_walk_final_step(self._h, R, ptr(s1), ptr(t1n), ptr(e1), ptr(t1), ptr(step2), ptr(t2), s, row_offset, ptr(inj),
                                       ptr(nodes), ptr(eidx), ptr(t), ptr(anony), self._stream()), "tm_walk_final_step")
        return (nodes[:, 0].cpu().numpy(), nodes[:, 1].cpu().numpy(), eidx[:, 0].cpu().numpy(), t[:, 0].cpu().numpy(), anony.cpu().numpy())


# ---------------------------------------------------------------------- class ids / histograms / edge identity
def class_hist_device(anony, want_cat=True):
    """statistic (utils/null_model.py:75-82) + marginal ids (processed/data_preprocess.py:171-208) on device."""
    a = anony.contiguous().view(-1, 3)
    dev = a.device
    hn = torch.zeros(12, dtype=torch.int64, device=dev); hp = torch.zeros(12, dtype=torch.int64, device=dev)
    cat = torch.empty(a.shape[0], dtype=torch.uint8, device=dev) if want_cat else None
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        check(lib().tm_class_hist(a.shape[0], ptr(a), ptr(hn), ptr(hp), ptr(cat), ptr(err), st), "tm_class_hist")
    return hn, hp, (cat.view(anony.shape[:-1]) if want_cat else None), err


def edge_identity_device(eidx, out=None, u8=False):
    """new_edge_info (processed/data_preprocess.py:327-343): [B, W, 3] int32 -> [B, W, 3, 3] float32 (the reference's values), or with
    ``u8`` the same counts as bytes, [B, W, 3, 4] = three counts and a pad byte (W <= 255): the compact form MotifPipeline hands to the scorer."""
    e = eidx.contiguous()
    B, W, _ = e.shape
    dt = torch.uint8 if u8 else torch.float32
    shape = (B, W, 3, 4) if u8 else (B, W, 3, 3)        # bytes: three counts and a pad byte per walk event
    if out is None:
        out = torch.empty(shape, dtype=dt, device=e.device)
    elif out.dtype != dt or tuple(out.shape) != shape:
        raise ValueError("edge_identity_device: out must be %s %s" % (dt, shape))
    st = C.c_void_p(torch.cuda.current_stream(e.device).cuda_stream)
    with torch.cuda.device(e.device):
        if u8:
            check(lib().tm_edge_identity_u8(B, W, ptr(e), ptr(out), st), "tm_edge_identity_u8")
        else:
            check(lib().tm_edge_identity(B, W, ptr(e), ptr(out), st), "tm_edge_identity")
    return out


def new_edge_info(edge_ids):
    """Host-array front end with the reference's signature/dtype (float64 [bsz, n_walks, 3, 3])."""
    e = torch.as_tensor(np.ascontiguousarray(edge_ids)).to(torch.int32).cuda()
    return edge_identity_device(e).cpu().numpy().astype(np.float64)

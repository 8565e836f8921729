"""Device-resident motif pipeline: the whole hot path for a batch of query events without leaving
the GPU -- first-hop lookup/sampling, 3-event walks + anonymisation class + histogram, edge-identity
counts and the fused TempME scorer.  This is what bench.py times; the pieces are the same C-ABI calls
the reference-facing classes (NeighborFinder, TempME) make.

A call is processed in *chunks* of ``chunk_events`` query events: each chunk runs the five kernels back to back
on preallocated per-chunk workspaces (no allocator traffic on the call), so the walk tensors of a chunk
(~85 B per motif) are produced and consumed while they are still in the 126 MB L2, and a call of any size needs a
bounded amount of HBM.  The host API (``submit_host`` / ``collect``) replays the chunk train of a query-set shape as
one CUDA graph.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import TM_EIDX_NONE
from .graph import edge_identity_device


class MotifPipeline:
    """roots = (src, tgt, bgd) of every query event, as processed/data_preprocess.py:106-134 does per event;
    ``group`` = events per reference batch (temp_exp_main.py --bs, default 100): the explainer is called
    once per root type per batch, so the attention's batch-global std runs over [group, W, 2].

    ``hist_null`` / ``hist_prep`` / ``scanned`` are THIS process's running totals over every call since construction
    (or the last ``reset_counters``).  For the job-wide class histogram of a sharded run all-reduce a *copy* once
    (``global_hist``); never all-reduce the accumulators in place."""

    def __init__(self, finder, explainer, n, N2, group=100, seed=0, chunk_events=None, use_graph=True):
        self.finder, self.explainer = finder, explainer
        self.n, self.N2, self.W, self.group = int(n), int(N2), int(n) * int(N2), int(group)
        self.seed = int(seed)
        self.device = finder.device
        self.chunk_events = None if not chunk_events else max(self.group, int(chunk_events) // self.group * self.group)
        self.use_graph = bool(use_graph)
        self.hist_null = torch.zeros(12, dtype=torch.int64, device=self.device)
        self.hist_prep = torch.zeros(12, dtype=torch.int64, device=self.device)
        self.scanned = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._ws = {}
        self._graphs = {}
        self.graph_error = None

    # ------------------------------------------------------------------ bookkeeping
    def reset_counters(self):
        self.hist_null.zero_(); self.hist_prep.zero_(); self.scanned.zero_()

    def global_hist(self, which="null"):
        """Class histogram summed over all ranks: all-reduce of a copy of this rank's running total (the path's only reduction)."""
        h = (self.hist_null if which == "null" else self.hist_prep).clone()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(h)
        return h

    def check_errors(self):
        """IndexError of the reference (utils/graph.py:134-135) for a root whose e_idx is not in its node's list (or a node id out of
        range) in any device call since the last check.  Synchronises the stream; run_host / collect call it."""
        self.finder.check_errors("MotifPipeline")

    def _layout(self, Q):
        g = self.group if Q >= self.group else Q
        if g == 0 or Q % g:
            raise ValueError(f"{Q} query events are not a whole number of reference batches of {self.group}")
        return Q // g, g

    def stage_queries(self, src, dst, fake, ts, eidx):
        """Host arrays -> device tensors of 3Q rows laid out batch-major, [n_batches, 3 (src|tgt|bgd), group]:
        the global row index of a root therefore does not depend on how whole batches are split over GPUs.
        Returns (roots i32, e_idx i32, cut f64, cut f32): the float64 cut time drives the time-cut lookup, its float32
        rounding is what the explainer sees (models/explainer.py:816)."""
        dev = self.device
        Q = len(src)
        nb, g = self._layout(Q)
        def lay(a, b, c, dt):
            x = np.stack([np.asarray(a).reshape(nb, g), np.asarray(b).reshape(nb, g), np.asarray(c).reshape(nb, g)], axis=1)
            return torch.as_tensor(np.ascontiguousarray(x.astype(dt))).to(dev, non_blocking=True).view(-1)
        roots = lay(src, dst, fake, np.int32)
        e = lay(eidx, eidx, np.full(Q, TM_EIDX_NONE, np.int64), np.int32)           # bgd roots are cut by time
        cut64 = lay(ts, ts, ts, np.float64)
        return roots, e, cut64, cut64.to(torch.float32)

    def _workspace(self, R):
        """Per-chunk walk tensors for R roots (allocated once per chunk size, reused by every chunk: launches on one stream)."""
        ws = self._ws.get(R)
        if ws is None:
            dev, n, W = self.device, self.n, self.W
            i32, f32 = torch.int32, torch.float32
            ws = dict(h1=(torch.empty((R, n), dtype=i32, device=dev), torch.empty((R, n), dtype=i32, device=dev), torch.empty((R, n), dtype=f32, device=dev)),
                      walks=(torch.empty((R, W, 6), dtype=i32, device=dev), torch.empty((R, W, 3), dtype=i32, device=dev),
                             torch.empty((R, W, 3), dtype=f32, device=dev), torch.empty((R, W), dtype=torch.uint8, device=dev)),
                      # edge-identity counts in their compact form (bytes, W <= 255) when the walks stay inside the pipeline
                      eid=torch.empty((R, W, 3, 4), dtype=torch.uint8, device=dev) if W <= 255 else torch.empty((R, W, 3, 3), dtype=f32, device=dev))
            self._ws[R] = ws
        return ws

    def run_device(self, roots, e, cut64, cut32=None, row_offset=0, want_walks=False, timers=None, out=None, peer_ptrs=None):
        """roots/e/cut64(/cut32): [3Q] device tensors in the stage_queries layout; row_offset = global index of the first
        root row (3 * events before this shard).  Returns scores [3Q, W] in the same row order (written into ``out`` if given;
        ``peer_ptrs``: the same [3Q, W] segment on peer GPUs, see TempME.score_device).  ``timers``: dict name -> list of
        (start, end) CUDA event pairs, one per chunk and stage."""
        f, n, N2, W = self.finder, self.n, self.N2, self.W
        R = roots.numel()
        g = self.group if R >= 3 * self.group else R // 3
        if cut32 is None:
            cut32 = cut64.to(torch.float32)
        if out is None:
            out = torch.empty((R, W), dtype=torch.float32, device=self.device)
        out2 = out.view(R, W)
        rows_chunk = R if not self.chunk_events else min(R, 3 * self.chunk_events)
        full = None
        if want_walks:
            dev, i32, f32 = self.device, torch.int32, torch.float32
            full = (torch.empty((R, W, 6), dtype=i32, device=dev), torch.empty((R, W, 3), dtype=i32, device=dev), torch.empty((R, W, 3), dtype=f32, device=dev),
                    torch.empty((R, W), dtype=torch.uint8, device=dev), torch.empty((R, W, 3, 3), dtype=f32, device=dev))
        prev = None
        def mark(name):
            nonlocal prev
            if timers is None:
                return
            ev = torch.cuda.Event(enable_timing=True); ev.record()
            if name is not None:
                timers.setdefault(name, []).append((prev, ev))
            prev = ev
        for c0 in range(0, R, rows_chunk):
            c1 = min(R, c0 + rows_chunk)
            rows = c1 - c0
            ws = self._workspace(rows)
            sl = slice(c0, c1)
            mark(None)
            h1 = f.sample_hop_device(roots[sl], cut64[sl], n, e[sl], seed=self.seed, stage=0, row_offset=row_offset + c0, out=ws["h1"])
            mark("sample_hop")
            wout = ws["walks"] if full is None else tuple(x[sl] for x in full[:4])
            nodes, eidx, t, _, cat = f.find_k_walks_device(n, roots[sl], N2, ([h1[0]], [h1[1]], [h1[2]]), seed=self.seed + 1,
                                                           row_offset=row_offset + c0, want_anony=False, want_cat=True,
                                                           hist_null=self.hist_null, hist_prep=self.hist_prep, scanned=self.scanned, out=wout)
            mark("sample_walks")
            eid = edge_identity_device(eidx, out=ws["eid"], u8=ws["eid"].dtype == torch.uint8) if full is None else edge_identity_device(eidx, out=full[4][sl])
            mark("edge_identity")
            pp = [int(p) + c0 * W * 4 for p in peer_ptrs] if peer_ptrs else None
            self.explainer.score_device(nodes, eidx, t, cat, cut32[sl], eid, group=max(g, 1), out=out2[sl], peer_ptrs=pp, fanout=N2)
            mark("encode")
        return (out2, full) if want_walks else out2

    def explain_device(self, roots, e, cut64, row_offset=0):
        """The explanation pass of temp_exp_main.py's evaluation for a staged batch, without leaving the GPU: 2-hop subgraph
        (find_k_hop, graph.py:233-262), walks on its first hop, scores, and TempME.retrieve_edge_imp_node (eval) on them.
        Returns (scores [3Q, W], edge_imp_0 [3Q, n], edge_imp_1 [3Q, n^2], subgraph records) in the staged row order."""
        f, n, N2 = self.finder, self.n, self.N2
        R = roots.numel()
        g = self.group if R >= 3 * self.group else R // 3
        # find_k_hop(2): hop 0 as in run_device (e_idx window for src / tgt roots, time cut for the bgd roots), hop 1 by e_idx (graph.py:247-250)
        h0 = f.sample_hop_device(roots, cut64, n, e, seed=self.seed, stage=0, row_offset=row_offset)
        h1 = f.sample_hop_device(h0[0].reshape(-1), None, n, h0[1].reshape(-1), seed=self.seed, stage=1, row_offset=row_offset * n)
        sub = ([h0[0], h1[0].view(R, n * n)], [h0[1], h1[1].view(R, n * n)], [h0[2], h1[2].view(R, n * n)])
        nodes, eidx, t, _, cat = f.find_k_walks_device(n, roots, N2, ([sub[0][0]], [sub[1][0]], [sub[2][0]]), seed=self.seed + 1,
                                                       row_offset=row_offset, want_anony=False, want_cat=True,
                                                       hist_null=self.hist_null, hist_prep=self.hist_prep, scanned=self.scanned)
        eid = edge_identity_device(eidx)
        scores = self.explainer.score_device(nodes, eidx, t, cat, cut64.to(torch.float32), eid, group=max(g, 1), fanout=N2)
        imp0, imp1 = self.explainer.edge_importance_device(scores, eidx, t, sub[0][0].view(R, n), sub[1][0].view(R, n),
                                                           sub[0][1].view(R, n * n), sub[1][1].view(R, n * n))
        return scores, imp0, imp1, sub

    def unstage_scores(self, scores, Q):
        """[3Q, W] in batch-major row order -> [3, Q, W] (src | tgt | bgd)."""
        nb, g = self._layout(Q)
        return scores.view(nb, 3, g, self.W).permute(1, 0, 2, 3).reshape(3, Q, self.W)

    # ------------------------------------------------------------------ host API (pinned staging, two batches in flight)
    def _slots(self, Q):
        if getattr(self, "_sl_q", None) != Q:
            dev = self.device
            R = 3 * Q
            # one pinned / device block per slot: roots i32 | e_idx i32 | cut f32 (int32 words), then cut f64
            self._sl = [dict(pin_i=torch.empty((3, R), dtype=torch.int32).pin_memory(), pin_t=torch.empty(R, dtype=torch.float64).pin_memory(),
                             pin_out=torch.empty((3, Q, self.W), dtype=torch.float32).pin_memory(),
                             dev_i=torch.empty((3, R), dtype=torch.int32, device=dev), dev_t=torch.empty(R, dtype=torch.float64, device=dev),
                             dev_scores=torch.empty((R, self.W), dtype=torch.float32, device=dev),
                             dev_out=torch.empty((3, Q, self.W), dtype=torch.float32, device=dev),
                             ready=torch.cuda.Event(), copied=torch.cuda.Event()) for _ in range(2)]
            self._sl_next = 0
            self._sl_q = Q
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._graphs = {}
        return self._sl

    def _fill_slot(self, sl, src, dst, fake, ts, eidx):
        Q = len(src)
        nb, g = self._layout(Q)
        pi = sl["pin_i"].numpy()
        r = pi[0].reshape(nb, 3, g); e = pi[1].reshape(nb, 3, g); c32 = pi[2].view(np.float32).reshape(nb, 3, g)
        c = sl["pin_t"].numpy().reshape(nb, 3, g)
        r[:, 0] = np.asarray(src).reshape(nb, g); r[:, 1] = np.asarray(dst).reshape(nb, g); r[:, 2] = np.asarray(fake).reshape(nb, g)
        e[:, 0] = e[:, 1] = np.asarray(eidx).reshape(nb, g); e[:, 2] = TM_EIDX_NONE                  # bgd roots are cut by time
        t64 = np.asarray(ts, np.float64).reshape(nb, 1, g)
        c[:] = t64
        c32[:] = t64                                                                                   # .float(), explainer.py:816

    def _device_train(self, sl, Q, row_offset):
        """The device work of one host call on slot-owned buffers: chunk train, then the [3, Q, W] reordering."""
        di = sl["dev_i"]
        scores = self.run_device(di[0], di[1], sl["dev_t"], di[2].view(torch.float32), row_offset, out=sl["dev_scores"])
        sl["dev_out"].copy_(self.unstage_scores(scores, Q))

    def _launch(self, ticket, Q, row_offset):
        """Replays the slot's CUDA graph of the chunk train (captured on first use per (slot, row_offset)); direct launches when graphs are off."""
        sl = self._sl[ticket]
        if not self.use_graph or self.graph_error is not None:
            self._device_train(sl, Q, row_offset)
            return
        key = (ticket, int(row_offset))
        g = self._graphs.get(key)
        if g is None:
            self._device_train(sl, Q, row_offset)                 # warm-up outside capture: workspaces, packed weights, function attributes
            torch.cuda.current_stream(self.device).synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._device_train(sl, Q, row_offset)
                self._graphs[key] = g
            except Exception as ex:                               # noqa: BLE001 -- capture is an optimisation: fall back to direct launches, keep the reason
                self.graph_error = f"{type(ex).__name__}: {ex}"
                torch.cuda.synchronize(self.device)
            return                                                # the warm-up run produced this call's results; capture itself launches nothing
        g.replay()

    def submit_host(self, src, dst, fake, ts, eidx, row_offset=0):
        """Asynchronous ``run_host``: stages the queries into one of two pinned slots, enqueues H2D + the device pipeline on the
        current stream and the D2H of the scores on a copy stream, and returns a ticket for ``collect``.  At most two batches are in
        flight; collect batch k-1 before submitting batch k+1."""
        Q = len(src)
        sl = self._slots(Q)[self._sl_next]
        ticket = self._sl_next
        self._sl_next ^= 1
        self._fill_slot(sl, src, dst, fake, ts, eidx)
        with torch.cuda.device(self.device):
            sl["dev_i"].copy_(sl["pin_i"], non_blocking=True)
            sl["dev_t"].copy_(sl["pin_t"], non_blocking=True)
            self._launch(ticket, Q, row_offset)
            sl["ready"].record()
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(sl["ready"])
                sl["pin_out"].copy_(sl["dev_out"], non_blocking=True)
                sl["copied"].record()
        return ticket

    def collect(self, ticket, check=False):
        """Scores [3, Q, W] of a submitted batch (a view of the slot's pinned buffer, valid until the slot is submitted again).
        check=True also raises the reference's IndexError for bad roots (one more 4-byte D2H)."""
        sl = self._sl[ticket]
        sl["copied"].synchronize()
        if check:
            self.check_errors()
        return sl["pin_out"].numpy()

    def run_host(self, src, dst, fake, ts, eidx, row_offset=0):
        """End-to-end call with host buffers: H2D of the queries, the device pipeline, D2H of the scores.
        Returns a [3, Q, W] float32 array backed by an internal pinned buffer (valid until the slot is reused, two calls later)."""
        return self.collect(self.submit_host(src, dst, fake, ts, eidx, row_offset), check=True)

"""Device-resident motif pipeline: the whole hot path for a batch of query events without leaving
the GPU -- first-hop lookup/sampling, 3-event walks + anonymisation class + histogram, edge-identity
counts and the fused TempME scorer.  This is what bench.py times; the pieces are the same C-ABI calls
the reference-facing classes (NeighborFinder, TempME) make.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import TM_EIDX_NONE
from .graph import edge_identity_device


class MotifPipeline:
    """roots = (src, tgt, bgd) of every query event, as processed/data_preprocess.py:106-134 does per event;
    ``group`` = events per reference batch (temp_exp_main.py --bs, default 100): the explainer is called
    once per root type per batch, so the attention's batch-global std runs over [group, W, 2]."""

    def __init__(self, finder, explainer, n, N2, group=100, seed=0):
        self.finder, self.explainer = finder, explainer
        self.n, self.N2, self.W, self.group = int(n), int(N2), int(n) * int(N2), int(group)
        self.seed = int(seed)
        self.device = finder.device
        self.hist_null = torch.zeros(12, dtype=torch.int64, device=self.device)
        self.hist_prep = torch.zeros(12, dtype=torch.int64, device=self.device)
        self.scanned = torch.zeros(1, dtype=torch.int64, device=self.device)

    def _layout(self, Q):
        g = self.group if Q >= self.group else Q
        if g == 0 or Q % g:
            raise ValueError(f"{Q} query events are not a whole number of reference batches of {self.group}")
        return Q // g, g

    def stage_queries(self, src, dst, fake, ts, eidx):
        """Host arrays -> device tensors of 3Q rows laid out batch-major, [n_batches, 3 (src|tgt|bgd), group]:
        the global row index of a root therefore does not depend on how whole batches are split over GPUs."""
        dev = self.device
        Q = len(src)
        nb, g = self._layout(Q)
        def lay(a, b, c, dt):
            x = np.stack([np.asarray(a).reshape(nb, g), np.asarray(b).reshape(nb, g), np.asarray(c).reshape(nb, g)], axis=1)
            return torch.as_tensor(np.ascontiguousarray(x.astype(dt))).to(dev, non_blocking=True).view(-1)
        roots = lay(src, dst, fake, np.int32)
        e = lay(eidx, eidx, np.full(Q, TM_EIDX_NONE, np.int64), np.int32)           # bgd roots are cut by time
        cut64 = lay(ts, ts, ts, np.float64)
        return roots, e, cut64

    def run_device(self, roots, e, cut64, row_offset=0, want_walks=False, timers=None, out=None, peer_ptrs=None):
        """roots/e/cut64: [3Q] device tensors in the stage_queries layout; row_offset = global index of the first
        root row (3 * events before this shard).  Returns scores [3Q, W] in the same row order."""
        f, n, N2, W = self.finder, self.n, self.N2, self.W
        R = roots.numel()
        g = self.group if R >= 3 * self.group else R // 3
        def mark(name):
            if timers is not None:
                ev = torch.cuda.Event(enable_timing=True); ev.record(); timers.append((name, ev))
        mark("start")
        h1 = f.sample_hop_device(roots, cut64, n, e, seed=self.seed, stage=0, row_offset=row_offset)
        mark("sample_hop")
        nodes, eidx, t, _, cat = f.find_k_walks_device(n, roots, N2, ([h1[0]], [h1[1]], [h1[2]]), seed=self.seed + 1,
                                                       row_offset=row_offset, want_anony=False, want_cat=True,
                                                       hist_null=self.hist_null, hist_prep=self.hist_prep, scanned=self.scanned)
        mark("sample_walks")
        eid = edge_identity_device(eidx)
        mark("edge_identity")
        scores = self.explainer.score_device(nodes, eidx, t, cat, cut64.to(torch.float32), eid, group=max(g, 1), out=out, peer_ptrs=peer_ptrs)
        mark("encode")
        return (scores, (nodes, eidx, t, cat, eid)) if want_walks else scores

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
        scores = self.explainer.score_device(nodes, eidx, t, cat, cut64.to(torch.float32), eid, group=max(g, 1))
        imp0, imp1 = self.explainer.edge_importance_device(scores, eidx, t, sub[0][0].view(R, n), sub[1][0].view(R, n),
                                                           sub[0][1].view(R, n * n), sub[1][1].view(R, n * n))
        return scores, imp0, imp1, sub

    def unstage_scores(self, scores, Q):
        """[3Q, W] in batch-major row order -> [3, Q, W] (src | tgt | bgd)."""
        nb, g = self._layout(Q)
        return scores.view(nb, 3, g, self.W).permute(1, 0, 2, 3).reshape(3, Q, self.W)

    def _host_buffers(self, Q):
        """Pinned host staging (inputs [3Q] in the batch-major layout, scores [3, Q, W]) + matching device tensors,
        allocated once per query count."""
        if getattr(self, "_hb_q", None) != Q:
            dev = self.device
            self._pin_i32 = torch.empty((2, 3 * Q), dtype=torch.int32).pin_memory()      # roots | e_idx
            self._pin_f64 = torch.empty(3 * Q, dtype=torch.float64).pin_memory()         # cut times
            self._pin_out = torch.empty((3, Q, self.W), dtype=torch.float32).pin_memory()
            self._dev_i32 = torch.empty((2, 3 * Q), dtype=torch.int32, device=dev)
            self._dev_f64 = torch.empty(3 * Q, dtype=torch.float64, device=dev)
            self._hb_q = Q
        return self._pin_i32, self._pin_f64, self._pin_out

    # ------------------------------------------------------------------ asynchronous host API (two batches in flight)
    def _slots(self, Q):
        if getattr(self, "_sl_q", None) != Q:
            dev = self.device
            self._sl = [dict(pin_i=torch.empty((2, 3 * Q), dtype=torch.int32).pin_memory(), pin_t=torch.empty(3 * Q, dtype=torch.float64).pin_memory(),
                             pin_out=torch.empty((3, Q, self.W), dtype=torch.float32).pin_memory(),
                             dev_i=torch.empty((2, 3 * Q), dtype=torch.int32, device=dev), dev_t=torch.empty(3 * Q, dtype=torch.float64, device=dev),
                             dev_out=torch.empty((3, Q, self.W), dtype=torch.float32, device=dev),
                             ready=torch.cuda.Event(), copied=torch.cuda.Event()) for _ in range(2)]
            self._sl_next = 0
            self._sl_q = Q
            self._copy_stream = torch.cuda.Stream(device=dev)
        return self._sl

    def submit_host(self, src, dst, fake, ts, eidx, row_offset=0):
        """Asynchronous ``run_host``: stages the queries into one of two pinned slots, enqueues H2D + the device pipeline on the
        current stream and the D2H of the scores on a copy stream, and returns a ticket for ``collect``.  At most two batches are in
        flight; collect batch k-1 before submitting batch k+1."""
        Q = len(src)
        nb, g = self._layout(Q)
        sl = self._slots(Q)[self._sl_next]
        ticket = self._sl_next
        self._sl_next ^= 1
        r = sl["pin_i"][0].numpy().reshape(nb, 3, g); e = sl["pin_i"][1].numpy().reshape(nb, 3, g); c = sl["pin_t"].numpy().reshape(nb, 3, g)
        r[:, 0] = np.asarray(src).reshape(nb, g); r[:, 1] = np.asarray(dst).reshape(nb, g); r[:, 2] = np.asarray(fake).reshape(nb, g)
        e[:, 0] = e[:, 1] = np.asarray(eidx).reshape(nb, g); e[:, 2] = TM_EIDX_NONE                  # bgd roots are cut by time
        c[:] = np.asarray(ts, np.float64).reshape(nb, 1, g)
        sl["dev_i"].copy_(sl["pin_i"], non_blocking=True)
        sl["dev_t"].copy_(sl["pin_t"], non_blocking=True)
        sl["dev_out"].copy_(self.unstage_scores(self.run_device(sl["dev_i"][0], sl["dev_i"][1], sl["dev_t"], row_offset), Q))   # slot-owned: no allocator traffic
        sl["ready"].record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(sl["ready"])
            sl["pin_out"].copy_(sl["dev_out"], non_blocking=True)
            sl["copied"].record()
        return ticket

    def collect(self, ticket):
        """Scores [3, Q, W] of a submitted batch (a view of the slot's pinned buffer, valid until the slot is submitted again)."""
        sl = self._sl[ticket]
        sl["copied"].synchronize()
        return sl["pin_out"].numpy()

    def run_host(self, src, dst, fake, ts, eidx, row_offset=0):
        """End-to-end call with host buffers: H2D of the queries, the device pipeline, D2H of the scores.
        Returns a [3, Q, W] float32 array backed by an internal pinned buffer (valid until the next run_host call)."""
        Q = len(src)
        nb, g = self._layout(Q)
        pin_i, pin_t, pin_out = self._host_buffers(Q)
        r = pin_i[0].numpy().reshape(nb, 3, g); e = pin_i[1].numpy().reshape(nb, 3, g); c = pin_t.numpy().reshape(nb, 3, g)
        r[:, 0] = np.asarray(src).reshape(nb, g); r[:, 1] = np.asarray(dst).reshape(nb, g); r[:, 2] = np.asarray(fake).reshape(nb, g)
        e[:, 0] = e[:, 1] = np.asarray(eidx).reshape(nb, g); e[:, 2] = TM_EIDX_NONE                  # bgd roots are cut by time
        c[:] = np.asarray(ts, np.float64).reshape(nb, 1, g)
        self._dev_i32.copy_(pin_i, non_blocking=True)
        self._dev_f64.copy_(pin_t, non_blocking=True)
        scores = self.run_device(self._dev_i32[0], self._dev_i32[1], self._dev_f64, row_offset)
        pin_out.copy_(self.unstage_scores(scores, Q), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return pin_out.numpy()

"""``TempME`` with the reference's constructor, parameter names and ``forward`` contract
(reference models/explainer.py:99-201), scoring motifs with the fused sm_100a kernel of
libtempme_b200 (``tm_encode_score``).

The sub-module / parameter names equal the reference's so a reference
``state_dict`` loads with ``load_state_dict`` and ours loads into the reference class.  Forward values
come from the fused kernel whenever no dropout is active; when gradients are requested the methods return
tensors with an autograd graph (``tempme_b200.training``: recompute backward of the scorer, Beta ``rsample``
and ``kl_loss`` kernels with their own backward), so ``temp_exp_main.py``'s training loop runs on this class.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import training as _tr
from ._lib import EncoderDesc, EncoderParams, GateDesc, GateParams, check, lib, ptr


class TimeEncode(nn.Module):
    """Parameters of the reference TimeEncode (explainer.py:45-50): basis_freq = 1/10^linspace(0,9,D), phase = 0."""

    def __init__(self, expand_dim):
        super().__init__()
        self.time_dim = expand_dim
        self.basis_freq = nn.Parameter(torch.from_numpy(1 / 10 ** np.linspace(0, 9, expand_dim)).float())
        self.phase = nn.Parameter(torch.zeros(expand_dim).float())


class _EventGCN(nn.Module):  # parameter container for event_gcn (explainer.py:79-84)
    def __init__(self, event_dim, node_dim, hid_dim):
        super().__init__()
        self.lin_event = nn.Linear(event_dim, node_dim)
        self.relu = nn.ReLU()
        self.MLP = nn.Sequential(nn.Linear(node_dim, hid_dim), nn.ReLU(), nn.Linear(hid_dim, hid_dim))


class _Attention(nn.Module):  # Attention (explainer.py:12-23)
    def __init__(self, input_dim, hid_dim):
        super().__init__()
        self.hidden_size = hid_dim
        self.W1 = nn.Linear(input_dim, input_dim)
        self.W2 = nn.Linear(input_dim, input_dim)
        self.MLP = nn.Sequential(nn.Linear(input_dim, hid_dim), nn.ReLU(), nn.Linear(hid_dim, hid_dim))
        nn.init.xavier_uniform_(self.W2.weight.data)
        self.W2.bias.data.fill_(0.1)


class _TemporalAwareAttention(nn.Module):  # TemporalAwareAttention (explainer.py:768-787)
    def __init__(self, input_dim, hid_dim, dropout_p=0.1):
        super().__init__()
        self.hidden_size = hid_dim
        self.W1 = nn.Linear(input_dim, input_dim)
        self.W2 = nn.Linear(input_dim, input_dim)
        self.W_time = nn.Linear(1, input_dim)
        self.dropout = nn.Dropout(dropout_p)
        self.MLP = nn.Sequential(nn.Linear(input_dim, hid_dim), nn.ReLU(), nn.Dropout(dropout_p), nn.Linear(hid_dim, hid_dim))
        nn.init.xavier_uniform_(self.W2.weight.data)
        self.W2.bias.data.fill_(0.1)
        nn.init.xavier_uniform_(self.W_time.weight.data)


class _MergeLayer(nn.Module):  # _MergeLayer (explainer.py:62-69)
    def __init__(self, input_dim, hid_dim):
        super().__init__()
        self.fc1 = nn.Linear(2 * input_dim, hid_dim)
        self.fc2 = nn.Linear(hid_dim, 1)
        nn.init.xavier_normal_(self.fc1.weight)
        nn.init.xavier_normal_(self.fc2.weight)
        self.act = nn.ReLU()

    def forward(self, x1, x2):  # explainer.py:71-76; one row per query event
        return self.fc2(self.act(self.fc1(torch.cat([x1, x2], dim=-1))))


class TempME(nn.Module):
    def __init__(self, base, base_model_type, data, out_dim, hid_dim, prior="empirical", temp=0.07,
                 if_cat_feature=True, dropout_p=0.1, device=None, use_temporal_guidance=True,
                 use_dependency_aware_sampling=True, null_model=None, batch_group=None):
        super().__init__()
        self.node_dim = base.n_feat_th.shape[1]
        self.edge_dim = base.e_feat_th.shape[1]
        self.time_dim = self.node_dim
        self.out_dim = out_dim
        self.hid_dim = hid_dim
        self.base_type = base_model_type
        self.dropout_p = dropout_p
        self.temp = temp
        self.prior = prior
        self.if_cat = if_cat_feature
        self.dropout = nn.Dropout(dropout_p)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("tempme_b200.TempME scores on a CUDA device (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.event_dim = self.edge_dim + self.time_dim + 3
        self.event_conv = _EventGCN(self.event_dim, self.node_dim, self.hid_dim)
        self.use_temporal_guidance = use_temporal_guidance
        self.attention = (_TemporalAwareAttention if use_temporal_guidance else _Attention)(2 * self.hid_dim, self.hid_dim)
        self.mlp_dim = self.hid_dim + 12 if self.if_cat else self.hid_dim
        self.MLP = nn.Sequential(nn.Linear(self.mlp_dim, self.mlp_dim), nn.ReLU(), nn.Dropout(self.dropout_p),
                                 nn.Linear(self.mlp_dim, self.hid_dim), nn.ReLU(), nn.Linear(self.hid_dim, 1))
        self.final_linear = nn.Linear(2 * self.hid_dim, self.hid_dim)
        self.node_emd_dim = self.hid_dim + 12 + self.node_dim if self.if_cat else self.hid_dim + self.node_dim
        self.affinity_score = _MergeLayer(self.node_emd_dim, self.node_emd_dim)
        self.edge_raw_embed = base.edge_raw_features
        self.node_raw_embed = base.node_raw_features
        self.time_encoder = TimeEncode(expand_dim=self.time_dim)
        if null_model is None:
            from .null_model import get_null_distribution
            null_model = get_null_distribution(data_name=data)
        self.null_model = null_model
        num_nodes = base.n_feat_th.shape[0]
        self.node_degree = torch.ones(num_nodes, device=self.device)
        self.use_dependency_aware_sampling = use_dependency_aware_sampling
        if use_dependency_aware_sampling:  # parameters of the edge-level modules (explainer.py:141-171), unused by forward
            self.edge_dependency_gcn = nn.Sequential(
                nn.Linear(self.edge_dim + self.time_dim, self.hid_dim), nn.ReLU(), nn.Dropout(dropout_p * 1.5),
                nn.Linear(self.hid_dim, self.hid_dim // 2), nn.ReLU(), nn.Dropout(dropout_p), nn.Linear(self.hid_dim // 2, 1))
            self.edge_importance_attention = nn.MultiheadAttention(embed_dim=self.hid_dim, num_heads=4, dropout=dropout_p, batch_first=True)
            self.edge_to_node_transform = nn.Sequential(nn.Linear(self.edge_dim, self.hid_dim), nn.ReLU(), nn.Linear(self.hid_dim, self.hid_dim))
            self.gumbel_temperature = 1.0
            self.min_gumbel_temperature = 0.5
            self.gumbel_anneal_rate = 0.003
        # ---- device-side state of the fused scorer
        self.batch_group = batch_group          # roots per reference batch; None = the whole call is one batch
        self._desc = EncoderDesc(self.node_dim, self.edge_dim, self.hid_dim, int(bool(use_temporal_guidance)), int(bool(self.if_cat)), 0)
        # edge-projection mode of the scorer: lin_event's edge columns applied once per edge id (a [rows, node_dim] table next to the
        # feature table, rebuilt when the weights change) instead of once per walk event.  TEMPME_EDGE_PROJECTION=0 disables it.
        self._desc_proj = EncoderDesc(self.node_dim, self.edge_dim, self.hid_dim, int(bool(use_temporal_guidance)), int(bool(self.if_cat)), 1)
        self.edge_projection = os.environ.get("TEMPME_EDGE_PROJECTION", "1") != "0" and self.edge_dim <= 256
        self._proj = None
        self._proj_key = None
        self._desc_fanout = {}
        self.projection_ms = None
        self._blob = None
        self._blob_key = None
        self.autograd_in_eval = False           # see _wants_grad
        self._ws = None
        self._ws_retired = []
        self._gate_desc = GateDesc(self.edge_dim, self.time_dim, self.hid_dim)
        self._gate_blob = None
        self._gate_key = None

    # ------------------------------------------------------------------ weights -> packed device blob
    def _forward_params(self):
        a3 = self.attention.MLP[3] if self.use_temporal_guidance else self.attention.MLP[2]
        return [self.event_conv.lin_event.weight, self.event_conv.lin_event.bias,
                self.event_conv.MLP[0].weight, self.event_conv.MLP[0].bias, self.event_conv.MLP[2].weight, self.event_conv.MLP[2].bias,
                self.attention.W1.weight, self.attention.W1.bias, self.attention.W2.weight, self.attention.W2.bias,
                self.attention.MLP[0].weight, self.attention.MLP[0].bias, a3.weight, a3.bias,
                self.MLP[0].weight, self.MLP[0].bias, self.MLP[3].weight, self.MLP[3].bias, self.MLP[5].weight, self.MLP[5].bias,
                self.time_encoder.basis_freq, self.time_encoder.phase]

    def packed_weights(self):
        ps = self._forward_params()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._blob is None or key != self._blob_key:
            host = [p.detach().to("cpu", torch.float32).contiguous() for p in ps]
            prm = EncoderParams(*[C.c_void_p(h.data_ptr()) for h in host])
            n = lib().tm_encoder_blob_floats(C.byref(self._desc))
            blob = torch.empty(n, dtype=torch.float32)
            check(lib().tm_encoder_pack(C.byref(self._desc), C.byref(prm), ptr(blob)), "tm_encoder_pack")
            self._blob = blob.to(self.device)
            self._blob_key = key
        return self._blob

    def _tables(self):
        nf = self.node_raw_embed.weight if hasattr(self.node_raw_embed, "weight") else self.node_raw_embed
        ef = self.edge_raw_embed.weight if hasattr(self.edge_raw_embed, "weight") else self.edge_raw_embed
        if nf.device != self.device or nf.dtype != torch.float32 or not nf.is_contiguous():
            nf = nf.detach().to(self.device, torch.float32).contiguous()
        if ef.device != self.device or ef.dtype != torch.float32 or not ef.is_contiguous():
            ef = ef.detach().to(self.device, torch.float32).contiguous()
        return nf, ef

    def _edge_table(self, blob, ef):
        """(desc, table) the scorer reads edge rows from: the projected table P = edge_feat @ lin_event.weight[:, :Ed]^T (cached per
        weights version and feature table; tm_encoder_project_edges) or, with the projection off, the raw feature table."""
        if not self.edge_projection or os.environ.get("TEMPME_ENCODER") == "ffma":
            return self._desc, ef
        key = (self._blob_key, ef.data_ptr(), tuple(ef.shape))
        if self._proj is None or key != self._proj_key:
            if self._proj is not None:
                self._ws_retired.append(self._proj)          # a captured graph may still gather from it
            P = torch.empty((ef.shape[0], self.node_dim), dtype=torch.float32, device=self.device)
            st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            capturing = torch.cuda.is_current_stream_capturing()
            if not capturing:
                t0.record()
            check(lib().tm_encoder_project_edges(C.byref(self._desc), ptr(blob), ptr(ef), ef.shape[0], ptr(P), self.device.index, st),
                  "tm_encoder_project_edges")
            if not capturing:
                t1.record(); t1.synchronize()
                self.projection_ms = t0.elapsed_time(t1)
            self._proj, self._proj_key = P, key
        return self._desc_proj, self._proj

    def _workspace(self, B, W, group):
        """Scratch of the scorer (per-batch std + the resident CTAs' h slabs).  It only grows, and a replaced buffer stays alive:
        a captured CUDA graph (MotifPipeline) may still hold its address."""
        with torch.cuda.device(self.device):
            nws = lib().tm_encoder_workspace_floats(C.byref(self._desc), B, W, group)
        if self._ws is None or self._ws.numel() < nws:
            if self._ws is not None:
                self._ws_retired.append(self._ws)
            self._ws = torch.empty(max(nws, 1024), dtype=torch.float32, device=self.device)
        return self._ws

    def _t(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a)).to(dtype).to(self.device, non_blocking=True)

    # ------------------------------------------------------------------ forward (explainer.py:174-201)
    def _with_fanout(self, desc, fanout, eid_u8=False):
        """The descriptor with the walk-layout hint (tm_encoder_desc.walk_fanout) and the edge-identity dtype flag set."""
        fanout = int(fanout or 0)
        if fanout < 2 and not eid_u8:
            return desc
        key = (desc.edge_projected, max(fanout, 0), bool(eid_u8))
        if key not in self._desc_fanout:
            self._desc_fanout[key] = EncoderDesc(desc.node_dim, desc.edge_dim, desc.hid_dim, desc.use_temporal, desc.if_cat, desc.edge_projected,
                                                 fanout if fanout >= 2 else 0, int(bool(eid_u8)))
        return self._desc_fanout[key]

    _fanout_seen = {}               # W -> fan-out found last time (tried first)

    @classmethod
    def detect_fanout(cls, eidx, nodes, min_motifs=0):
        """Largest c <= 10 dividing W such that every c consecutive walks have the same event next to the root -- find_k_walks' layout
        w = i1 * N2 + j (utils/graph.py:290-300) gives c = N2.  Only a hint: the scorer verifies it tile by tile.  Calls below
        ``min_motifs`` walks skip the test (each candidate costs a small reduction and a synchronisation)."""
        B, W = eidx.shape[0], eidx.shape[1]
        if B == 0 or B * W < min_motifs:
            return 1
        key = torch.stack([eidx[:, :, 2], nodes[:, :, 4], nodes[:, :, 5]], dim=-1)

        def uniform(c):
            g = key.view(B, W // c, c, 3)
            return bool((g == g[:, :, :1]).all())
        last = cls._fanout_seen.get(W)
        cands = [c for c in range(min(10, W), 1, -1) if W % c == 0]
        if last in cands and uniform(last) and not any(uniform(c) for c in cands if c > last and c % last == 0):
            return last
        for c in cands:
            if uniform(c):
                cls._fanout_seen[W] = c
                return c
        cls._fanout_seen[W] = 1
        return 1

    def score_device(self, nodes, eidx, t, cat, cut_time, edge_identity, group=None, out=None, peer_ptrs=None, fanout=None):
        """All arguments CUDA tensors: nodes i32 [B,W,6], eidx i32 [B,W,3], t f32 [B,W,3], cat u8 [B,W],
        cut_time f32 [B], edge_identity f32 [B,W,3,3] (or the byte counts of edge_identity_device(u8=True), [B,W,3,4]) -> scores f32 [B,W].  out: preallocated [B,W] result (e.g. this rank's segment of a
        gathered buffer); peer_ptrs: device addresses of the same segment on up to 7 peer GPUs -- the kernel stores every score there too
        (tm_encode_score_gather; tempme_b200.dist.ScoreExchange).  fanout: N2 when the walks come from find_k_walks (w = i1 * N2 + j): the
        event next to the root is then evaluated once per N2 walks (verified by the kernel; the scores do not depend on the hint)."""
        B, W = nodes.shape[0], nodes.shape[1]
        group = int(group or self.batch_group or max(B, 1))
        blob = self.packed_weights()
        nf, ef = self._tables()
        desc, ef = self._edge_table(blob, ef)
        if edge_identity.dtype not in (torch.float32, torch.uint8):
            raise ValueError("score_device: edge_identity must be float32 (tm_edge_identity) or uint8 (tm_edge_identity_u8)")
        desc = self._with_fanout(desc, fanout, eid_u8=edge_identity.dtype == torch.uint8)
        self._workspace(B, W, group)
        if out is None:
            scores = torch.empty((B, W), dtype=torch.float32, device=self.device)
        else:
            scores = out
            if scores.dtype != torch.float32 or scores.numel() != B * W or not scores.is_contiguous() or scores.device != self.device:
                raise ValueError("score_device: out must be a contiguous float32 [B, W] tensor on the explainer's device")
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if peer_ptrs:
            arr = (C.c_uint64 * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
            check(lib().tm_encode_score_gather(C.byref(desc), ptr(blob), B, W, group, ptr(nodes), ptr(eidx), ptr(t), ptr(cat),
                                               ptr(cut_time), ptr(edge_identity), ptr(nf), nf.shape[0], ptr(ef), ef.shape[0],
                                               ptr(self._ws), ptr(scores), arr, len(peer_ptrs), self.device.index, st), "tm_encode_score_gather")
            return scores
        check(lib().tm_encode_score(C.byref(desc), ptr(blob), B, W, group, ptr(nodes), ptr(eidx), ptr(t), ptr(cat),
                                    ptr(cut_time), ptr(edge_identity), ptr(nf), nf.shape[0], ptr(ef), ef.shape[0],
                                    ptr(self._ws), ptr(scores), self.device.index, st), "tm_encode_score")
        return scores

    def scorer_parameters(self):
        """The parameters ``forward`` depends on, in a fixed order (the autograd inputs of training.FusedScore)."""
        att = self.attention
        mods = [self.event_conv, att.W1, att.W2, att.MLP, self.MLP, self.time_encoder]
        return [p for mod in mods for p in mod.parameters()]

    def _dropout_active(self):
        """train() mode with at least one live Dropout (the attention module keeps its own default p = 0.1 whatever dropout_p is,
        explainer.py:121): the fused kernel has no dropout, so such calls take the layer-by-layer route."""
        return self.training and any(isinstance(mod, nn.Dropout) and mod.p > 0 for mod in self.modules())

    def _wants_grad(self):
        """Build an autograd graph?  In train() mode whenever gradients are enabled and a scorer parameter requires them.  In eval()
        mode the reference's own evaluation loops run with gradients enabled and never call backward (temp_exp_main.py:300-330), so
        eval() returns plain values from the fused kernels unless ``autograd_in_eval`` is set."""
        return (self.training or self.autograd_in_eval) and torch.is_grad_enabled() and any(p.requires_grad for p in self.scorer_parameters())

    def forward(self, walks, cut_time_l, edge_identify):
        node_idx, edge_idx, time_idx, cat_feat, _ = walks
        nodes = self._t(node_idx, torch.int32)
        eidx = self._t(edge_idx, torch.int32)
        t = self._t(time_idx, torch.float32)                              # .float(), explainer.py:325
        B, W = nodes.shape[0], nodes.shape[1]
        cat = self._t(cat_feat, torch.uint8).view(B, W) if self.if_cat else None
        cut = self._t(cut_time_l, torch.float32)                          # .float(), explainer.py:816
        eid = self._t(edge_identify, torch.float32)                       # .float(), explainer.py:177
        if self._wants_grad():                                            # training loop (temp_exp_main.py:605-632): scores with an autograd graph
            return _tr.score_autograd(self, nodes, eidx, t, cat, cut, eid)
        if self._dropout_active():                                        # train() under no_grad: dropout is part of the value
            return _tr.scores_layerwise(self, nodes.long(), eidx.long(), t, cat, cut, eid)
        return self.score_device(nodes, eidx, t, cat, cut, eid, fanout=self.detect_fanout(eidx, nodes, min_motifs=8192)).view(B, W, 1)

    # ------------------------------------------------------------------ enhance path (explainer.py:203-306), eval mode
    def _walk_tensors(self, walks, cut_time_l, edge_identify):
        node_idx, edge_idx, time_idx, cat_feat, _ = walks
        nodes = self._t(node_idx, torch.int32); eidx = self._t(edge_idx, torch.int32); t = self._t(time_idx, torch.float32)
        B, W = nodes.shape[0], nodes.shape[1]
        cat = self._t(cat_feat, torch.uint8).view(B, W) if self.if_cat else None
        return nodes, eidx, t, cat, self._t(cut_time_l, torch.float32), self._t(edge_identify, torch.float32)

    def compute_walk_importance(self, time_idx, node_idx, cut_time_l, group=None):
        """explainer.py:257-306 -> soft walk weights [B, W] (CUDA tensor); the statistics run over the call's batch (or `group` roots)."""
        t = self._t(time_idx, torch.float32); nodes = self._t(node_idx, torch.int32); cut = self._t(cut_time_l, torch.float32)
        B, W = t.shape[0], t.shape[1]
        deg = self.node_degree.detach().to(self.device, torch.float32).contiguous()
        w = torch.empty((B, W), dtype=torch.float32, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(lib().tm_walk_importance(B, W, int(group or self.batch_group or max(B, 1)), ptr(t), ptr(nodes), ptr(cut), ptr(deg), deg.shape[0], ptr(w), st),
              "tm_walk_importance")
        return w

    def enhance_predict_walks(self, walks, cut_time_l, edge_identify):
        """explainer.py:222-255 -> [B, hid_dim (+ 12)] CUDA tensor: the attention output of every walk (the scorer kernel with its
        hidden-vector output), weighted by compute_walk_importance, summed over the walks; class counts appended with if_cat."""
        nodes, eidx, t, cat, cut, eid = self._walk_tensors(walks, cut_time_l, edge_identify)
        B, W = nodes.shape[0], nodes.shape[1]
        group = int(self.batch_group or max(B, 1))
        if self._wants_grad() or self._dropout_active():                      # enhance_main.py:321-357 trains through the walk embeddings
            y = _tr.attention_layerwise(self, nodes.long(), eidx.long(), t, cut, eid)
            w = self.compute_walk_importance(t, nodes, cut, group=group)     # a function of timestamps and degrees only: no parameters
            emb = (y * w.unsqueeze(-1)).sum(1)
            if self.if_cat:
                emb = torch.cat([emb, torch.nn.functional.one_hot(cat.long(), 12).sum(1).to(emb.dtype)], dim=-1)
            return emb
        blob = self.packed_weights()
        nf, ef = self._tables()
        desc, ef = self._edge_table(blob, ef)
        desc = self._with_fanout(desc, self.detect_fanout(eidx, nodes, min_motifs=8192))
        self._workspace(B, W, group)
        scores = torch.empty((B, W), dtype=torch.float32, device=self.device)
        y = torch.empty((B, W, self.hid_dim), dtype=torch.float32, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(lib().tm_encode_attention(C.byref(desc), ptr(blob), B, W, group, ptr(nodes), ptr(eidx), ptr(t), ptr(cat), ptr(cut), ptr(eid),
                                        ptr(nf), nf.shape[0], ptr(ef), ef.shape[0], ptr(self._ws), ptr(scores), ptr(y), self.device.index, st),
              "tm_encode_attention")
        w = self.compute_walk_importance(t, nodes, cut, group=group)
        a3 = self.attention.MLP[3] if self.use_temporal_guidance else self.attention.MLP[2]
        a3w = a3.weight.detach().to(self.device, torch.float32).contiguous(); a3b = a3.bias.detach().to(self.device, torch.float32).contiguous()
        out = torch.empty((B, self.hid_dim + (12 if self.if_cat else 0)), dtype=torch.float32, device=self.device)
        check(lib().tm_enhance_reduce(B, W, self.hid_dim, ptr(y), ptr(w), ptr(a3w), ptr(a3b), ptr(cat), ptr(out), st), "tm_enhance_reduce")
        return out

    def enhance_predict_pairs(self, walks_src, walks_tgt, cut_time_l, src_edge, tgt_edge):
        return self.enhance_predict_walks(walks_src, cut_time_l, src_edge), self.enhance_predict_walks(walks_tgt, cut_time_l, tgt_edge)

    def enhance_predict_agg(self, ts_l_cut, walks_src, walks_tgt, walks_bgd, edge_id_info, src_gat, tgt_gat, bgd_gat):
        """explainer.py:203-213 -> (pos_score, neg_score), each [B, 1]."""
        src_edge, tgt_edge, bgd_edge = edge_id_info
        gat = [g.to(self.device, torch.float32) if isinstance(g, torch.Tensor) else torch.as_tensor(np.asarray(g), dtype=torch.float32, device=self.device)
               for g in (src_gat, tgt_gat, bgd_gat)]
        src_emb, tgt_emb = self.enhance_predict_pairs(walks_src, walks_tgt, ts_l_cut, src_edge, tgt_edge)
        pos = self.affinity_score(torch.cat([src_emb, gat[0]], dim=-1), torch.cat([tgt_emb, gat[1]], dim=-1))
        src_emb, bgd_emb = self.enhance_predict_pairs(walks_src, walks_bgd, ts_l_cut, src_edge, bgd_edge)
        neg = self.affinity_score(torch.cat([src_emb, gat[0]], dim=-1), torch.cat([bgd_emb, gat[2]], dim=-1))
        return pos, neg

    # ------------------------------------------------------------------ motif -> edge aggregation (explainer.py:354-430)
    def _packed_gate(self):
        """edge_dependency_gcn + time_encoder packed for the tensor-core gate kernel (None without dependency-aware sampling)."""
        if not self.use_dependency_aware_sampling:
            return None
        g = self.edge_dependency_gcn
        ps = [g[0].weight, g[0].bias, g[3].weight, g[3].bias, g[6].weight, g[6].bias, self.time_encoder.basis_freq, self.time_encoder.phase]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._gate_blob is None or key != self._gate_key:
            host = [p.detach().to("cpu", torch.float32).contiguous() for p in ps]
            prm = GateParams(*[C.c_void_p(h.data_ptr()) for h in host])
            blob = torch.empty(lib().tm_gate_blob_floats(C.byref(self._gate_desc)), dtype=torch.float32)
            check(lib().tm_gate_pack(C.byref(self._gate_desc), C.byref(prm), ptr(blob)), "tm_gate_pack")
            self._gate_blob = blob.to(self.device)
            self._gate_key = key
        return self._gate_blob

    def beta_sample(self, prob, training):
        """explainer.py:421-430: Beta(max(10 p, 1), max(10 (1 - p), 1)) -- its mean in eval, a reparameterised sample when training
        (device Philox sampler; the draws follow torch.manual_seed through the key of each call)."""
        if training:
            return _tr.BetaRSample.apply(prob if isinstance(prob, torch.Tensor) else self._t(prob, torch.float32), None, _tr.next_seed(), 0)
        alpha = torch.clamp(prob * 10, min=1.0)
        beta = torch.clamp((1 - prob) * 10, min=1.0)
        return alpha / (alpha + beta)

    def edge_importance_device(self, scores, eidx, t, h0_node, h0_eidx, h1_node, h1_eidx, training=False, seed=0):
        """All CUDA tensors: scores f32 [B,W], eidx i32 [B,W,3], t f32 [B,W,3], hop slots i32 [B,K0] / [B,K1] -> (imp0 [B,K0], imp1 [B,K1]).
        training: one Beta draw per slot inside the aggregation kernel (key = seed) instead of the Beta mean; values only, no graph."""
        B, W = eidx.shape[0], eidx.shape[1]
        K0, K1 = h0_eidx.shape[1], h1_eidx.shape[1]
        gate = self._packed_gate()
        _, ef = self._tables()
        walk_imp = torch.empty(B * W * 3, dtype=torch.float32, device=self.device) if gate is not None else None
        imp0 = torch.empty((B, K0), dtype=torch.float32, device=self.device)
        imp1 = torch.empty((B, K1), dtype=torch.float32, device=self.device)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(lib().tm_edge_importance(C.byref(self._gate_desc), ptr(gate) if gate is not None else None, B, W, ptr(scores), ptr(eidx), ptr(t),
                                       ptr(ef), ef.shape[0], K0, ptr(h0_node), ptr(h0_eidx), K1, ptr(h1_node), ptr(h1_eidx),
                                       ptr(walk_imp) if walk_imp is not None else None, ptr(imp0), ptr(imp1), int(bool(training)), int(seed),
                                       self.device.index, st),
              "tm_edge_importance")
        return imp0, imp1

    def retrieve_edge_imp_node(self, subgraph, graphlet_imp, walks, training=True):
        """explainer.py:354-406: dependency gate, per-root scatter-max over edge ids, gather to the hop-1 / hop-2 slots, Beta sample
        (training=True, the reference's default and what its eval loops pass, temp_exp_main.py:312-318) or Beta mean, padding mask.
        Returns CUDA tensors (edge_imp_0 [B,n], edge_imp_1 [B,n^2]).  When graphlet_imp carries an autograd graph (training loop) the
        result does too; otherwise the fused aggregation kernels run.  In train() mode the gate's dropout is part of the value, so
        that case also takes the layer-by-layer route."""
        node_record, eidx_record = subgraph[0], subgraph[1]
        eidx = self._t(walks[1], torch.int32)
        B, W = eidx.shape[0], eidx.shape[1]
        t = self._t(walks[2], torch.float32)                               # .float(), explainer.py:371
        h_nodes = [self._t(node_record[0], torch.int32), self._t(node_record[1], torch.int32)]
        h_eidx = [self._t(eidx_record[0], torch.int32), self._t(eidx_record[1], torch.int32)]
        g_imp = graphlet_imp if isinstance(graphlet_imp, torch.Tensor) else self._t(graphlet_imp, torch.float32)
        gate_dropout = self.use_dependency_aware_sampling and self._dropout_active()
        want = (self.training or self.autograd_in_eval) and torch.is_grad_enabled() and (
            g_imp.requires_grad or (self.use_dependency_aware_sampling and any(p.requires_grad for p in self.edge_dependency_gcn.parameters())))
        if want or gate_dropout:
            imp0, imp1 = _tr.edge_importance_autograd(self, g_imp.to(self.device, torch.float32), eidx.long(), t, [x.long() for x in h_nodes],
                                                      [x.long() for x in h_eidx], bool(training))
            return imp0, imp1
        scores = g_imp.detach().to(self.device, torch.float32).reshape(B, W).contiguous()
        return self.edge_importance_device(scores, eidx, t, h_nodes[0], h_eidx[0], h_nodes[1], h_eidx[1], training=bool(training),
                                           seed=_tr.next_seed() if training else 0)

    def kl_loss(self, prob, walks, target=0.3):
        """explainer.py:432-453 as a 0-dim CUDA tensor; carries an autograd graph (tm_kl_loss_backward) when ``prob`` does.
        The classes are paired with ``list(self.null_model.values())`` by position, as the reference does."""
        cat = self._t(walks[3], torch.uint8)
        B, W = cat.shape[0], cat.shape[1]
        cat = cat.reshape(B, W).contiguous()
        empirical = self.prior == "empirical"
        null = torch.tensor([float(v) for v in self.null_model.values()], dtype=torch.float32, device=self.device)
        p = prob if isinstance(prob, torch.Tensor) else self._t(prob, torch.float32)
        return _tr.KLLoss.apply(p.to(self.device), cat, null, float(target), int(empirical))

    def retrieve_explanation(self, subgraph_src, graphlet_imp_src, walks_src, subgraph_tgt, graphlet_imp_tgt, walks_tgt,
                             subgraph_bgd, graphlet_imp_bgd, walks_bgd, training=True):
        """explainer.py:408-419."""
        src_0, src_1 = self.retrieve_edge_imp_node(subgraph_src, graphlet_imp_src, walks_src, training=training)
        tgt_0, tgt_1 = self.retrieve_edge_imp_node(subgraph_tgt, graphlet_imp_tgt, walks_tgt, training=training)
        bgd_0, bgd_1 = self.retrieve_edge_imp_node(subgraph_bgd, graphlet_imp_bgd, walks_bgd, training=training)
        if self.base_type == "tgn":
            return [torch.cat([src_0, tgt_0, bgd_0], dim=0), torch.cat([src_1, tgt_1, bgd_1], dim=0)]
        return [torch.cat([src_0, tgt_0, bgd_0], dim=0)]

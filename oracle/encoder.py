"""numpy restatement of ``TempME.forward`` -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/models/explainer.py line by line (eval mode: every Dropout is the
identity).  ``dtype=np.float32`` mirrors the reference's arithmetic type; ``dtype=np.float64``
is the arbiter used to decide which of two fp32 results is closer to the exact value.

Parity status: PINNED against the unmodified reference module (run on CPU with a
``torch_scatter`` stand-in) by tests/golden/make_golden.py -> tests/golden/encoder_*.npz.

``params`` uses the reference's ``state_dict`` names:
  event_conv.lin_event.{weight,bias}, event_conv.MLP.{0,2}.{weight,bias},
  attention.{W1,W2}.{weight,bias}, attention.MLP.{0,3}.{weight,bias},
  MLP.{0,3,5}.{weight,bias}, time_encoder.{basis_freq,phase}
"""
from __future__ import annotations

import numpy as np


def _lin(x, p, name, dt):
    # nn.Linear: x @ W^T + b
    return x @ p[name + ".weight"].astype(dt).T + p[name + ".bias"].astype(dt)


def time_encode(ts, p, dt, arg32=False):
    """TimeEncode.forward, explainer.py:51-59: cos(ts * basis_freq + phase).  arg32 (with dt = float64): the ARGUMENT is formed in fp32 as
    the reference does -- at |argument| ~ 1e8 its fp32 rounding is part of the function's definition -- and only the cosine and
    everything downstream run in float64: the arbiter for "which fp32 result is closer to the exact value of the reference's formula"."""
    if arg32:
        m = ts.astype(np.float32)[..., None] * p["time_encoder.basis_freq"].astype(np.float32)
        m = (m + p["time_encoder.phase"].astype(np.float32)).astype(dt)
        return np.cos(m)
    m = ts[..., None] * p["time_encoder.basis_freq"].astype(dt)
    m = m + p["time_encoder.phase"].astype(dt)
    return np.cos(m)


def event_gcn(ev, src, tgt, p, dt):
    """event_gcn.forward, explainer.py:85-96."""
    event = _lin(ev, p, "event_conv.lin_event", dt)
    msg = np.maximum(tgt + event, 0)
    h = np.maximum(_lin(src + msg, p, "event_conv.MLP.0", dt), 0)
    return _lin(h, p, "event_conv.MLP.2", dt)


def temporal_attention(feat, time_idx, cut_time, p, dt, use_temporal=True):
    """TemporalAwareAttention.forward, explainer.py:789-846 (Attention.forward :25-43 when not use_temporal)."""
    src = feat[:, :, 2, :]                      # [B, W, 2H]   :799
    tgt = feat[:, :, 0:2, :]                    # [B, W, 2, 2H] :800
    Wp = _lin(src, p, "attention.W1", dt)       # :806
    Wq = _lin(tgt, p, "attention.W2", dt)       # :807
    scores = np.einsum("bwd,bwkd->bwk", Wp, Wq)  # bmm, :808
    if use_temporal:
        sel = time_idx[:, :, :2]                # :820
        td = np.abs(cut_time[:, None, None] - sel)          # :826
        std = td.astype(np.float64).std(ddof=1) if td.size > 1 else np.float64("nan")   # torch .std() is unbiased, :828
        tw = np.exp(-td / (dt(std) + dt(1e-6)))
        scores = scores * (dt(1.0) - dt(0.3) + dt(0.3) * tw)    # :835-836
    scores = scores - scores.max(-1, keepdims=True)
    e = np.exp(scores)
    alpha = e / e.sum(-1, keepdims=True)        # softmax :839
    out = np.einsum("bwk,bwkd->bwd", alpha, Wq)  # bmm :841
    out = src + out                             # :842
    h = np.maximum(_lin(out, p, "attention.MLP.0", dt), 0)
    last = "attention.MLP.3" if "attention.MLP.3.weight" in p else "attention.MLP.2"   # Attention.MLP has no Dropout, :18
    return _lin(h, p, last, dt)                 # [B, W, H]


def forward(p, node_feat, edge_feat, walks, cut_time, edge_identify, dtype=np.float32,
            use_temporal=True, if_cat=True, return_hidden=False, attention_only=False, arg32=False):
    """TempME.forward, explainer.py:174-201.

    walks = (node_idx [B,W,6], edge_idx [B,W,3], time_idx [B,W,3], cat_feat [B,W,1] or [B,W], _)
    """
    dt = dtype
    node_idx, edge_idx, time_idx, cat_feat = walks[0], walks[1], walks[2], walks[3]
    node_feat = np.asarray(node_feat).astype(dt); edge_feat = np.asarray(edge_feat).astype(dt)
    # timestamps and cut times go through .float() first (:325, :814-816)
    t32 = np.asarray(time_idx).astype(np.float32)
    cut32 = np.asarray(cut_time).astype(np.float32)
    edge_features = edge_feat[np.asarray(edge_idx).astype(np.int64)]            # :332-338
    edge_count = np.asarray(edge_identify).astype(np.float32).astype(dt)        # :177
    delta = (t32[:, :, 2:3] - t32)                                               # :326 (fp32 subtraction)
    B, W = delta.shape[:2]
    time_features = time_encode(delta.astype(dt), p, dt, arg32)                  # :328-329
    ev = np.concatenate([edge_features, edge_count, time_features], axis=-1)    # :179
    nid = np.asarray(node_idx).astype(np.int64)
    srcf = node_feat[nid[:, :, [0, 2, 4]]]                                       # :348-351
    tgtf = node_feat[nid[:, :, [1, 3, 5]]]
    up_src = event_gcn(ev, srcf, tgtf, p, dt)                                    # :182
    up_tgt = event_gcn(ev, tgtf, srcf, p, dt)                                    # :184
    feat = np.concatenate([up_src, up_tgt], axis=-1)                             # :185
    h = temporal_attention(feat, t32.astype(dt), cut32.astype(dt), p, dt, use_temporal)   # :190-193
    if attention_only:
        return h                                                                 # [B, W, H] (enhance_predict_walks :240-243)
    if if_cat:
        cat = np.asarray(cat_feat).astype(np.int64).reshape(B, W)
        onehot = np.eye(12, dtype=dt)[cat]                                       # :308-315
        h = np.concatenate([h, onehot], axis=-1)                                 # :197
    z = np.maximum(_lin(h, p, "MLP.0", dt), 0)
    z = np.maximum(_lin(z, p, "MLP.3", dt), 0)
    z = _lin(z, p, "MLP.5", dt)
    out = 1.0 / (1.0 + np.exp(-z))                                               # :200
    out = out.astype(dt)
    return (out, h) if return_hidden else out


def beta_mean(prob, dt=np.float32):
    """TempME.beta_sample in eval mode, explainer.py:421-430: E[Beta(max(10 p, 1), max(10 (1 - p), 1))]."""
    alpha = np.maximum(prob * dt(10), dt(1.0))
    beta = np.maximum((dt(1) - prob) * dt(10), dt(1.0))
    return alpha / (alpha + beta)


def edge_importance(p, edge_feat, subgraph, graphlet_imp, walks, use_dependency=True, dtype=np.float32):
    """TempME.retrieve_edge_imp_node, explainer.py:354-406, eval mode (training=False).

    subgraph = (node_records [hop0 [B,n], hop1 [B,n^2]], eidx_records [same shapes], _); graphlet_imp [B,W,1] scores;
    walks as in ``forward``.  Returns (edge_imp_0 [B,n], edge_imp_1 [B,n^2])."""
    dt = dtype
    node_record, eidx_record = subgraph[0], subgraph[1]
    edge_walk = np.asarray(walks[1]).reshape(walks[1].shape[0], -1).astype(np.int64)       # [B, 3W]            :360
    B = edge_walk.shape[0]
    walk_imp = np.repeat(np.asarray(graphlet_imp).astype(dt).reshape(B, -1, 1), 3, axis=2).reshape(B, -1)   # :363
    if use_dependency:
        ef = np.asarray(edge_feat).astype(dt)[edge_walk]                                   # :369
        tw = np.asarray(walks[2]).reshape(B, -1).astype(np.float32).astype(dt)             # :371 (.float())
        te = time_encode(tw, p, dt)                                                         # :372 (the raw timestamps)
        x = np.concatenate([ef, te], axis=-1)                                               # :375
        h = np.maximum(_lin(x, p, "edge_dependency_gcn.0", dt), 0)                          # :143-151, Dropout = identity
        h = np.maximum(_lin(h, p, "edge_dependency_gcn.3", dt), 0)
        dep = _lin(h, p, "edge_dependency_gcn.6", dt)[..., 0]                               # :379
        gate = dt(1) / (dt(1) + np.exp(-dep))                                               # :383
        walk_imp = walk_imp * (dt(0.5) + dt(0.5) * gate)                                    # :386
    outs = []
    for l in range(2):
        ids = np.asarray(eidx_record[l]).astype(np.int64)
        imp = np.zeros(ids.shape, dt)
        for b in range(B):                                                                  # scatter(max) over edge ids, then gather (:389-393)
            m = {}
            for e, v in zip(edge_walk[b], walk_imp[b]):
                m[e] = max(m.get(e, dt(0)), v)
            imp[b] = [m.get(e, dt(0)) for e in ids[b]]
        imp = beta_mean(imp, dt)                                                            # :396-397
        imp[np.asarray(node_record[l]) == 0] = 0                                            # :400-404
        outs.append(imp.astype(dt))
    return outs[0], outs[1]


def walk_importance(time_idx, node_idx, cut_time, node_degree, dtype=np.float32):
    """TempME.compute_walk_importance, explainer.py:257-306: soft weights [B, W] from recency (batch-global std) and the mean degree
    of the walk's non-padding nodes (batch-global mean / std), normalised to sum to W per root."""
    dt = dtype
    t = np.asarray(time_idx).astype(np.float32).astype(dt)
    nid = np.asarray(node_idx).astype(np.int64)
    cut = np.asarray(cut_time).astype(np.float32).astype(dt)
    W = t.shape[1]
    diff = np.abs(cut[:, None] - t.max(-1))                                              # :274-277
    rec = np.exp(-diff / (dt(diff.astype(np.float64).std(ddof=1)) + dt(1e-6)) / dt(1.0))  # :281
    valid = nid > 0
    deg = np.where(valid, np.asarray(node_degree).astype(dt)[nid], dt(0))               # :286-291
    avg = deg.sum(-1) / (valid.sum(-1).astype(dt) + dt(1e-6))                            # :292
    z = (avg - dt(avg.astype(np.float64).mean())) / (dt(avg.astype(np.float64).std(ddof=1)) + dt(1e-6))
    dw = dt(1) / (dt(1) + np.exp(-z))                                                     # :295
    imp = dt(0.5) * rec + dt(0.5) * dw                                                    # :298
    return (imp / (imp.sum(-1, keepdims=True) / dt(W) + dt(1e-6))).astype(dt)             # :301


def enhance_predict_walks(p, node_feat, edge_feat, walks, cut_time, edge_identify, node_degree, dtype=np.float32, use_temporal=True, if_cat=True,
                          arg32=False):
    """TempME.enhance_predict_walks, explainer.py:222-255: attention output per walk, weighted by walk_importance, summed over the
    walks; with if_cat the per-root class counts are appended (:307-313)."""
    dt = dtype
    h = forward(p, node_feat, edge_feat, walks, cut_time, edge_identify, dtype=dt, use_temporal=use_temporal, if_cat=if_cat, attention_only=True, arg32=arg32)
    w = walk_importance(walks[2], walks[0], cut_time, node_degree, dt)
    out = (h * w[..., None]).sum(1)                                                       # :245-249
    if if_cat:
        cat = np.asarray(walks[3]).astype(np.int64).reshape(h.shape[0], h.shape[1])
        out = np.concatenate([out, np.eye(12, dtype=dt)[cat].sum(1)], axis=-1)            # :251-253
    return out.astype(dt)


def affinity_score(p, x1, x2, dtype=np.float32):
    """_MergeLayer.forward, explainer.py:71-76."""
    dt = dtype
    x = np.concatenate([x1, x2], axis=-1).astype(dt)
    return _lin(np.maximum(_lin(x, p, "affinity_score.fc1", dt), 0), p, "affinity_score.fc2", dt)


def kl_loss(prob, cat_feat, null_values, target=0.3, prior="empirical", dtype=np.float32):
    """TempME.kl_loss, explainer.py:432-453.  prob [B,W] scores, cat_feat [B,W] classes, null_values = list(null_model.values())."""
    dt = dtype
    cat = np.asarray(cat_feat).astype(np.int64).reshape(np.asarray(cat_feat).shape[0], -1)
    B, W = cat.shape
    p = np.clip(np.asarray(prob).astype(dt).reshape(B, W), dt(1e-6), dt(1 - 1e-6))          # :435
    t = dt(target)
    if prior != "empirical":
        v = p * np.log(p / t + dt(1e-6)) + (dt(1) - p) * np.log((dt(1) - p) / (dt(1) - t + dt(1e-6)) + dt(1e-6))   # :450-451
        return dt(v.astype(np.float64).mean())
    null = np.asarray(null_values).astype(dt)
    C = null.shape[0]
    s = p.mean(1, keepdims=True)                                                              # :438
    onehot = (cat[:, :, None] == np.arange(C)[None, None, :])
    cnt = np.maximum(onehot.sum(1), 1).astype(dt)
    emp = s * ((p[:, :, None] * onehot).sum(1) / cnt)                                         # :443-444 (scatter mean: absent class -> 0)
    nd = t * null[None, :]                                                                    # :445
    v = (dt(1) - s) * np.log((dt(1) - s) / (dt(1) - t + dt(1e-6)) + dt(1e-6)) + emp * np.log(emp / (nd + dt(1e-6)) + dt(1e-6))
    return dt(v.astype(np.float64).mean())                                                    # :447-448

"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The reference's randomness is replaced by the deterministic-draw
contract through tests/refshim.py (no reference file is modified or copied); functions that live
in the un-importable script processed/data_preprocess.py (it needs h5py and runs on import) are
pulled out of that file with ``ast`` at run time and executed as they are.

Fixtures:
  tie_star.npz     SURVEY App. A.2 example: CSR arrays, nodeedge2idx values, find_before windows
  rand_small.npz   40-node multigraph with ties, self-loops and a real node 0: k-hop + walks (+offset shard)
  native_rng.npz   the reference driven by its own numpy MT19937 stream, with the draws it consumed recorded
  rand_bigts.npz   timestamps ~1e8 (float32 != float64), train/full split: e_idx missing from a finder
  uslegis.npz      the bundled processed/ml_uslegis_sampled events + reference outputs on test queries
  nullmodel.npz    utils/null_model.py pre_processing on an endpoint-shuffled copy (class histogram)
  encoder_*.npz    TempME.forward scores with the weights/features that produced them
  edgeimp_*.npz    retrieve_edge_imp_node (eval mode) on those scores: `python tests/golden/make_golden.py edgeimp`
  enhance_*.npz    enhance_predict_walks / compute_walk_importance / enhance_predict_agg (eval): `python tests/golden/make_golden.py enhance`
  kl_loss.npz      TempME.kl_loss on fixed scores / classes, both priors: `python tests/golden/make_golden.py kl`
  nextstep_time.npz  get_next_step(e_idx_l=None): second events drawn from time-cut prefixes: `python tests/golden/make_golden.py nextstep`
  trainstep_*.npz  one temp_exp_main.py:605-632-shaped step in train() mode with dropout_p = 0 (deterministic): scores, edge importances,
                   kl_loss, the loss and d loss / d parameter for every explainer parameter: `python tests/golden/make_golden.py train`
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refshim  # noqa: E402

rg = refshim.import_reference()
REF = refshim.REF


def ref_functions_from_script(path, names):
    """exec selected top-level ``def``s of a reference script without importing/running the script."""
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"np": np, "tqdm": (lambda x: x)}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


marginal, new_edge_info = ref_functions_from_script(
    os.path.join(REF, "processed", "data_preprocess.py"), ["marginal", "new_edge_info"])


def khop_and_walks(nf, shim, roots, ts, eidx, n, N2, k=2):
    """One root type: find_k_hop(k) then find_k_walks -- two counted top-level calls."""
    sub = nf.find_k_hop(k, roots, ts, n, e_idx_l=eidx)
    walks = nf.find_k_walks(n, roots, N2, sub)
    return sub, walks


def pack_sub(prefix, sub, out):
    for name, rec in zip(("node", "eidx", "ts"), sub):
        for l, a in enumerate(rec):
            out[f"{prefix}_hop{l}_{name}"] = a


def pack_walks(prefix, walks, out):
    for name, a in zip(("nodes", "eidx", "t", "anony"), walks):
        out[f"{prefix}_w_{name}"] = a


def dict_table(nf, n_nodes):
    rows = []
    for v in range(n_nodes):
        for e, val in nf.nodeedge2idx[v].items():
            rows.append((v, e, val))
    return np.array(rows, np.int64).reshape(-1, 3)


def gen_tie_star():
    # node 1 is the hub; ts = [1,2,2,2,3,3], eidx 1..6
    src = np.array([1] * 6); dst = np.array([2, 3, 4, 5, 6, 7]); eidx = np.arange(1, 7)
    ts = np.array([1, 2, 2, 2, 3, 3], np.float64)
    nf = rg.NeighborFinder(refshim.adj_list_from_events(8, src, dst, eidx, ts))
    out = dict(n_nodes=8, src=src, dst=dst, eidx=eidx, ts=ts, off=nf.off_set_l, nbr=nf.node_idx_l,
               e=nf.edge_idx_l, t=nf.node_ts_l, dict=dict_table(nf, 8))
    q = []
    for e in range(1, 7):
        a = nf.find_before(1, 3.0, e_idx=e)
        q.append((1, e, len(a[0])))
    out["fb_eidx"] = np.array(q, np.int64)
    out["fb_time"] = np.array([(1, t, len(nf.find_before(1, t)[0])) for t in (0.5, 1.0, 2.0, 2.5, 3.0, 9.0)], np.float64)
    np.savez_compressed(os.path.join(HERE, "tie_star.npz"), **out)


def gen_rand_small():
    rng = np.random.default_rng(1)
    N, E = 40, 600
    src = rng.integers(0, N, E); dst = rng.integers(0, N, E)
    ts = np.sort(rng.integers(0, 60, E)).astype(np.float64)
    eidx = np.arange(1, E + 1)
    nf = rg.NeighborFinder(refshim.adj_list_from_events(N, src, dst, eidx, ts))
    q = np.arange(E - 150, E - 1)[:48]
    fake = rng.integers(0, N, len(q))
    n, N2 = 7, 3
    out = dict(n_nodes=N, src=src, dst=dst, eidx=eidx, ts=ts, q=q, fake=fake, n=n, N2=N2, base_seed=77,
               off=nf.off_set_l, nbr=nf.node_idx_l, e=nf.edge_idx_l, t=nf.node_ts_l, dict=dict_table(nf, N))
    shim = refshim.DrawShim(base_seed=77)
    with shim.patched(rg):
        for name, roots, e in (("src", src[q], eidx[q]), ("tgt", dst[q], eidx[q]), ("bgd", fake, None)):
            sub, walks = khop_and_walks(nf, shim, roots, ts[q], e, n, N2)
            pack_sub(name, sub, out); pack_walks(name, walks, out)
    # the same src roots as the second half of a 2-way shard: row_offset = 24, fresh call numbering
    shim = refshim.DrawShim(base_seed=77, row_offset=24)
    with shim.patched(rg):
        sub, walks = khop_and_walks(nf, shim, src[q][24:], ts[q][24:], eidx[q][24:], n, N2)
        pack_sub("shard1", sub, out); pack_walks("shard1", walks, out)
    # N2 = 1 (null-model setting) and a wide fan-out n = 40 > 32
    shim = refshim.DrawShim(base_seed=5)
    with shim.patched(rg):
        sub, walks = khop_and_walks(nf, shim, dst[q], ts[q], eidx[q], 40, 1, k=1)
        pack_sub("wide", sub, out); pack_walks("wide", walks, out)
    out["src_edge_identity"] = new_edge_info(out["src_w_eidx"].astype(int))
    np.savez_compressed(os.path.join(HERE, "rand_small.npz"), **out)


def gen_native_rng():
    """The reference with its OWN numpy MT19937 stream (np.random.seed(12345)); the draws are only recorded."""
    rng = np.random.default_rng(21)
    N, E = 50, 1200
    src = rng.integers(0, N, E); dst = rng.integers(0, N, E)
    ts = np.sort(rng.integers(0, 300, E)).astype(np.float64)
    eidx = np.arange(1, E + 1)
    nf = rg.NeighborFinder(refshim.adj_list_from_events(N, src, dst, eidx, ts))
    q = np.arange(E - 400, E - 1)[::7][:40]
    fake = rng.integers(0, N, len(q))
    n, N2 = 9, 3
    out = dict(n_nodes=N, src=src, dst=dst, eidx=eidx, ts=ts, q=q, fake=fake, n=n, N2=N2)
    np.random.seed(12345)
    for name, roots, e in (("src", src[q], eidx[q]), ("bgd", fake, None)):
        rec = refshim.DrawRecorder()
        with rec.patched():
            sub = nf.find_k_hop(2, roots, ts[q], n, e_idx_l=e)
            walks = nf.find_k_walks(n, roots, N2, sub)
        pack_sub(name, sub, out); pack_walks(name, walks, out)
        B = len(q)
        out[f"{name}_inj_hop0"] = rec.dense(("get_temporal_neighbor", 0), B, n)
        out[f"{name}_inj_hop1"] = rec.dense(("get_temporal_neighbor", 1), B * n, n)
        out[f"{name}_inj_step2"] = rec.dense(("get_next_step", 0), B * n, N2)
        out[f"{name}_inj_step3"] = rec.dense(("get_final_step", 0), B * n * N2, 1)
    np.savez_compressed(os.path.join(HERE, "native_rng.npz"), **out)


def gen_rand_bigts():
    rng = np.random.default_rng(2)
    N, E = 60, 2000
    src = rng.integers(1, N, E); dst = rng.integers(1, N, E)
    keep = src != dst
    src, dst = src[keep], dst[keep]; E = len(src)
    ts = np.sort(rng.integers(100_000_000, 100_020_000, E)).astype(np.float64)
    ts[rng.random(E) < 0.3] += 0.0  # keep integers; ~duplicates come from the narrow range
    eidx = np.arange(1, E + 1)
    n_train = int(E * 0.7)
    nf_train = rg.NeighborFinder(refshim.adj_list_from_events(N, src[:n_train], dst[:n_train], eidx[:n_train], ts[:n_train]))
    nf_full = rg.NeighborFinder(refshim.adj_list_from_events(N, src, dst, eidx, ts))
    q = np.arange(n_train + 10, n_train + 10 + 32)
    fake = rng.integers(1, N, len(q))
    n, N2 = 10, 2
    out = dict(n_nodes=N, src=src, dst=dst, eidx=eidx, ts=ts, n_train=n_train, q=q, fake=fake, n=n, N2=N2, base_seed=9)
    shim = refshim.DrawShim(base_seed=9)
    with shim.patched(rg):
        for name, roots, e in (("src", src[q], eidx[q]), ("tgt", dst[q], eidx[q]), ("bgd", fake, None)):
            sub, walks = khop_and_walks(nf_full, shim, roots, ts[q], e, n, N2)
            pack_sub(name, sub, out); pack_walks(name, walks, out)
        # train finder: query e_idx are not in it -> only the time-cut (bgd-style) path is legal at hop 0
        sub, walks = khop_and_walks(nf_train, shim, src[q], ts[q], None, n, N2)
        pack_sub("train", sub, out); pack_walks("train", walks, out)
    try:
        nf_train.find_before(int(src[q][0]), float(ts[q][0]), e_idx=int(eidx[q][0]))
        out["train_raises"] = 0
    except IndexError:
        out["train_raises"] = 1
    np.savez_compressed(os.path.join(HERE, "rand_bigts.npz"), **out)
    return out


def load_uslegis():
    import pandas as pd
    g = pd.read_csv(os.path.join(REF, "processed", "ml_uslegis_sampled.csv"))
    return g.u.values, g.i.values, g.idx.values, g.ts.values


def gen_uslegis():
    src, dst, eidx, ts = load_uslegis()
    N = int(max(src.max(), dst.max())) + 1
    nf = rg.NeighborFinder(refshim.adj_list_from_events(N, src, dst, eidx, ts))
    test_time = np.quantile(ts, 0.85)
    q = np.nonzero(ts > test_time)[0][:40]
    rng = np.random.default_rng(3)
    fake = rng.integers(0, N, len(q))
    n, N2 = 30, 3
    out = dict(n_nodes=N, src=src.astype(np.int16), dst=dst.astype(np.int16), eidx=eidx.astype(np.int32),
               ts=ts.astype(np.float32), q=q, fake=fake, n=n, N2=N2, base_seed=2024)
    shim = refshim.DrawShim(base_seed=2024)
    allw = {}
    with shim.patched(rg):
        for name, roots, e in (("src", src[q], eidx[q]), ("tgt", dst[q], eidx[q]), ("bgd", fake, None)):
            sub, walks = khop_and_walks(nf, shim, roots, ts[q], e, n, N2)
            pack_sub(name, sub, out); pack_walks(name, walks, out)
            allw[name] = np.concatenate([w.astype(np.float64) for w in walks], axis=-1)  # [Q, W, 15], data_preprocess.py:130
    new = marginal(allw["src"], allw["tgt"], allw["bgd"])                                   # [Q, W, 14]
    for name, w in zip(("src", "tgt", "bgd"), new):
        out[f"{name}_cat"] = w[:, :, 12].astype(np.int8)
        out[f"{name}_marginal"] = w[:, :, 13]
        out[f"{name}_edge_identity"] = new_edge_info(w[:, :, 6:9].astype(int)).astype(np.int16)
    for k in list(out):
        if k.endswith("hop1_node") or k.endswith("hop1_eidx"):
            out[k] = out[k].astype(np.int16 if k.endswith("node") else np.int32)
    np.savez_compressed(os.path.join(HERE, "uslegis.npz"), **out)
    return out


def gen_nullmodel():
    import utils.null_model as rnull  # the reference's module
    src, dst, eidx, ts = load_uslegis()
    rng = np.random.default_rng(4)
    perm = rng.permutation(len(src))           # load_data_shuffle permutes src/dst but not ts/e_idx, null_model.py:23-29
    s2, d2 = src[perm], dst[perm]
    N = int(max(src.max(), dst.max())) + 1
    nf = rg.NeighborFinder(refshim.adj_list_from_events(N, s2, d2, eidx, ts))
    test = ts > np.quantile(ts, 0.85)
    fakes = []

    class Sampler:  # stands in for RandEdgeSampler.sample (batch_loader.py:39-42) and records what it returned
        def sample(self, size):
            f = rng.integers(0, N, size)
            fakes.append(f)
            return f, f

    n = rnull.degree_dict["uslegis_sampled"]
    shim = refshim.DrawShim(base_seed=31337)
    with shim.patched(rg):
        sat = rnull.pre_processing(nf, Sampler(), s2[test], d2[test], ts[test], eidx[test], n)
    out = dict(n_nodes=N, src=s2.astype(np.int16), dst=d2.astype(np.int16), eidx=eidx.astype(np.int32), ts=ts.astype(np.float32),
               test_idx=np.nonzero(test)[0][:500], fakes=np.stack(fakes), n=n, base_seed=31337,
               dist=np.array([sat[k] for k in range(1, 13)], np.float64))
    np.savez_compressed(os.path.join(HERE, "nullmodel.npz"), **out)


FWD_PARAMS = ["event_conv.lin_event", "event_conv.MLP.0", "event_conv.MLP.2", "attention.W1", "attention.W2",
              "attention.MLP.0", "attention.MLP.2", "attention.MLP.3", "MLP.0", "MLP.3", "MLP.5"]


def gen_encoder(tag, walks5, edge_identity, cut_time, n_nodes, n_edges, D, Ed, seed, use_temporal=True, zero_node=False, if_cat=True, hid=64):
    import torch
    import models.explainer as rexp  # the reference's module
    rexp.get_null_distribution = lambda data_name: {k: 1.0 / 12 for k in range(1, 13)}  # skip the 8 s CSV pass
    torch.manual_seed(seed)
    nfeat = torch.zeros(n_nodes, D) if zero_node else torch.randn(n_nodes, D)
    efeat = torch.randn(n_edges, Ed)
    nfeat[0] = 0; efeat[0] = 0  # padding rows of the base model tables

    class Base:
        n_feat_th = nfeat; e_feat_th = efeat
        node_raw_features = torch.nn.Embedding.from_pretrained(nfeat, padding_idx=0, freeze=True)
        edge_raw_features = torch.nn.Embedding.from_pretrained(efeat, padding_idx=0, freeze=True)

    m = rexp.TempME(Base(), "tgn", "uslegis_sampled", out_dim=40, hid_dim=hid, device=torch.device("cpu"),
                    use_temporal_guidance=use_temporal, if_cat_feature=if_cat)
    with torch.no_grad():  # move the trainable phase off zero so the +phase step is exercised
        m.time_encoder.phase.copy_(0.1 * torch.randn(D))
    m.eval()
    with torch.no_grad():
        score = m(walks5, cut_time, edge_identity).numpy()
    sd = m.state_dict()
    out = {"p:" + k: v.numpy() for k, v in sd.items()
           if any(k.startswith(p + ".") for p in FWD_PARAMS) or k.startswith("time_encoder.")}
    out.update(node_feat=nfeat.numpy(), edge_feat=efeat.numpy(), w_nodes=walks5[0].astype(np.int32),
               w_eidx=walks5[1].astype(np.int32), w_t=walks5[2].astype(np.float32), w_cat=walks5[3].astype(np.int8),
               cut_time=cut_time, edge_identity=edge_identity.astype(np.float32), score=score,
               use_temporal=int(use_temporal), if_cat=int(if_cat), hid_dim=int(hid))
    np.savez_compressed(os.path.join(HERE, f"encoder_{tag}.npz"), **out)


def gen_edge_imp(tag, walks5, edge_identity, cut_time, subgraph, n_nodes, n_edges, D, Ed, seed):
    """retrieve_edge_imp_node (reference models/explainer.py:354-406) in eval mode on the scores of the same module:
    dependency gate (edge_dependency_gcn over [edge features | TimeEncode(raw t)]), per-query scatter-max over edge ids,
    gather to the hop-1 / hop-2 slots, Beta mean, padding mask."""
    import torch
    import models.explainer as rexp
    rexp.get_null_distribution = lambda data_name: {k: 1.0 / 12 for k in range(1, 13)}
    torch.manual_seed(seed)
    nfeat = torch.randn(n_nodes, D); efeat = torch.randn(n_edges, Ed)
    nfeat[0] = 0; efeat[0] = 0

    class Base:
        n_feat_th = nfeat; e_feat_th = efeat
        node_raw_features = torch.nn.Embedding.from_pretrained(nfeat, padding_idx=0, freeze=True)
        edge_raw_features = torch.nn.Embedding.from_pretrained(efeat, padding_idx=0, freeze=True)

    out = {}
    for dep in (True, False):
        torch.manual_seed(seed + 100)
        m = rexp.TempME(Base(), "tgn", "uslegis_sampled", out_dim=40, hid_dim=64, device=torch.device("cpu"),
                        use_dependency_aware_sampling=dep)
        with torch.no_grad():
            m.time_encoder.phase.copy_(0.1 * torch.randn(D))
        m.eval()
        with torch.no_grad():
            score = m(walks5, cut_time, edge_identity)
            imp0, imp1 = m.retrieve_edge_imp_node(subgraph, score, walks5, training=False)
        k = "dep" if dep else "nodep"
        out.update({f"{k}_score": score.numpy(), f"{k}_imp0": imp0.numpy(), f"{k}_imp1": imp1.numpy()})
        if dep:
            sd = m.state_dict()
            out.update({"p:" + n: v.numpy() for n, v in sd.items() if n.startswith("edge_dependency_gcn.") or n.startswith("time_encoder.")})
    out.update(edge_feat=efeat.numpy(), w_eidx=walks5[1].astype(np.int32), w_t=walks5[2].astype(np.float32),
               h0_node=subgraph[0][0].astype(np.int32), h1_node=subgraph[0][1].astype(np.int32),
               h0_eidx=subgraph[1][0].astype(np.int32), h1_eidx=subgraph[1][1].astype(np.int32))
    np.savez_compressed(os.path.join(HERE, f"edgeimp_{tag}.npz"), **out)


def gen_enhance(tag, walks_src, walks_tgt, ei_src, ei_tgt, cut_time, n_nodes, n_edges, D, Ed, seed, hid=64, out_dim=40):
    """enhance_predict_walks / enhance_predict_agg (reference models/explainer.py:203-306) in eval mode: attention output per walk,
    soft walk-importance weights (batch-global statistics, node_degree gather), weighted sum over the walks, category counts,
    affinity score of the (src, tgt) and (src, bgd) pairs."""
    import torch
    import models.explainer as rexp
    rexp.get_null_distribution = lambda data_name: {k: 1.0 / 12 for k in range(1, 13)}
    torch.manual_seed(seed)
    nfeat = torch.randn(n_nodes, D); efeat = torch.randn(n_edges, Ed)
    nfeat[0] = 0; efeat[0] = 0

    class Base:
        n_feat_th = nfeat; e_feat_th = efeat
        node_raw_features = torch.nn.Embedding.from_pretrained(nfeat, padding_idx=0, freeze=True)
        edge_raw_features = torch.nn.Embedding.from_pretrained(efeat, padding_idx=0, freeze=True)

    m = rexp.TempME(Base(), "tgn", "uslegis_sampled", out_dim=out_dim, hid_dim=hid, device=torch.device("cpu"))
    with torch.no_grad():
        m.time_encoder.phase.copy_(0.1 * torch.randn(D))
        m.node_degree = torch.randint(1, 60, (n_nodes,)).float()
    m.eval()
    B = walks_src[0].shape[0]
    src_gat, tgt_gat, bgd_gat = torch.randn(B, D), torch.randn(B, D), torch.randn(B, D)
    with torch.no_grad():
        emb_src = m.enhance_predict_walks(walks_src, cut_time, ei_src)
        emb_tgt = m.enhance_predict_walks(walks_tgt, cut_time, ei_tgt)
        w_src = m.compute_walk_importance(walks_src[2], walks_src[0], cut_time)
        pos, neg = m.enhance_predict_agg(cut_time, walks_src, walks_tgt, walks_src, (ei_src, ei_tgt, ei_src), src_gat, tgt_gat, bgd_gat)
    sd = m.state_dict()
    out = {"p:" + k: v.numpy() for k, v in sd.items()
           if any(k.startswith(q + ".") for q in FWD_PARAMS) or k.startswith("time_encoder.") or k.startswith("affinity_score.")}
    for pre, w, ei in (("src", walks_src, ei_src), ("tgt", walks_tgt, ei_tgt)):
        out.update({f"{pre}_nodes": w[0].astype(np.int32), f"{pre}_eidx": w[1].astype(np.int32), f"{pre}_t": w[2].astype(np.float32),
                    f"{pre}_cat": w[3].astype(np.int8), f"{pre}_ei": ei.astype(np.float32)})
    out.update(node_feat=nfeat.numpy(), edge_feat=efeat.numpy(), node_degree=m.node_degree.numpy(), cut_time=cut_time,
               src_gat=src_gat.numpy(), tgt_gat=tgt_gat.numpy(), bgd_gat=bgd_gat.numpy(),
               emb_src=emb_src.numpy(), emb_tgt=emb_tgt.numpy(), w_src=w_src.numpy(), pos=pos.numpy(), neg=neg.numpy(), hid_dim=hid, out_dim=out_dim)
    np.savez_compressed(os.path.join(HERE, f"enhance_{tag}.npz"), **out)


def gen_enhance_all():
    big = dict(np.load(os.path.join(HERE, "rand_bigts.npz")))
    Bq = 12
    ws = []
    for pre in ("src", "tgt"):
        wn, we, wt, wa = (big[f"{pre}_w_{k}"][:Bq] for k in ("nodes", "eidx", "t", "anony"))
        allw = np.concatenate([x.astype(np.float64) for x in (wn, we, wt, wa)], axis=-1)
        new = marginal(allw, allw, allw)[0]
        ws.append(((wn.astype(np.int64), we.astype(np.int64), wt.astype(np.float64), new[:, :, 12:13].astype(np.int64), new[:, :, 13:14]),
                   new_edge_info(we.astype(int))))
    gen_enhance("d32", ws[0][0], ws[1][0], ws[0][1], ws[1][1], big["ts"][big["q"][:Bq]], int(big["n_nodes"]), len(big["eidx"]) + 1, 32, 32, seed=5)
    # enhance_main.py's own defaults: --hid_dim 32 --out_dim 32 (enhance_main.py:65-66)
    gen_enhance("d32_hid32", ws[0][0], ws[1][0], ws[0][1], ws[1][1], big["ts"][big["q"][:Bq]], int(big["n_nodes"]), len(big["eidx"]) + 1, 32, 32, seed=8,
                hid=32, out_dim=32)


def gen_encoder_nocat():
    """TempME(if_cat_feature=False): MLP over the attention output alone (mlp_dim = hid_dim); `python tests/golden/make_golden.py nocat`."""
    big = dict(np.load(os.path.join(HERE, "rand_bigts.npz")))
    Bq = 12
    wn, we, wt, wa = (big[f"src_w_{k}"][:Bq] for k in ("nodes", "eidx", "t", "anony"))
    allw = np.concatenate([x.astype(np.float64) for x in (wn, we, wt, wa)], axis=-1)
    new = marginal(allw, allw, allw)[0]
    walks5 = (wn.astype(np.int64), we.astype(np.int64), wt.astype(np.float64), new[:, :, 12:13].astype(np.int64), new[:, :, 13:14])
    gen_encoder("d32_nocat", walks5, new_edge_info(we.astype(int)), big["ts"][big["q"][:Bq]], big["n_nodes"], len(big["eidx"]) + 1, 32, 32, seed=6,
                if_cat=False)
    # enhance_main.py's defaults: hid_dim = 32 (enhance_main.py:65-66)
    gen_encoder("d32_hid32", walks5, new_edge_info(we.astype(int)), big["ts"][big["q"][:Bq]], big["n_nodes"], len(big["eidx"]) + 1, 32, 32, seed=7, hid=32)


def gen_kl_all():
    """TempME.kl_loss (reference models/explainer.py:432-453) on the classes of the committed walk fixtures and seeded scores (with
    exact 0 / 1 entries for the clamp), for both priors and two targets.  The null model is non-uniform, in the reference's dict order."""
    import torch
    import models.explainer as rexp
    rs = np.random.RandomState(11)
    null = {k: float(v) for k, v in zip(range(1, 13), rs.dirichlet(np.ones(12) * 0.7))}
    rexp.get_null_distribution = lambda data_name: null
    nfeat = torch.zeros(4, 8); efeat = torch.zeros(4, 8)

    class Base:
        n_feat_th = nfeat; e_feat_th = efeat
        node_raw_features = torch.nn.Embedding.from_pretrained(nfeat, padding_idx=0, freeze=True)
        edge_raw_features = torch.nn.Embedding.from_pretrained(efeat, padding_idx=0, freeze=True)

    us = dict(np.load(os.path.join(HERE, "uslegis.npz")))
    out = {"null_values": np.array(list(null.values()), np.float64)}
    cases = {"us": us["src_cat"][:16].astype(np.int64), "few": (rs.randint(0, 3, (5, 20)) * 4).astype(np.int64), "one": np.full((1, 7), 11, np.int64)}
    for name, cat in cases.items():
        B, W = cat.shape
        prob = rs.rand(B, W, 1).astype(np.float32) ** 2
        prob[0, 0] = 0.0; prob[-1, -1] = 1.0
        out[f"{name}_cat"] = cat.astype(np.int8); out[f"{name}_prob"] = prob
        for prior in ("empirical", "uniform"):
            m = rexp.TempME(Base(), "tgn", "x", out_dim=8, hid_dim=16, prior=prior, device=torch.device("cpu"))
            for target in (0.3, 0.05):
                with torch.no_grad():
                    v = m.kl_loss(torch.from_numpy(prob), (None, None, None, cat[:, :, None], None), target=target)
                out[f"{name}_{prior}_{target}"] = np.float64(v.item())
    np.savez_compressed(os.path.join(HERE, "kl_loss.npz"), **out)


def gen_train_step(tag, walks5, edge_identity, cut_time, subgraph, n_nodes, n_edges, D, Ed, seed, hid=64, prior="empirical", use_temporal=True):
    """forward -> retrieve_edge_imp_node(training=False: Beta mean, deterministic) -> kl_loss -> backward on the unmodified reference
    module in train() mode with every Dropout probability set to 0, as temp_exp_main.py:605-632 does per root type.  loss = sum(w0 imp0) + sum(w1 imp1) + 0.5 kl."""
    import torch
    import models.explainer as rexp
    rs = np.random.RandomState(seed)
    null = {k: float(v) for k, v in zip(range(1, 13), rs.dirichlet(np.ones(12) * 0.7))}
    rexp.get_null_distribution = lambda data_name: null
    torch.manual_seed(seed)
    nfeat = torch.randn(n_nodes, D); efeat = torch.randn(n_edges, Ed)
    nfeat[0] = 0; efeat[0] = 0

    class Base:
        n_feat_th = nfeat; e_feat_th = efeat
        node_raw_features = torch.nn.Embedding.from_pretrained(nfeat, padding_idx=0, freeze=True)
        edge_raw_features = torch.nn.Embedding.from_pretrained(efeat, padding_idx=0, freeze=True)

    m = rexp.TempME(Base(), "tgn", "uslegis_sampled", out_dim=40, hid_dim=hid, prior=prior, dropout_p=0.0, device=torch.device("cpu"),
                    use_temporal_guidance=use_temporal)
    with torch.no_grad():
        m.time_encoder.phase.copy_(0.1 * torch.randn(D))
    m.train()
    for mod in m.modules():       # dropout_p does not reach the attention module (it keeps its default 0.1, explainer.py:121): zero every Dropout
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    score = m(walks5, cut_time, edge_identity)
    imp0, imp1 = m.retrieve_edge_imp_node(subgraph, score, walks5, training=False)
    kl = m.kl_loss(score, walks5, target=0.3)
    w0 = torch.from_numpy(rs.rand(*imp0.shape).astype(np.float32)); w1 = torch.from_numpy(rs.rand(*imp1.shape).astype(np.float32))
    loss = (imp0 * w0).sum() + (imp1 * w1).sum() + 0.5 * kl
    loss.backward()
    out = {"p:" + k: v.detach().numpy() for k, v in m.state_dict().items()
           if not (k.startswith("node_raw_embed") or k.startswith("edge_raw_embed") or k.startswith("node_degree"))}
    out.update({"g:" + k: v.grad.numpy() for k, v in m.named_parameters() if v.grad is not None})
    out.update(node_feat=nfeat.numpy(), edge_feat=efeat.numpy(), w_nodes=walks5[0].astype(np.int32), w_eidx=walks5[1].astype(np.int32),
               w_t=walks5[2].astype(np.float32), w_cat=walks5[3].astype(np.int8), cut_time=cut_time, edge_identity=edge_identity.astype(np.float32),
               h0_node=subgraph[0][0].astype(np.int32), h1_node=subgraph[0][1].astype(np.int32), h0_eidx=subgraph[1][0].astype(np.int32),
               h1_eidx=subgraph[1][1].astype(np.int32), w0=w0.numpy(), w1=w1.numpy(), score=score.detach().numpy(), imp0=imp0.detach().numpy(),
               imp1=imp1.detach().numpy(), kl=np.float64(kl.item()), loss=np.float64(loss.item()), null_values=np.array(list(null.values()), np.float64),
               hid_dim=int(hid), use_temporal=int(use_temporal), prior=np.array(prior))
    np.savez_compressed(os.path.join(HERE, f"trainstep_{tag}.npz"), **out)


def gen_train_all():
    big = dict(np.load(os.path.join(HERE, "rand_bigts.npz")))
    Bq = 12
    wn, we, wt, wa = (big[f"tgt_w_{k}"][:Bq] for k in ("nodes", "eidx", "t", "anony"))
    allw = np.concatenate([x.astype(np.float64) for x in (wn, we, wt, wa)], axis=-1)
    new = marginal(allw, allw, allw)[0]
    walks5 = (wn.astype(np.int64), we.astype(np.int64), wt.astype(np.float64), new[:, :, 12:13].astype(np.int64), new[:, :, 13:14])
    ei = new_edge_info(we.astype(int))
    sub = ([big["tgt_hop0_node"][:Bq].astype(np.int64), big["tgt_hop1_node"][:Bq].astype(np.int64)],
           [big["tgt_hop0_eidx"][:Bq].astype(np.int64), big["tgt_hop1_eidx"][:Bq].astype(np.int64)], None)
    args = (walks5, ei, big["ts"][big["q"][:Bq]], sub, int(big["n_nodes"]), len(big["eidx"]) + 1)
    gen_train_step("d32", *args, 32, 32, seed=21)
    gen_train_step("d32_hid32_uniform", *args, 32, 32, seed=22, hid=32, prior="uniform")
    us = dict(np.load(os.path.join(HERE, "uslegis.npz")))
    B = 4
    src, dst, eidx, ts = load_uslegis()
    walks5 = (us["src_w_nodes"][:B].astype(np.int64), us["src_w_eidx"][:B].astype(np.int64), us["src_w_t"][:B].astype(np.float64),
              us["src_cat"][:B, :, None].astype(np.int64), us["src_marginal"][:B, :, None])
    sub = ([us["src_hop0_node"][:B].astype(np.int64), us["src_hop1_node"][:B].astype(np.int64)],
           [us["src_hop0_eidx"][:B].astype(np.int64), us["src_hop1_eidx"][:B].astype(np.int64)], None)
    gen_train_step("d172", walks5, us["src_edge_identity"][:B].astype(np.float64), ts[us["q"][:B]].astype(np.float64), sub,
                   int(us["n_nodes"]), len(eidx) + 1, 172, 1, seed=23)


def gen_next_step_time():
    """get_next_step with e_idx_l = None (utils/graph.py:308-333, the bisect branch of find_before_walk) on the rand_small graph, run on
    the unmodified reference under the draw contract (stage 16, row = i)."""
    z = dict(np.load(os.path.join(HERE, "rand_small.npz")))
    n_nodes = int(z["n_nodes"])
    nf = rg.NeighborFinder(refshim.adj_list_from_events(n_nodes, z["src"], z["dst"], z["eidx"], z["ts"]))
    roots = z["src"][z["q"]].astype(np.int64)
    degree, N2 = int(z["n"]), 3
    nbr = z["src_hop0_node"].reshape(-1).astype(np.int64)            # first-hop neighbours [B * degree] and their (float32) times
    cut = z["src_hop0_ts"].reshape(-1).astype(np.float32)
    shim = refshim.DrawShim(base_seed=31)
    with shim.patched(rg):
        shim.seed = 31
        s2, t2, e2, ts2 = nf.get_next_step(nbr, cut, N2, degree, e_idx_l=None, source_id=roots)
    np.savez_compressed(os.path.join(HERE, "nextstep_time.npz"), roots=roots, nbr=nbr, cut=cut, degree=degree, N2=N2, seed=31,
                        o_src=s2, o_tgt=t2, o_eidx=e2, o_ts=ts2)


def gen_edge_imp_all():
    """Fixtures of the motif -> edge aggregation; reads the committed walk fixtures (does not regenerate them)."""
    us = dict(np.load(os.path.join(HERE, "uslegis.npz")))
    big = dict(np.load(os.path.join(HERE, "rand_bigts.npz")))
    B = 6
    src, dst, eidx, ts = load_uslegis()
    walks5 = (us["src_w_nodes"][:B].astype(np.int64), us["src_w_eidx"][:B].astype(np.int64), us["src_w_t"][:B].astype(np.float64),
              us["src_cat"][:B, :, None].astype(np.int64), us["src_marginal"][:B, :, None])
    sub = ([us["src_hop0_node"][:B].astype(np.int64), us["src_hop1_node"][:B].astype(np.int64)],
           [us["src_hop0_eidx"][:B].astype(np.int64), us["src_hop1_eidx"][:B].astype(np.int64)], None)
    gen_edge_imp("d172", walks5, us["src_edge_identity"][:B].astype(np.float64), ts[us["q"][:B]].astype(np.float64), sub,
                 int(us["n_nodes"]), len(eidx) + 1, 172, 1, seed=3)
    Bq = 12
    wn, we, wt, wa = (big[f"tgt_w_{k}"][:Bq] for k in ("nodes", "eidx", "t", "anony"))
    allw = np.concatenate([x.astype(np.float64) for x in (wn, we, wt, wa)], axis=-1)
    new = marginal(allw, allw, allw)[0]
    walks5 = (wn.astype(np.int64), we.astype(np.int64), wt.astype(np.float64), new[:, :, 12:13].astype(np.int64), new[:, :, 13:14])
    ei = new_edge_info(we.astype(int))
    sub = ([big["tgt_hop0_node"][:Bq].astype(np.int64), big["tgt_hop1_node"][:Bq].astype(np.int64)],
           [big["tgt_hop0_eidx"][:Bq].astype(np.int64), big["tgt_hop1_eidx"][:Bq].astype(np.int64)], None)
    gen_edge_imp("d32", walks5, ei, big["ts"][big["q"][:Bq]], sub, int(big["n_nodes"]), len(big["eidx"]) + 1, 32, 32, seed=4)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "edgeimp":
        gen_edge_imp_all()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "nocat":
        gen_encoder_nocat()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "kl":
        gen_kl_all()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "nextstep":
        gen_next_step_time()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        gen_train_all()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "enhance":
        gen_enhance_all()
        return
    gen_tie_star()
    gen_rand_small()
    gen_native_rng()
    big = gen_rand_bigts()
    us = gen_uslegis()
    gen_nullmodel()
    # encoder on uslegis-shaped walks: D = 172, Ed = 1 (SURVEY cfg 1), 6 roots x 90 walks
    B = 6
    src, dst, eidx, ts = load_uslegis()
    walks5 = (us["src_w_nodes"][:B].astype(np.int64), us["src_w_eidx"][:B].astype(np.int64), us["src_w_t"][:B].astype(np.float64),
              us["src_cat"][:B, :, None].astype(np.int64), us["src_marginal"][:B, :, None])
    gen_encoder("d172", walks5, us["src_edge_identity"][:B].astype(np.float64), ts[us["q"][:B]].astype(np.float64),
                us["n_nodes"], len(eidx) + 1, 172, 1, seed=0)
    # encoder with 32-d features and ~1e8 timestamps (cfg 2 shape): float32 rounding of ts and big cos arguments
    Bq = 12
    wn, we, wt, wa = (big[f"tgt_w_{k}"][:Bq] for k in ("nodes", "eidx", "t", "anony"))
    allw = np.concatenate([x.astype(np.float64) for x in (wn, we, wt, wa)], axis=-1)
    new = marginal(allw, allw, allw)[0]
    walks5 = (wn.astype(np.int64), we.astype(np.int64), wt.astype(np.float64), new[:, :, 12:13].astype(np.int64), new[:, :, 13:14])
    ei = new_edge_info(we.astype(int))
    gen_encoder("d32", walks5, ei, big["ts"][big["q"][:Bq]], big["n_nodes"], len(big["eidx"]) + 1, 32, 32, seed=1)
    gen_encoder("d32_plainattn", walks5, ei, big["ts"][big["q"][:Bq]], big["n_nodes"], len(big["eidx"]) + 1, 32, 32, seed=2,
                use_temporal=False)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()

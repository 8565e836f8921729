"""The drop-in boundary exercised the way the reference's own callers use it (-m gpu): the sys.modules injection of
INTEGRATION.md 2, then the exact call shapes of TGN/tgn.py:280-285 and GraphM/graphmixer.py:224-234 (set_neighbor_sampler /
grab_subgraph), temp_exp_main.py:135-144 (finder from an adj_list), models/explainer.py:8,132 (get_null_distribution inside the
explainer's constructor), processed/data_preprocess.py (pack on disk) and temp_exp_main.py:300-330 / 590-632 (get_item ->
Explainer(...) -> retrieve_explanation(training=args.if_bern) -> kl_loss -> backward).  /root/reference does not exist on the GPU
box, so the callers are restated here line for line as small classes; results are checked against the CPU oracle."""
import sys
import time
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tempme_b200
    return tempme_b200


class TGNLike:                      # TGN/tgn.py:280-285
    def __init__(self, num_neighbors):
        self.num_neighbors = num_neighbors
        self.embedding_module = types.SimpleNamespace(neighbor_sampler=None)

    def set_neighbor_sampler(self, neighbor_finder):
        self.embedding_module.neighbor_sampler = neighbor_finder

    def grab_subgraph(self, src_idx_l, cut_time_l):
        return self.embedding_module.neighbor_sampler.find_k_hop(2, src_idx_l, cut_time_l, num_neighbors=self.num_neighbors, e_idx_l=None)


class GraphMixerLike:               # GraphM/graphmixer.py:224-234
    def __init__(self, num_neighbors):
        self.num_neighbors = num_neighbors

    def set_neighbor_sampler(self, neighbor_sampler):
        self.neighbor_sampler = neighbor_sampler

    def grab_subgraph(self, src_idx_l, cut_time_l):
        return self.neighbor_sampler.find_k_hop(2, src_idx_l, cut_time_l, num_neighbors=self.num_neighbors, e_idx_l=None)


def get_item(input_pack, batch_id):                     # utils/batch_loader.py:204-234
    *subs, walks_src, walks_tgt, walks_bgd, dst_fake = input_pack
    subs = [tuple([i[batch_id] for i in rec] for rec in s) for s in subs]
    walks = [tuple(item[batch_id] for item in w) for w in (walks_src, walks_tgt, walks_bgd)]
    return (*subs, *walks, dst_fake[batch_id])


def test_reference_callers_on_the_drop_in(tm, golden, tmp_path, monkeypatch):
    import oracle
    from tempme_b200 import compat
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.graph", "utils.null_model", "processed", "processed.utils", "models")}
    try:
        compat.install(also_models=True)
        import utils                                    # noqa: F401  (the injected module)
        from utils import NeighborFinder, RandEdgeSampler, get_null_distribution          # models/explainer.py:8, temp_exp_main.py
        from processed.utils import NeighborFinder as NF2                                  # processed/data_preprocess.py:16
        from models import TempME
        assert NeighborFinder is tm.NeighborFinder and NF2 is NeighborFinder

        g = golden("uslegis")
        src, dst, eidx, ts = g["src"].astype(np.int64), g["dst"].astype(np.int64), g["eidx"].astype(np.int64), g["ts"].astype(np.float64)
        n_nodes = int(g["n_nodes"])
        # temp_exp_main.py:135-144: adjacency lists of (neighbour, e_idx, ts) tuples, every event appended to both endpoints
        adj_list = [[] for _ in range(n_nodes)]
        for s, d, e, t in zip(src, dst, eidx, ts):
            adj_list[s].append((d, e, t)); adj_list[d].append((s, e, t))
        finder = NeighborFinder(adj_list)
        og = oracle.OracleGraph.from_events(n_nodes, src, dst, eidx, ts)
        q = np.nonzero(ts > np.quantile(ts, 0.85))[0][:200]
        for base in (TGNLike(30), GraphMixerLike(30)):
            base.set_neighbor_sampler(finder)
            finder.seed, finder.calls = 77, 0           # the draw contract: top-level call c of a finder uses seed + c
            sub = base.grab_subgraph(src[q], ts[q])
            ref = og.find_k_hop(2, src[q], ts[q], 30, None, seed=77)
            assert isinstance(sub, tuple) and len(sub) == 3 and [len(r) for r in sub] == [2, 2, 2]
            for a, b in zip(sub, ref):
                for x, y in zip(a, b):
                    assert isinstance(x, np.ndarray) and x.dtype == y.dtype and x.shape == y.shape and (x == y).all()
        # scalar find_before, the per-event lookups of TGAT-style callers: views, reference dtypes, IndexError
        t0 = time.perf_counter()
        for i in q[:100]:
            nb, ee, tt, _ = finder.find_before(int(src[i]), float(ts[i]), e_idx=int(eidx[i]))
            s_o, c_o = og.find_before_batch(src[i:i + 1], None, eidx[i:i + 1])
            assert len(nb) == int(c_o[0]) and nb.dtype == np.int64 and tt.dtype == np.float64 and (np.diff(tt) >= 0).all()
        per_call = (time.perf_counter() - t0) / 100
        assert per_call < 2e-3, f"scalar find_before takes {per_call * 1e6:.0f} us"
        with pytest.raises(IndexError):
            finder.find_before(int(src[q[0]]), float(ts[q[0]]), e_idx=10 ** 6)

        # the offline pack (processed/data_preprocess.py:99-145,393-419) on disk, read back the way the drivers do
        rs = RandEdgeSampler((src,), (dst,))
        np.random.seed(0)
        fake = rs.sample(len(q))[1]
        pack, edge = tm.build_pack(finder, src[q], dst[q], ts[q], eidx[q], fake, 30, 3, seed=5)
        cat_path, edge_path = tm.save_pack(pack, edge, str(tmp_path), "uslegis_sampled", "test")
        args = types.SimpleNamespace(n_degree=30, if_bern=True, prior_p=0.3, beta=0.5)
        test_pack = utils.load_subgraph_margin(args, tm.load_pack(cat_path))               # temp_exp_main.py:705-714
        test_edge = np.load(edge_path)
        assert test_edge.shape == (3, len(q), 90, 3, 3)

        # models/explainer.py:132: the constructor computes the null model from processed/ml_<data>.csv
        import pandas as pd
        pd.DataFrame({"u": src, "i": dst, "ts": ts, "label": np.zeros(len(src)), "idx": eidx}).to_csv(tmp_path / "ml_uslegis_sampled.csv")
        monkeypatch.setenv("TEMPME_DATA_ROOT", str(tmp_path))
        null = get_null_distribution(data_name="uslegis_sampled")
        assert sorted(null) == list(range(1, 13)) and abs(sum(null.values()) - 1.0) < 1e-9

        dev = torch.device("cuda:0")
        torch.manual_seed(0)
        nfeat = torch.randn(n_nodes, 172); efeat = torch.randn(len(src) + 1, 4); nfeat[0] = 0; efeat[0] = 0
        base = types.SimpleNamespace(n_feat_th=nfeat.to(dev), e_feat_th=efeat.to(dev),
                                     node_raw_features=torch.nn.Embedding.from_pretrained(nfeat.to(dev), padding_idx=0, freeze=True),
                                     edge_raw_features=torch.nn.Embedding.from_pretrained(efeat.to(dev), padding_idx=0, freeze=True))
        Explainer = TempME(base, base_model_type="tgn", data="uslegis_sampled", out_dim=40, hid_dim=64, temp=0.07, if_cat_feature=True,
                           dropout_p=0.1, device=dev)                                      # temp_exp_main.py:550-553
        Explainer = Explainer.to(dev)
        optimizer = torch.optim.Adam(Explainer.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
        batch_idx = np.arange(0, 100)
        ts_l_cut = ts[q][batch_idx]
        subgraph_src, subgraph_tgt, subgraph_bgd, walks_src, walks_tgt, walks_bgd, dst_l_fake = get_item(test_pack, batch_idx)
        src_edge, tgt_edge, bgd_edge = (test_edge[:, batch_idx, :, :, :][k] for k in range(3))    # get_item_edge
        # ---- evaluation loop (temp_exp_main.py:300-330): eval(), gradients enabled, training=args.if_bern
        Explainer.eval()
        imps = [Explainer(w, ts_l_cut, e) for w, e in ((walks_src, src_edge), (walks_tgt, tgt_edge), (walks_bgd, bgd_edge))]
        assert all(x.shape == (100, 90, 1) and not x.requires_grad for x in imps)
        explanation = Explainer.retrieve_explanation(subgraph_src, imps[0], walks_src, subgraph_tgt, imps[1], walks_tgt,
                                                     subgraph_bgd, imps[2], walks_bgd, training=args.if_bern)
        assert [tuple(x.shape) for x in explanation] == [(300, 30), (300, 900)]
        kl = sum(Explainer.kl_loss(x, w, target=args.prior_p) for x, w in zip(imps, (walks_src, walks_tgt, walks_bgd)))
        assert np.isfinite(float(kl))
        # ---- training step (temp_exp_main.py:590-632)
        Explainer.train()
        optimizer.zero_grad()
        imps = [Explainer(w, ts_l_cut, e) for w, e in ((walks_src, src_edge), (walks_tgt, tgt_edge), (walks_bgd, bgd_edge))]
        explanation = Explainer.retrieve_explanation(subgraph_src, imps[0], walks_src, subgraph_tgt, imps[1], walks_tgt,
                                                     subgraph_bgd, imps[2], walks_bgd, training=args.if_bern)
        pred = torch.cat([explanation[0].mean(1, keepdim=True)[:100], explanation[1].mean(1, keepdim=True)[200:]], dim=0)   # stands in for base_model.contrast
        y_ori = torch.cat([torch.ones(100, 1), torch.zeros(100, 1)]).to(dev)
        pred_loss = torch.nn.BCEWithLogitsLoss()(pred, y_ori)
        kl_loss = sum(Explainer.kl_loss(x, w, target=args.prior_p) for x, w in zip(imps, (walks_src, walks_tgt, walks_bgd)))
        loss = pred_loss + args.beta * kl_loss
        before = Explainer.MLP[5].weight.detach().clone()
        loss.backward()
        optimizer.step()
        assert torch.isfinite(loss) and not torch.equal(before, Explainer.MLP[5].weight)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

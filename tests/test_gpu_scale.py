"""GPU parity AT SCALE (-m gpu): the CUDA path through the C ABI against the CPU oracle on the graph shapes the bench
runs -- cfg3 and cfg4 at full size, cfg5 at 5 % (5 M events: hub windows of ~10^5 entries, deep 33-ary searches, long
runs of the secondary index, a multi-million-slot run directory) -- with the top hubs among the roots, and a time cut
on a >10^6-entry window.  Integer outputs bit-exact, scores rtol 1e-5.  Nothing here reads /root/reference."""
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


@pytest.fixture(scope="module")
def orc():
    import oracle
    return oracle


class _Base:
    def __init__(self, nfeat, efeat):
        self.n_feat_th = nfeat.cuda(); self.e_feat_th = efeat.cuda()
        self.node_raw_features = torch.nn.Embedding.from_pretrained(self.n_feat_th, padding_idx=0, freeze=True)
        self.edge_raw_features = torch.nn.Embedding.from_pretrained(self.e_feat_th, padding_idx=0, freeze=True)


def hub_queries(g, rng, count, n_hubs=10):
    """Query events of the test split, `count` of them, forced to include the latest events of the `n_hubs` largest nodes
    (so that hub windows are searched, sampled and id-filtered) and hub ids among the background roots."""
    from tempme_b200 import synth
    src, dst, fake, ts, eidx = [np.array(a) for a in synth.make_queries(g, rng, count)]
    deg = np.bincount(g["src"], minlength=g["n_nodes"]) + np.bincount(g["dst"], minlength=g["n_nodes"])
    hubs = np.argsort(-deg)[:n_hubs]
    tail = slice(int(0.9 * len(g["src"])), None)
    s_t, d_t = np.asarray(g["src"][tail]), np.asarray(g["dst"][tail])
    picks = []
    for h in hubs:
        hit = np.nonzero((s_t == h) | (d_t == h))[0]
        if len(hit):
            picks.append(tail.start + hit[-1]); picks.append(tail.start + hit[len(hit) // 2])
    picks = np.array(sorted(set(picks)), dtype=np.int64)
    k = len(picks)
    src[:k], dst[:k], ts[:k], eidx[:k] = g["src"][picks], g["dst"][picks], g["ts"][picks], g["eidx"][picks]
    fake[:len(hubs)] = hubs                       # hub as a background root: time cut over its whole window
    order = np.argsort(ts, kind="stable")
    return src[order], dst[order], fake[order], ts[order], eidx[order], hubs, deg


@pytest.mark.parametrize("cfg,scale,count", [("cfg5", 0.05, 300), ("cfg4", 1.0, 300), ("cfg3", 1.0, 200)])
def test_bench_graph_shapes_against_oracle(tm, orc, cfg, scale, count):
    from oracle import encoder as orc_enc
    from tempme_b200 import synth
    sh = synth.SHAPES[cfg]
    n, N2, D, Ed = sh["n"], sh["N2"], sh["D"], sh["Ed"]
    g = synth.make_graph(cfg, scale)
    N, E = g["n_nodes"], len(g["src"])
    f = tm.NeighborFinder.from_events(N, g["src"], g["dst"], g["eidx"], g["ts"])
    og = orc.OracleGraph.from_events(N, g["src"], g["dst"], g["eidx"], g["ts"])
    off, nbr, e, t = og.export()
    assert (f.off_set_l == off).all() and (f.node_idx_l == nbr).all() and (f.edge_idx_l == e).all() and (f.node_ts_l == t).all()
    src, dst, fake, ts, eidx, hubs, deg = hub_queries(g, np.random.default_rng(77), count)
    if cfg == "cfg5":
        assert deg[hubs[0]] >= 50_000               # the hub windows this test is about
    # find_before, both cut modes, incl. the hubs' whole windows
    for roots, ct, ee in ((src, None, eidx), (dst, None, eidx), (fake, ts, None)):
        s_o, c_o = og.find_before_batch(roots, ct, ee)
        s_d, c_d = f.find_before_batch_device(roots, ct, ee)
        assert (s_d.cpu().numpy() == s_o).all() and (c_d.cpu().numpy() == c_o).all()
    f.check_errors()
    nfeat, efeat = synth.make_features(cfg, N, E)
    torch.manual_seed(5)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "unit", 40, 64, device="cuda", null_model={}).cuda().eval()
    p = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    nf, ef = nfeat.numpy(), efeat.numpy()
    for roots, ee, sd in ((src, eidx, 21), (dst, eidx, 22), (fake, None, 23)):
        sub = f.find_k_hop(2, roots, ts, n, ee, seed=sd, row_offset=5)
        osub = og.find_k_hop(2, roots, ts, n, ee, seed=sd, row_offset=5)
        for a, b in zip(sub, osub):
            for x, y in zip(a, b):
                assert x.dtype == y.dtype and (x == y).all()
        scanned = torch.zeros(1, dtype=torch.int64, device="cuda")
        nodes, we, wt, anony, cat = f.find_k_walks_device(n, roots, N2, sub, seed=sd + 50, row_offset=5, scanned=scanned)
        on, oe, ot, oa, osc = og.sample_walks(roots, osub[0][0], osub[1][0], osub[2][0], N2, seed=sd + 50, row_offset=5, want_scanned=True)
        assert (nodes.cpu().numpy() == on).all() and (we.cpu().numpy() == oe).all()
        assert (wt.cpu().numpy() == ot).all() and (anony.cpu().numpy() == oa).all()
        ocat = orc.class_ids_prep(oa)[0]
        assert (cat.cpu().numpy() == ocat).all()
        assert int(scanned.item()) == int(osc.sum())
        eid = tm.edge_identity_device(we)
        oeid = orc.edge_identity(oe)
        assert (eid.cpu().numpy() == oeid).all()
        cut = torch.as_tensor(ts.astype(np.float32)).cuda()
        scores = m.score_device(nodes, we, wt, cat, cut, eid, group=100).cpu().numpy()
        for s in range(0, len(roots), 100):
            sl = slice(s, s + 100)
            ref = orc_enc.forward(p, nf, ef, (on[sl], oe[sl], ot[sl], ocat[sl], None), ts[sl], oeid[sl])
            np.testing.assert_allclose(scores[sl], ref[..., 0], rtol=1e-5, atol=0)
    f.check_errors()


def test_time_cut_on_a_million_entry_window(tm, orc):
    """find_before by time (bisect_left_adapt, utils/graph.py:511-530) on one node with 1.25 M entries: ties, exact hits, both ends."""
    rng = np.random.default_rng(3)
    E = 1_250_000
    hub, leaves = 1, rng.integers(2, 5000, E)
    ts = np.sort(rng.integers(0, 400_000, E)).astype(np.float64) * 0.5          # ~3 events per distinct timestamp
    src = np.full(E, hub, np.int64)
    f = tm.NeighborFinder.from_events(5000, src, leaves, np.arange(1, E + 1), ts)
    cuts = np.concatenate([ts[rng.integers(0, E, 2000)], ts[rng.integers(0, E, 2000)] + 0.25, [-1.0, 0.0, ts[0], ts[-1], ts[-1] + 1, 1e300]])
    start, cut = f.find_before_batch_device(np.full(len(cuts), hub), cuts, None)
    win = f.node_ts_l[f.off_set_l[hub]:f.off_set_l[hub + 1]]
    assert len(win) == E and (np.diff(win) >= 0).all()
    assert (cut.cpu().numpy() == np.searchsorted(win, cuts, side="left")).all()        # strict lower bound: events at the cut time excluded
    assert (start.cpu().numpy() == f.off_set_l[hub]).all()
    og = orc.OracleGraph.from_events(5000, src, leaves, np.arange(1, E + 1), ts)
    s_o, c_o = og.find_before_batch(np.full(len(cuts), hub), cuts, None)
    assert (cut.cpu().numpy() == c_o).all()
    # e_idx cuts on the same window (tie groups collapse to the group's first slot unless trailing, App. A.2)
    qe = rng.integers(1, E + 1, 3000)
    s_d, c_d = f.find_before_batch_device(np.full(len(qe), hub), None, qe)
    s_o, c_o = og.find_before_batch(np.full(len(qe), hub), None, qe)
    assert (c_d.cpu().numpy() == c_o).all()
    # sampling from the full window: rows sorted by position, every entry precedes the cut
    sub = f.find_k_hop(1, np.full(64, hub), np.full(64, 1e300), 30, None, seed=9)
    osub = og.find_k_hop(1, np.full(64, hub), np.full(64, 1e300), 30, None, seed=9)
    for a, b in zip(sub, osub):
        assert (a[0] == b[0]).all()


def _rand_events(seed, N, E, tmax, lo=1, loops=0):
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, N - lo + 1) ** 1.1
    p /= p.sum()
    src = rng.choice(np.arange(lo, N), E, p=p); dst = rng.choice(np.arange(lo, N), E, p=p)
    if loops:
        k = rng.choice(E, loops, replace=False)
        dst[k] = src[k]
    ts = np.sort(rng.integers(0, tmax, E)).astype(np.float64)
    return src, dst, np.arange(1, E + 1), ts


@pytest.mark.parametrize("case", ["ties", "no_ties", "self_loops", "unsorted_input", "node0", "cfg5_5pct"])
def test_device_build_equals_host_build(tm, monkeypatch, case):
    """K1 on the device (radix sorts, scan, per-edge table, secondary index) == the literal host pass, array for array."""
    from tempme_b200 import synth
    if case == "cfg5_5pct":
        g = synth.make_graph("cfg5", 0.05)
        N, (src, dst, eidx, ts) = g["n_nodes"], (g["src"], g["dst"], g["eidx"], g["ts"])
    else:
        N = 400
        src, dst, eidx, ts = _rand_events(11, N, 60000, {"ties": 300, "no_ties": 10 ** 9}.get(case, 5000), lo=0 if case == "node0" else 1,
                                          loops=50 if case == "self_loops" else 0)
        if case == "unsorted_input":                        # the API does not require chronological input: the sort must do the work
            perm = np.random.default_rng(5).permutation(len(src))
            src, dst, eidx, ts = src[perm], dst[perm], eidx[perm], ts[perm]
    monkeypatch.delenv("TEMPME_GRAPH_BUILD", raising=False)
    dev = tm.NeighborFinder.from_events(N, src, dst, eidx, ts)
    monkeypatch.setenv("TEMPME_GRAPH_BUILD", "host")
    host = tm.NeighborFinder.from_events(N, src, dst, eidx, ts)
    monkeypatch.delenv("TEMPME_GRAPH_BUILD")
    assert (dev.off_set_l == host.off_set_l).all() and (dev.node_idx_l == host.node_idx_l).all()
    assert (dev.edge_idx_l == host.edge_idx_l).all() and (dev.node_ts_l == host.node_ts_l).all()
    assert (dev.edge_table() == host.edge_table()).all()
    assert (dev.secondary_index() == host.secondary_index()).all()
    # the adj_list entry point takes the same route
    if case in ("ties", "self_loops"):
        node = np.empty(2 * len(src), np.int32); nbr = np.empty_like(node)
        node[0::2] = src; node[1::2] = dst; nbr[0::2] = dst; nbr[1::2] = src
        ent = tm.NeighborFinder(None, _entries=(N, (node, nbr, np.repeat(eidx.astype(np.int32), 2), np.repeat(ts, 2))))
        assert (ent.off_set_l == host.off_set_l).all() and (ent.edge_table() == host.edge_table()).all() and (ent.secondary_index() == host.secondary_index()).all()

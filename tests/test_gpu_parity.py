"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI via
tempme_b200's reference-facing classes, against (a) the committed golden vectors produced by the
unmodified reference and (b) the CPU oracle on seeded inputs.  Integer / index outputs and copied
float32 timestamps are compared bit-exactly; fp32 scores within rtol 1e-5 (the north_star tolerance).
Nothing here reads /root/reference."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOTS = ("src", "tgt", "bgd")


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


def finder_of(tm, g, upto=None, seed=0):
    s = slice(None, upto)
    return tm.NeighborFinder.from_events(int(g["n_nodes"]), g["src"][s].astype(np.int64), g["dst"][s].astype(np.int64),
                                         g["eidx"][s], g["ts"][s].astype(np.float64), seed=seed)


def assert_sub(g, prefix, sub):
    for name, rec in zip(("node", "eidx", "ts"), sub):
        for l, a in enumerate(rec):
            ref = g[f"{prefix}_hop{l}_{name}"]
            assert a.shape == ref.shape and a.dtype == (np.float32 if name == "ts" else np.int32)
            assert (a == ref).all(), f"{prefix} hop{l} {name}"


def assert_walks(g, prefix, walks):
    for name, a in zip(("nodes", "eidx", "t", "anony"), walks):
        ref = g[f"{prefix}_w_{name}"]
        assert a.shape == ref.shape, (a.shape, ref.shape)
        assert (a == ref).all(), f"{prefix} walks {name}"


# ------------------------------------------------------------------------------------------ graph
def test_tie_star_csr_and_cuts(tm, golden):
    g = golden("tie_star")
    f = finder_of(tm, g)
    assert (f.off_set_l == g["off"]).all() and (f.node_idx_l == g["nbr"]).all()
    assert (f.edge_idx_l == g["e"]).all() and (f.node_ts_l == g["t"]).all()
    d = f.nodeedge2idx
    for v, e, val in g["dict"]:
        assert d[int(v)][int(e)] == val
    assert [d[1][k] for k in range(1, 7)] == [0, 1, 1, 1, 4, 5]
    for v, e, n in g["fb_eidx"]:
        assert len(f.find_before(int(v), 3.0, e_idx=int(e))[0]) == n
    for v, t, n in g["fb_time"]:
        nb, ee, ts, _ = f.find_before(int(v), float(t))
        assert len(nb) == int(n) and (ts < t).all()
    with pytest.raises(IndexError):
        f.find_before(1, 3.0, e_idx=99)
    assert len(f.find_before(0, 3.0, e_idx=1)[0]) == 0                 # node 0 -> cut 0 (graph.py:133)


@pytest.mark.parametrize("name", ["rand_small", "rand_bigts", "uslegis"])
def test_golden_csr_hops_walks(tm, orc, golden, name):
    g = golden(name)
    f = finder_of(tm, g)
    if "off" in g:
        assert (f.off_set_l == g["off"]).all() and (f.node_idx_l == g["nbr"]).all()
        assert (f.edge_idx_l == g["e"]).all() and (f.node_ts_l == g["t"]).all()
        d = f.nodeedge2idx
        for v, e, val in g["dict"]:
            L = g["off"][v + 1] - g["off"][v]
            eff = max(0, L + val) if val < 0 else min(val, L)
            assert d[int(v)][int(e)] == eff
    q, n, N2, seed = g["q"], int(g["n"]), int(g["N2"]), int(g["base_seed"])
    ts = g["ts"].astype(np.float64)
    call = 0
    for r in ROOTS:
        roots = {"src": g["src"][q], "tgt": g["dst"][q], "bgd": g["fake"]}[r].astype(np.int64)
        e = None if r == "bgd" else g["eidx"][q]
        sub = f.find_k_hop(2, roots, ts[q], n, e, seed=seed + call)
        assert_sub(g, r, sub)
        walks = f.find_k_walks(n, roots, N2, sub, seed=seed + call + 1)
        assert walks[0].dtype == np.int64
        assert_walks(g, r, walks)
        call += 2
        if f"{r}_edge_identity" in g:
            ei = tm.new_edge_info(walks[1])
            assert ei.dtype == np.float64 and (ei == g[f"{r}_edge_identity"]).all()
        if f"{r}_cat" in g:
            an = torch.as_tensor(walks[3]).cuda()
            hn, hp, cat, err = tm.class_hist_device(an)
            assert int(err.item()) == 0 and (cat.cpu().numpy() == g[f"{r}_cat"]).all()
            assert (hn.cpu().numpy() == orc.class_hist_null(walks[3])).all()
            assert (hp.cpu().numpy() == orc.class_ids_prep(walks[3])[1]).all()


def test_shard_offset_wide_fanout_and_call_counter(tm, golden):
    g = golden("rand_small")
    f = finder_of(tm, g, seed=77)
    q, n, N2 = g["q"], int(g["n"]), int(g["N2"])
    ts = g["ts"]
    sub = f.find_k_hop(2, g["src"][q][24:], ts[q][24:], n, g["eidx"][q][24:], seed=77, row_offset=24)
    assert_sub(g, "shard1", sub)
    walks = f.find_k_walks(n, g["src"][q][24:], N2, sub, seed=78, row_offset=24)
    assert_walks(g, "shard1", walks)
    assert (sub[0][1] == g["src_hop1_node"][24:]).all() and (walks[0] == g["src_w_nodes"][24:]).all()
    sub = f.find_k_hop(1, g["dst"][q], ts[q], 40, g["eidx"][q], seed=5)
    assert_sub(g, "wide", sub)
    assert_walks(g, "wide", f.find_k_walks(40, g["dst"][q], 1, sub, seed=6))
    # implicit seeds: call number c uses seed + c, exactly how the golden run numbered its calls
    f2 = finder_of(tm, g, seed=77)
    sub = f2.find_k_hop(2, g["src"][q], ts[q], n, g["eidx"][q])
    assert_sub(g, "src", sub)
    assert_walks(g, "src", f2.find_k_walks(n, g["src"][q], N2, sub))


def test_train_finder_missing_eidx(tm, golden):
    g = golden("rand_bigts")
    f = finder_of(tm, g, int(g["n_train"]))
    q, n, N2 = g["q"], int(g["n"]), int(g["N2"])
    with pytest.raises(IndexError):
        f.find_k_hop(1, g["src"][q], g["ts"][q], n, g["eidx"][q])
    with pytest.raises(IndexError):
        f.find_before(int(g["src"][q][0]), float(g["ts"][q][0]), e_idx=int(g["eidx"][q][0]))
    sub = f.find_k_hop(2, g["src"][q], g["ts"][q], n, None, seed=9 + 6)
    assert_sub(g, "train", sub)
    assert_walks(g, "train", f.find_k_walks(n, g["src"][q], N2, sub, seed=9 + 7))


def test_null_model_distribution_golden(tm, golden):
    g = golden("nullmodel")
    f = finder_of(tm, g, seed=int(g["base_seed"]))
    ti = g["test_idx"]
    dist = tm.pre_processing(f, None, g["src"][ti].astype(np.int64), g["dst"][ti].astype(np.int64),
                             g["ts"][ti].astype(np.float64), g["eidx"][ti], int(g["n"]), fakes=g["fakes"])
    assert list(dist.keys()) == list(range(1, 13))
    assert np.array_equal(np.array([dist[k] for k in range(1, 13)]), g["dist"])


def test_empty_and_edge_inputs(tm, golden):
    g = golden("rand_small")
    f = finder_of(tm, g)
    e = np.zeros(0, np.int64)
    sub = f.find_k_hop(2, e, e.astype(np.float64), 5, e)
    assert sub[0][0].shape == (0, 5) and sub[0][1].shape == (0, 25)
    w = f.find_k_walks(5, e, 3, sub)
    assert w[0].shape == (0, 15, 6) and w[3].shape == (0, 15, 3)
    # a root with no history gives an all-padding row and [1, 3, 0] motifs (graph.py:214-215, 395-435)
    sub = f.find_k_hop(1, np.array([1]), np.array([-5.0]), 4, None, seed=1)
    assert (sub[0][0] == 0).all() and (sub[2][0] == 0).all()
    w = f.find_k_walks(4, np.array([1]), 2, sub, seed=2)
    assert (w[0][..., :4] == 0).all() and (w[3] == np.array([1, 3, 0])).all()
    # graph without any event
    f0 = tm.NeighborFinder([[] for _ in range(4)])
    sub = f0.find_k_hop(1, np.array([1, 2]), np.array([1.0, 2.0]), 3, None)
    assert (sub[0][0] == 0).all()
    with pytest.raises(NotImplementedError):
        tm.NeighborFinder([[]], bias=1.0)
    with pytest.raises(NotImplementedError):
        f.find_k_walks(4, np.array([1]), 33, f.find_k_hop(1, np.array([1]), np.array([5.0]), 4, None))


# ------------------------------------------------------------------------------------------ vs oracle, seeded
def synth_graph(seed, N, E, tmax, zipf=1.0, lo=1):
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, N - lo + 1) ** zipf
    p /= p.sum()
    src = rng.choice(np.arange(lo, N), E, p=p); dst = rng.choice(np.arange(lo, N), E, p=p)
    ts = np.sort(rng.integers(0, tmax, E)).astype(np.float64)
    return src, dst, np.arange(1, E + 1), ts


@pytest.mark.parametrize("seed,N,E,tmax,n,N2", [(1, 185, 60000, 10 ** 7, 30, 1), (2, 3000, 80000, 50000, 20, 3), (3, 500, 30000, 40, 20, 5)])
def test_oracle_parity_seeded(tm, orc, seed, N, E, tmax, n, N2):
    src, dst, eidx, ts = synth_graph(seed, N, E, tmax)
    f = tm.NeighborFinder.from_events(N, src, dst, eidx, ts)
    og = orc.OracleGraph.from_events(N, src, dst, eidx, ts)
    off, nbr, e, t = og.export()
    assert (f.off_set_l == off).all() and (f.node_idx_l == nbr).all() and (f.edge_idx_l == e).all() and (f.node_ts_l == t).all()
    rng = np.random.default_rng(seed + 100)
    q = rng.choice(np.arange(E // 2, E), 600, replace=False)
    # find_before, both modes, straight through the batch call
    fake = rng.integers(1, N, len(q))
    s_o, c_o = og.find_before_batch(src[q], None, eidx[q])
    s_d, c_d = f.find_before_batch_device(src[q], None, eidx[q])
    assert (s_d.cpu().numpy() == s_o).all() and (c_d.cpu().numpy() == c_o).all()
    s_o, c_o = og.find_before_batch(fake, ts[q], None)
    s_d, c_d = f.find_before_batch_device(fake, ts[q], None)
    assert (s_d.cpu().numpy() == s_o).all() and (c_d.cpu().numpy() == c_o).all()
    for roots, ee, sd in ((src[q], eidx[q], 11), (dst[q], eidx[q], 12), (fake, None, 13)):
        sub = f.find_k_hop(2, roots, ts[q], n, ee, seed=sd, row_offset=7)
        osub = og.find_k_hop(2, roots, ts[q], n, ee, seed=sd, row_offset=7)
        for a, b in zip(sub, osub):
            for x, y in zip(a, b):
                assert x.dtype == y.dtype and (x == y).all()
        scanned = torch.zeros(1, dtype=torch.int64, device="cuda")
        nodes, we, wt, anony, cat = f.find_k_walks_device(n, roots, N2, sub, seed=sd + 50, row_offset=7, scanned=scanned)
        on, oe, ot, oa, osc = og.sample_walks(roots, osub[0][0], osub[1][0], osub[2][0], N2, seed=sd + 50, row_offset=7, want_scanned=True)
        assert (nodes.cpu().numpy() == on).all() and (we.cpu().numpy() == oe).all()
        assert (wt.cpu().numpy() == ot).all() and (anony.cpu().numpy() == oa).all()
        assert (cat.cpu().numpy() == orc.class_ids_prep(oa)[0]).all()
        assert int(scanned.item()) == int(osc.sum())
        assert (tm.edge_identity_device(we).cpu().numpy() == orc.edge_identity(oe)).all()
        # size-independent properties: sampled first-hop timestamps ascend per row and precede the cut
        t0 = sub[2][0]
        assert (np.diff(t0, axis=1) >= 0).all()
        if ee is None:
            assert (t0.astype(np.float64)[sub[0][0] > 0] < np.repeat(ts[q], n).reshape(-1, n)[sub[0][0] > 0] + 1.0).all()


def test_injected_draws_replay(tm, orc):
    """INJECTED mode: indices recorded elsewhere (here: from the oracle's Philox run) reproduce the sample."""
    src, dst, eidx, ts = synth_graph(5, 300, 20000, 5000)
    f = tm.NeighborFinder.from_events(300, src, dst, eidx, ts)
    og = orc.OracleGraph.from_events(300, src, dst, eidx, ts)
    q = np.arange(15000, 15200)
    n = 12
    s, c = og.find_before_batch(src[q], None, eidx[q])
    rec = np.zeros((len(q), n), np.int64)
    for i in range(len(q)):
        for k in range(n):
            rec[i, k] = orc.draw_index(99, 0, i, k, int(c[i])) if c[i] else 0
    a = f.find_k_hop(1, src[q], ts[q], n, eidx[q], inject=[rec], seed=12345)   # seed ignored when injecting
    b = og.find_k_hop(1, src[q], ts[q], n, eidx[q], seed=99)
    for x, y in zip(a, b):
        assert (x[0] == y[0]).all()


def test_walk_pieces_reference_signatures(tm, golden):
    """get_next_step / get_final_step / find_before_walk with the reference's signatures reproduce the pieces of the
    golden find_k_walks run (graph.py:290-296 shows how find_k_walks composes them)."""
    g = golden("rand_small")
    f = finder_of(tm, g)
    q, n, N2 = g["q"], int(g["n"]), int(g["N2"])
    roots = g["src"][q].astype(np.int64)
    h1n, h1e, h1t = g["src_hop0_node"], g["src_hop0_eidx"], g["src_hop0_ts"]
    s2, t2n, e2, t2 = f.get_next_step(h1n.flatten(), h1t.flatten(), N2, n, e_idx_l=h1e.flatten(), source_id=roots, seed=78)
    wn, we, wt = g["src_w_nodes"], g["src_w_eidx"], g["src_w_t"]
    B = len(q)
    assert (s2.reshape(B, -1) == wn[..., 2]).all() and (t2n.reshape(B, -1) == wn[..., 3]).all()
    assert (e2.reshape(B, -1) == we[..., 1]).all() and (t2.reshape(B, -1) == wt[..., 1]).all()
    s3, t3n, e3, t3, anony = f.get_final_step(wn[..., 4], wn[..., 5], wn[..., 2], wn[..., 3], we[..., 2], we[..., 1], wt[..., 2], wt[..., 1], seed=78)
    assert (s3.reshape(B, -1) == wn[..., 0]).all() and (t3n.reshape(B, -1) == wn[..., 1]).all()
    assert (e3.reshape(B, -1) == we[..., 0]).all() and (t3.reshape(B, -1) == wt[..., 0]).all()
    assert (anony.reshape(B, -1, 3) == g["src_w_anony"]).all()
    # find_before_walk: prefixes of [root, neighbour] cut at the first-hop edge; a node that lacks the edge contributes nothing
    root, nb, e = int(roots[3]), int(h1n[3, 2]), int(h1e[3, 2])
    src_a, nbr_a, e_a, ts_a, _ = f.find_before_walk([root, nb], 0.0, e_idx=e)
    d = f.nodeedge2idx
    ca = d[root].get(e, 0) if root > 0 else 0
    cb = d[nb].get(e, 0) if nb > 0 else 0
    assert len(src_a) == ca + cb and (src_a[:ca] == root).all() and (src_a[ca:] == nb).all()
    off = f.off_set_l
    assert (e_a[:ca] == f.edge_idx_l[off[root]:off[root] + ca]).all() and (ts_a[ca:] == f.node_ts_l[off[nb]:off[nb] + cb]).all()


def test_reference_native_mt19937_replay(tm, golden):
    """The unmodified reference ran with its OWN numpy MT19937 stream (np.random.seed(12345)); its draws were only
    recorded.  Feeding them back (INJECTED mode) must reproduce hops, walks and anonymisation bit for bit."""
    g = golden("native_rng")
    f = finder_of(tm, g)
    q, n, N2 = g["q"], int(g["n"]), int(g["N2"])
    ts = g["ts"].astype(np.float64)
    for r in ("src", "bgd"):
        roots = (g["src"][q] if r == "src" else g["fake"]).astype(np.int64)
        e = g["eidx"][q] if r == "src" else None
        sub = f.find_k_hop(2, roots, ts[q], n, e, inject=[g[f"{r}_inj_hop0"], g[f"{r}_inj_hop1"]])
        assert_sub(g, r, sub)
        walks = f.find_k_walks(n, roots, N2, sub, inject2=g[f"{r}_inj_step2"], inject3=g[f"{r}_inj_step3"])
        assert_walks(g, r, walks)


def test_native_philox_distribution_chi_square(tm, golden):
    """Native-RNG mode: the 12-class motif histogram under our Philox stream is statistically indistinguishable from
    the histogram the reference produced with its MT19937 stream (two-sample chi-square, alpha = 1e-3), and the
    sampled first-hop indices are uniform over their windows."""
    from scipy import stats
    g = golden("nullmodel")
    ref = g["dist"] * 500 * 3 * int(g["n"])                      # counts of the reference run (Philox-contract draws)
    ti = g["test_idx"]
    tot = np.zeros(12)
    for seed in range(40, 48):                                    # 8 independent native runs
        f = finder_of(tm, g, seed=seed)
        d = tm.pre_processing(f, None, g["src"][ti].astype(np.int64), g["dst"][ti].astype(np.int64),
                              g["ts"][ti].astype(np.float64), g["eidx"][ti], int(g["n"]), fakes=g["fakes"])
        tot += np.array([d[k] for k in range(1, 13)]) * 500 * 3 * int(g["n"])
    keep = (ref + tot) > 40                                       # pool sparse classes out of the test
    table = np.stack([ref[keep], tot[keep]])
    chi2, p, dof, _ = stats.chi2_contingency(table)
    assert p > 1e-3, (chi2, p, dof)
    # uniformity of first-hop draws: position / window length ~ U(0,1)
    src, dst, eidx, ts = synth_graph(4, 60, 40000, 10 ** 6)
    f = tm.NeighborFinder.from_events(60, src, dst, eidx, ts, seed=3)
    q = np.arange(30000, 34000)
    start, cut = f.find_before_batch_device(src[q], None, eidx[q])
    sub = f.find_k_hop(1, src[q], ts[q], 16, eidx[q])
    off = f.off_set_l; e_sorted = f.edge_idx_l
    cutn = cut.cpu().numpy(); st = start.cpu().numpy()
    u = []
    for i in range(0, len(q), 7):
        if cutn[i] < 50:
            continue
        win = e_sorted[st[i]:st[i] + cutn[i]]
        pos = {int(e): k for k, e in enumerate(win)}
        u += [(pos[int(e)] + 0.5) / cutn[i] for e in sub[1][0][i]]
    ks = stats.kstest(np.array(u), "uniform")
    assert ks.pvalue > 1e-3, ks


# ------------------------------------------------------------------------------------------ encoder
class _Base:
    def __init__(self, nfeat, efeat):
        self.n_feat_th = torch.as_tensor(nfeat).cuda(); self.e_feat_th = torch.as_tensor(efeat).cuda()
        self.node_raw_features = torch.nn.Embedding.from_pretrained(self.n_feat_th, padding_idx=0, freeze=True)
        self.edge_raw_features = torch.nn.Embedding.from_pretrained(self.e_feat_th, padding_idx=0, freeze=True)


def explainer_from_golden(tm, g):
    base = _Base(g["node_feat"], g["edge_feat"])
    m = tm.TempME(base, "tgn", "unit", out_dim=40, hid_dim=int(g["hid_dim"]) if "hid_dim" in g else 64, device="cuda", use_temporal_guidance=bool(g["use_temporal"]),
                  if_cat_feature=bool(g["if_cat"]) if "if_cat" in g else True, null_model={k: 1 / 12 for k in range(1, 13)})
    sd = {k[2:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p:")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected                       # every reference parameter name exists in our module
    return m.cuda().eval()


@pytest.mark.parametrize("tag", ["d172", "d32", "d32_plainattn", "d32_nocat", "d32_hid32"])
def test_encoder_golden(tm, golden, tag):
    g = golden("encoder_" + tag)
    m = explainer_from_golden(tm, g)
    walks = (g["w_nodes"].astype(np.int64), g["w_eidx"].astype(np.int64), g["w_t"].astype(np.float64),
             g["w_cat"].astype(np.int64)[..., None], None)
    with torch.no_grad():
        out = m(walks, g["cut_time"], g["edge_identity"].astype(np.float64))
    assert out.shape == g["score"].shape and out.dtype == torch.float32
    np.testing.assert_allclose(out.cpu().numpy(), g["score"], rtol=1e-5, atol=0)


@pytest.mark.parametrize("D,Ed,if_cat,use_temporal,hid", [(32, 32, True, True, 64), (172, 172, True, True, 64), (100, 7, True, True, 64),
                                                          (32, 32, False, False, 64), (172, 1, False, True, 64), (64, 32, False, True, 64),
                                                          (32, 32, True, True, 32), (172, 172, True, True, 32), (100, 7, False, True, 32),
                                                          (64, 4, True, False, 32)])
def test_encoder_vs_oracle_larger(tm, orc, D, Ed, if_cat, use_temporal, hid):
    """Scores for many roots in reference batches of 100 (ragged last batch) vs the numpy oracle; with and without the category one-hot
    (if_cat_feature, explainer.py:122-126) and the temporal weighting (use_temporal_guidance)."""
    from oracle import encoder as orc_enc
    rng = np.random.default_rng(D)
    src, dst, eidx, ts = synth_graph(7, 400, 30000, 10 ** 6)
    f = tm.NeighborFinder.from_events(400, src, dst, eidx, ts)
    q = np.arange(20000, 20250)
    n, N2 = 10, 3
    sub = f.find_k_hop_device(1, src[q], ts[q], n, eidx[q], seed=3)
    nodes, we, wt, anony, cat = f.find_k_walks_device(n, src[q], N2, sub, seed=4)
    eid = tm.edge_identity_device(we)
    nfeat = rng.standard_normal((400, D)).astype(np.float32); efeat = rng.standard_normal((30001, Ed)).astype(np.float32)
    nfeat[0] = 0; efeat[0] = 0
    torch.manual_seed(D)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "unit", 40, hid, device="cuda", null_model={}, if_cat_feature=if_cat,
                  use_temporal_guidance=use_temporal).cuda().eval()
    with torch.no_grad():
        m.time_encoder.phase.normal_(0, 0.1)
    cut = torch.as_tensor(ts[q].astype(np.float32)).cuda()
    scores = m.score_device(nodes, we, wt, cat if if_cat else None, cut, eid, group=100).cpu().numpy()
    p = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    for s in range(0, len(q), 100):
        sl = slice(s, s + 100)
        walks = (nodes[sl].cpu().numpy(), we[sl].cpu().numpy(), wt[sl].cpu().numpy(), cat[sl].cpu().numpy(), None)
        ref = orc_enc.forward(p, nfeat, efeat, walks, ts[q][sl], eid[sl].cpu().numpy(), if_cat=if_cat, use_temporal=use_temporal)
        np.testing.assert_allclose(scores[sl], ref[..., 0], rtol=1e-5, atol=0)


def test_pipeline_matches_piecewise(tm, orc):
    """MotifPipeline (the bench path) == the reference-facing calls made one by one."""
    src, dst, eidx, ts = synth_graph(9, 185, 40000, 10 ** 7)
    f = tm.NeighborFinder.from_events(185, src, dst, eidx, ts)
    rng = np.random.default_rng(1)
    nfeat = rng.standard_normal((185, 32)).astype(np.float32); efeat = rng.standard_normal((40001, 32)).astype(np.float32)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "unit", 40, 64, device="cuda", null_model={}).cuda().eval()
    q = np.arange(30000, 30200)
    fake = rng.integers(1, 185, len(q))
    pipe = tm.MotifPipeline(f, m, 30, 1, group=100, seed=42)
    scores = pipe.run_host(src[q], dst[q], fake, ts[q], eidx[q])
    assert scores.shape == (3, 200, 30)
    assert int(pipe.hist_null.sum().item()) == 3 * 200 * 30 == int(pipe.hist_prep.sum().item())
    og = orc.OracleGraph.from_events(185, src, dst, eidx, ts)
    hist = np.zeros(12, np.int64)
    for b in range(2):                                   # reference batches of 100 events
        sl = slice(100 * b, 100 * b + 100)
        for k, (roots, ee) in enumerate(((src[q], eidx[q]), (dst[q], eidx[q]), (fake, None))):
            off = b * 300 + k * 100                      # batch-major global row of the first root
            osub = og.find_k_hop(1, roots[sl], ts[q][sl], 30, None if ee is None else ee[sl], seed=42, row_offset=off)
            on, oe, ot, oa = og.sample_walks(roots[sl], osub[0][0], osub[1][0], osub[2][0], 1, seed=43, row_offset=off)
            hist += orc.class_hist_null(oa)
            cat = orc.class_ids_prep(oa)[0]
            walks = (on.astype(np.int64), oe.astype(np.int64), ot.astype(np.float64), cat[..., None].astype(np.int64), None)
            with torch.no_grad():
                ref = m(walks, ts[q][sl], orc.edge_identity(oe))
            np.testing.assert_allclose(scores[k, sl], ref[..., 0].cpu().numpy(), rtol=1e-6, atol=0)
    assert (pipe.hist_null.cpu().numpy() == hist).all()
    # asynchronous host API: two batches in flight give the scores of run_host
    q2 = np.arange(31000, 31200)
    ref1, ref2 = pipe.run_host(src[q], dst[q], fake, ts[q], eidx[q]).copy(), pipe.run_host(src[q2], dst[q2], fake, ts[q2], eidx[q2]).copy()
    t1 = pipe.submit_host(src[q], dst[q], fake, ts[q], eidx[q])
    t2 = pipe.submit_host(src[q2], dst[q2], fake, ts[q2], eidx[q2])
    a1 = pipe.collect(t1).copy()
    t3 = pipe.submit_host(src[q], dst[q], fake, ts[q], eidx[q])
    a2, a3 = pipe.collect(t2).copy(), pipe.collect(t3).copy()
    assert (a1 == ref1).all() and (a2 == ref2).all() and (a3 == ref1).all()


# ---------------------------------------------------------------------------------------------
# motif -> edge aggregation: retrieve_edge_imp_node, eval mode
# ---------------------------------------------------------------------------------------------
def _edge_imp_model(tm, z, D, dep):
    n_nodes = 4
    nfeat = torch.zeros(n_nodes, D); efeat = torch.as_tensor(z["edge_feat"])

    class Base:
        n_feat_th = nfeat.cuda(); e_feat_th = efeat.cuda()
        node_raw_features = torch.nn.Embedding.from_pretrained(n_feat_th, padding_idx=0, freeze=True)
        edge_raw_features = torch.nn.Embedding.from_pretrained(e_feat_th, padding_idx=0, freeze=True)

    m = tm.TempME(Base(), "tgn", "t", 40, 64, device="cuda:0", null_model={}, use_dependency_aware_sampling=dep).cuda().eval()
    sd = {k[2:]: torch.as_tensor(z[k]) for k in z if k.startswith("p:")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected or not dep          # without dependency-aware sampling the module has no edge_dependency_gcn
    return m


@pytest.mark.parametrize("tag,D", [("d32", 32), ("d172", 172)])
def test_edge_importance_golden(tm, golden, tag, D):
    """CUDA retrieve_edge_imp_node (gate on tcgen05, per-root segmented max) == the unmodified reference, rtol 1e-5."""
    z = golden(f"edgeimp_{tag}")
    sub = ([z["h0_node"], z["h1_node"]], [z["h0_eidx"], z["h1_eidx"]], None)
    walks = (None, z["w_eidx"], z["w_t"], None, None)
    for key, dep in (("dep", True), ("nodep", False)):
        m = _edge_imp_model(tm, z, D, dep)
        i0, i1 = m.retrieve_edge_imp_node(sub, z[f"{key}_score"], walks, training=False)
        np.testing.assert_allclose(i0.cpu().numpy(), z[f"{key}_imp0"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(i1.cpu().numpy(), z[f"{key}_imp1"], rtol=1e-5, atol=1e-7)
    s0, s1 = m.retrieve_edge_imp_node(sub, z["nodep_score"], walks, training=True)       # Beta draws (tests/test_gpu_train.py checks their law)
    live = torch.as_tensor(z["h0_node"]).cuda() != 0
    assert s0.shape == i0.shape and bool((s0[~live] == 0).all()) and bool(((s0[live] > 0) & (s0[live] < 1)).all())


def test_edge_importance_vs_oracle_larger(tm, orc):
    """Seeded inputs against oracle/encoder.edge_importance: ragged ids, padding slots, ids no walk carries."""
    from oracle import encoder as enc
    rng = np.random.default_rng(11)
    B, W, n, D, Ed, Ne = 300, 30, 12, 32, 32, 700
    z = {"edge_feat": rng.standard_normal((Ne, Ed)).astype(np.float32)}
    m = _edge_imp_model(tm, z, D, True)
    with torch.no_grad():
        m.time_encoder.phase.copy_(0.1 * torch.randn(D))
    p = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    w_eidx = rng.integers(0, Ne, (B, W, 3)).astype(np.int32)
    w_eidx[rng.random((B, W, 3)) < 0.3] = 0
    w_t = np.sort(rng.integers(1e8, 1.1e8, (B, W, 3)).astype(np.float32), axis=-1)
    scores = rng.random((B, W, 1)).astype(np.float32)
    h_e = [rng.integers(0, Ne, (B, n)).astype(np.int32), rng.integers(0, Ne, (B, n * n)).astype(np.int32)]
    h_n = [rng.integers(0, 5, (B, n)).astype(np.int32), rng.integers(0, 5, (B, n * n)).astype(np.int32)]
    h_e[0][:, :4] = w_eidx[:, :4, 2]; h_e[1][:, :W] = w_eidx[:, :, 1]
    walks = (None, w_eidx, w_t, None, None)
    r0, r1 = enc.edge_importance(p, z["edge_feat"], (h_n, h_e, None), scores, walks)
    i0, i1 = m.retrieve_edge_imp_node((h_n, h_e, None), scores, walks, training=False)
    np.testing.assert_allclose(i0.cpu().numpy(), r0, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(i1.cpu().numpy(), r1, rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------------------------------------
# full BASELINE size (cfg2: 16,000 query events = 1.44 M motifs per step): size-independent properties
# ---------------------------------------------------------------------------------------------
def test_full_size_cfg2_properties(tm, monkeypatch):
    """At the bench size the oracle is too slow, so check properties: (1) the tensor-core scorer (host-folded chain, TMEM
    operands, dual rounds, gather4 staging, dynamic tiles) agrees with the layer-by-layer fp32 CUDA-core scorer of the same library
    within 1e-5 on every motif; (2) scores do not depend on how the query batches are split into calls (draws are keyed by the
    global row); (3) the class histogram counts every motif once; (4) repeated calls are bit-identical."""
    from tempme_b200 import synth
    from bench import random_params
    g = synth.make_graph("cfg2", 1.0)
    sh = synth.SHAPES["cfg2"]
    f = tm.NeighborFinder.from_events(g["n_nodes"], g["src"], g["dst"], g["eidx"], g["ts"], device="cuda:0", seed=3)
    nfeat, efeat = synth.make_features("cfg2", g["n_nodes"], len(g["src"]))
    m = tm.TempME(_Base(nfeat.numpy(), efeat.numpy()), "tgn", "cfg2", 40, 64, device="cuda", null_model={}).cuda().eval()
    m.load_state_dict({k: torch.as_tensor(v) for k, v in random_params(sh["D"], sh["Ed"]).items()}, strict=False)
    pipe = tm.MotifPipeline(f, m, sh["n"], sh["N2"], group=100, seed=7)
    Q = 16000
    q = synth.make_queries(g, np.random.default_rng(5), Q)
    dq = pipe.stage_queries(*q)
    s_tc = pipe.run_device(*dq).clone()
    assert int(pipe.hist_null.sum().item()) == 3 * Q * pipe.W == int(pipe.hist_prep.sum().item())
    assert torch.isfinite(s_tc).all() and float(s_tc.min()) > 0 and float(s_tc.max()) < 1
    assert torch.equal(pipe.run_device(*dq), s_tc)                       # (4)
    monkeypatch.setenv("TEMPME_ENCODER", "ffma")                          # (1) unfolded fp32 FFMA kernel, same inputs
    s_ff = pipe.run_device(*dq).clone()
    monkeypatch.delenv("TEMPME_ENCODER")
    rel = ((s_tc - s_ff).abs() / s_ff.abs()).max().item()
    assert rel < 1e-5, rel
    # (2) the same 16,000 events as two calls of 8,000 (whole reference batches), global row offsets
    half = 3 * (Q // 2)
    a = pipe.run_device(dq[0][:half], dq[1][:half], dq[2][:half], row_offset=0)
    b = pipe.run_device(dq[0][half:], dq[1][half:], dq[2][half:], row_offset=half)
    assert torch.equal(torch.cat([a, b]), s_tc)


def test_pipeline_explain_matches_oracle(tm, orc):
    """MotifPipeline.explain_device (2-hop subgraph + walks + scores + edge importance on the GPU) == the oracle's pieces."""
    from oracle import encoder as enc
    src, dst, eidx, ts = synth_graph(11, 120, 20000, 10 ** 7)
    f = tm.NeighborFinder.from_events(120, src, dst, eidx, ts)
    rng = np.random.default_rng(2)
    nfeat = rng.standard_normal((120, 32)).astype(np.float32); efeat = rng.standard_normal((20001, 32)).astype(np.float32)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "unit", 40, 64, device="cuda", null_model={}).cuda().eval()
    p = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    q = np.arange(15000, 15100)
    fake = rng.integers(1, 120, len(q))
    n = 8
    pipe = tm.MotifPipeline(f, m, n, 1, group=100, seed=17)
    roots, e, cut64, _ = pipe.stage_queries(src[q], dst[q], fake, ts[q], eidx[q])
    scores, imp0, imp1, sub = pipe.explain_device(roots, e, cut64)
    scores, imp0, imp1 = scores.cpu().numpy(), imp0.cpu().numpy(), imp1.cpu().numpy()
    og = orc.OracleGraph.from_events(120, src, dst, eidx, ts)
    for k, (rr, ee) in enumerate(((src[q], eidx[q]), (dst[q], eidx[q]), (fake, None))):
        off = k * 100
        osub = og.find_k_hop(2, rr, ts[q], n, ee, seed=17, row_offset=off)
        on, oe, ot, oa = og.sample_walks(rr, osub[0][0], osub[1][0], osub[2][0], 1, seed=18, row_offset=off)
        sl = slice(off, off + 100)
        assert (sub[0][1][sl].cpu().numpy() == osub[0][1]).all() and (sub[1][1][sl].cpu().numpy() == osub[1][1]).all()
        r0, r1 = enc.edge_importance(p, efeat, ([osub[0][0], osub[0][1]], [osub[1][0], osub[1][1]], None), scores[sl][..., None], (None, oe, ot, None, None))
        np.testing.assert_allclose(imp0[sl], r0, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(imp1[sl], r1, rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------------------------------------
# enhance path: compute_walk_importance, enhance_predict_walks, enhance_predict_agg, eval mode
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["d32", "d32_hid32"])      # hid32: enhance_main.py's own defaults (--hid_dim 32 --out_dim 32)
def test_enhance_path_golden(tm, golden, tag):
    z = golden("enhance_" + tag)
    hid, od = (int(z["hid_dim"]), int(z["out_dim"])) if "hid_dim" in z else (64, 40)
    m = tm.TempME(_Base(z["node_feat"], z["edge_feat"]), "tgn", "t", od, hid, device="cuda:0", null_model={}).cuda().eval()
    missing, unexpected = m.load_state_dict({k[2:]: torch.as_tensor(z[k]) for k in z if k.startswith("p:")}, strict=False)
    assert not unexpected
    m.node_degree = torch.as_tensor(z["node_degree"]).cuda()
    ws = {pre: (z[f"{pre}_nodes"], z[f"{pre}_eidx"], z[f"{pre}_t"], z[f"{pre}_cat"], None) for pre in ("src", "tgt")}
    w = m.compute_walk_importance(ws["src"][2], ws["src"][0], z["cut_time"])
    np.testing.assert_allclose(w.cpu().numpy(), z["w_src"], rtol=1e-5, atol=1e-7)
    # Tolerances: these outputs are signed sums over the W walks (then two more Linear layers), so the reference's own fp32 evaluation
    # deviates from the exact value of its formula by ~1e-5 of the terms' scale.  Arbiter = the formula in float64 with the reference's
    # fp32 TimeEncode argument (oracle, arg32); the CUDA result may differ from the golden by no more than twice the golden's own deviation.
    from oracle import encoder as enc
    p = {k[2:]: z[k] for k in z if k.startswith("p:")}
    exact = {}
    for pre in ws:
        emb = m.enhance_predict_walks(ws[pre], z["cut_time"], z[f"{pre}_ei"]).cpu().numpy()
        exact[pre] = enc.enhance_predict_walks(p, z["node_feat"], z["edge_feat"], ws[pre], z["cut_time"], z[f"{pre}_ei"], z["node_degree"], dtype=np.float64, arg32=True)
        scale = np.abs(exact[pre][:, :hid]).max(1, keepdims=True)
        dev = np.abs(z[f"emb_{pre}"] - exact[pre])
        print(f"enhance golden {tag} {pre}: |cuda - exact| / row scale = {(np.abs(emb - exact[pre]) / scale).max():.2e}, |reference - exact| / row scale = {(dev / scale).max():.2e}")
        assert (np.abs(emb - exact[pre]) <= 1e-5 * scale).all()
        assert (np.abs(emb - z[f"emb_{pre}"]) <= 1e-5 * scale + dev).all()
    with torch.no_grad():
        pos, neg = m.enhance_predict_agg(z["cut_time"], ws["src"], ws["tgt"], ws["src"], (z["src_ei"], z["tgt_ei"], z["src_ei"]), z["src_gat"], z["tgt_gat"], z["bgd_gat"])
    cat64 = lambda a, b: np.concatenate([a, np.asarray(b, np.float64)], axis=-1)
    ex_pos = enc.affinity_score(p, cat64(exact["src"], z["src_gat"]), cat64(exact["tgt"], z["tgt_gat"]), dtype=np.float64)
    ex_neg = enc.affinity_score(p, cat64(exact["src"], z["src_gat"]), cat64(exact["src"], z["bgd_gat"]), dtype=np.float64)
    for got, gold, ex in ((pos, z["pos"], ex_pos), (neg, z["neg"], ex_neg)):       # the affinity logits, relative to the logits' scale
        dev, bound = np.abs(gold - ex), 1e-5 * max(1.0, float(np.abs(ex).max()))
        print(f"enhance golden {tag} logits: |cuda - exact| max {np.abs(got.cpu().numpy() - ex).max():.2e}, |reference - exact| max {dev.max():.2e}, bound {bound:.2e}")
        assert (np.abs(got.cpu().numpy() - ex) <= bound).all()
        assert (np.abs(got.cpu().numpy() - gold) <= bound + dev).all()


@pytest.mark.parametrize("hid", [64, 32])      # 32: enhance_main.py's default (--hid_dim, enhance_main.py:66)
def test_enhance_walks_vs_oracle_larger(tm, orc, hid):
    """Seeded inputs at several hundred roots against oracle/encoder.enhance_predict_walks (padding nodes, ragged degrees)."""
    from oracle import encoder as enc
    rng = np.random.default_rng(21)
    B, W, D, Ed, Nn, Ne = 300, 30, 32, 32, 150, 900
    nfeat = rng.standard_normal((Nn, D)).astype(np.float32); efeat = rng.standard_normal((Ne, Ed)).astype(np.float32)
    nfeat[0] = 0; efeat[0] = 0
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "t", 40, hid, device="cuda:0", null_model={}).cuda().eval()
    m.node_degree = torch.as_tensor(rng.integers(1, 80, Nn).astype(np.float32)).cuda()
    p = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    nodes = rng.integers(0, Nn, (B, W, 6)); nodes[rng.random((B, W, 6)) < 0.15] = 0
    eidx = rng.integers(0, Ne, (B, W, 3))
    t = np.sort(rng.integers(1e8, 1.1e8, (B, W, 3)).astype(np.float64), -1)
    cat = rng.integers(0, 12, (B, W, 1)); cut = t[:, :, 2].max(1) + rng.integers(1, 1000, B)
    eid = rng.integers(0, W, (B, W, 3, 3)).astype(np.float64)
    walks = (nodes, eidx, t, cat, None)
    deg = m.node_degree.cpu().numpy()
    ref = enc.enhance_predict_walks(p, nfeat, efeat, walks, cut, eid, deg)
    got = m.enhance_predict_walks(walks, cut, eid).cpu().numpy()
    # The embedding is a signed sum over W walks of weight * attention output, so the reference's own fp32 evaluation is only accurate to
    # ~1e-5 of the terms' scale, not of the (partly cancelled) sum.  Arbiter: the same formula in float64 with the reference's fp32
    # TimeEncode argument.  The CUDA result must be as close to it as the layer-by-layer fp32 evaluation is, and within 1e-5 of it
    # relative to the row's scale.
    exact = enc.enhance_predict_walks(p, nfeat, efeat, walks, cut, eid, deg, dtype=np.float64, arg32=True)
    e_ref, e_got = np.abs(ref - exact), np.abs(got - exact)
    scale = np.abs(exact[:, :hid]).max(1, keepdims=True)          # the embedding's own scale (max-norm of the row)
    print(f"enhance hid={hid}: max |cuda - exact| / row scale = {(e_got / scale).max():.2e}, max |fp32 reference formula - exact| / row scale = "
          f"{(e_ref / scale).max():.2e}, row scale {scale.min():.2f} .. {scale.max():.2f}")
    assert (e_got <= 1e-5 * scale).all(), (e_got / scale).max()                      # the 1e-5 contract, relative to the vector it belongs to
    assert (np.abs(got - ref) <= 1e-5 * scale + e_ref).all()                         # and against the fp32 oracle, allowing for ITS distance to exact


# ---------------------------------------------------------------------------------------------
# offline explanation pack: pre_processing + marginal + calculate_edge of processed/data_preprocess.py
# ---------------------------------------------------------------------------------------------
def test_build_pack_matches_oracle(tm, orc, tmp_path):
    src, dst, eidx, ts = synth_graph(13, 150, 25000, 10 ** 6)
    f = tm.NeighborFinder.from_events(150, src, dst, eidx, ts)
    og = orc.OracleGraph.from_events(150, src, dst, eidx, ts)
    rng = np.random.default_rng(4)
    q = np.arange(20000, 20060)
    fake = rng.integers(1, 150, len(q))
    n, N2 = 6, 3
    pack, edge = tm.build_pack(f, src[q], dst[q], ts[q], eidx[q], fake, n, N2, seed=100)
    assert sorted(pack) == sorted(tm.pack.PACK_KEYS) and edge.shape == (3, len(q), n * N2, 3, 3)
    anonys, walks = [], {}
    for k, (name, roots, ee) in enumerate((("src", src[q], eidx[q]), ("tgt", dst[q], eidx[q]), ("bgd", fake, None))):
        osub = og.find_k_hop(2, roots, ts[q], n, ee, seed=100 + 2 * k)
        on, oe, ot, oa = og.sample_walks(roots, osub[0][0], osub[1][0], osub[2][0], N2, seed=100 + 2 * k + 1)
        for l in range(2):                                    # [node | eidx | t] (data_preprocess.py:115-116)
            ref = np.concatenate([osub[0][l], osub[1][l], osub[2][l]], axis=-1).astype(np.float64)
            assert (pack[f"subgraph_{name}_{l}"] == ref).all()
        anonys.append(oa); walks[name] = (on, oe, ot, oa)
        assert (edge[k] == orc.edge_identity(oe)).all()       # calculate_edge (:345-357)
    hist = sum(np.bincount(orc.class_ids_prep(a)[0].ravel(), minlength=12) for a in anonys)
    freq = hist / (len(q) * n * N2 * 3)                        # marginal (:190-192)
    for name, (on, oe, ot, oa) in walks.items():
        cat = orc.class_ids_prep(oa)[0]
        ref = np.concatenate([on.astype(np.float64), oe.astype(np.float64), ot.astype(np.float64), cat[..., None].astype(np.float64), freq[cat][..., None]], axis=-1)
        assert (pack[f"walks_{name}_new"] == ref).all()
    # the files temp_exp_main.py:705-714 opens: {data}_{mode}_cat.h5 (HDF5) and {data}_{mode}_edge.npy; the reference loader's slicing
    # (utils/batch_loader.py:120-201: file[name][:]) on what comes back
    path, edge_path = tm.save_pack(pack, edge, str(tmp_path), "unit", "test")
    assert path.endswith("unit_test_cat.h5")
    z = tm.load_pack(path)
    for k in pack:
        assert np.array_equal(z[k][:], pack[k]) and z[k].dtype == np.asarray(pack[k]).dtype, k
    x0 = z["subgraph_src_0"][:]
    assert x0[:, 0:n].shape == (len(q), n) and (z["walks_src_new"][:][:, :, 12:13].astype(int) == orc.class_ids_prep(walks["src"][3])[0][..., None]).all()
    assert np.load(edge_path).shape == edge.shape


# ---------------------------------------------------------------------------------------------
# kl_loss forward value (the part the reference's eval loops use)
# ---------------------------------------------------------------------------------------------
def test_kl_loss_golden_and_oracle(tm, golden):
    from oracle import encoder as enc
    z = golden("kl_loss")
    null = {k + 1: float(v) for k, v in enumerate(z["null_values"])}
    nfeat = np.zeros((4, 8), np.float32)
    for prior in ("empirical", "uniform"):
        m = tm.TempME(_Base(nfeat, nfeat), "tgn", "t", 8, 16, prior=prior, device="cuda:0", null_model=null).cuda().eval()
        for name in ("us", "few", "one"):
            for target in (0.3, 0.05):
                v = m.kl_loss(torch.as_tensor(z[f"{name}_prob"]).cuda(), (None, None, None, z[f"{name}_cat"][:, :, None], None), target=target)
                assert v.shape == () and v.is_cuda
                np.testing.assert_allclose(v.item(), float(z[f"{name}_{prior}_{target}"]), rtol=1e-5, atol=1e-7)
        rng = np.random.default_rng(5)                       # a full batch of cfg2's shape against the oracle
        B, W = 3000, 30
        prob = rng.random((B, W, 1)).astype(np.float32); cat = rng.integers(0, 12, (B, W, 1))
        v = m.kl_loss(prob, (None, None, None, cat, None), target=0.3)
        np.testing.assert_allclose(v.item(), enc.kl_loss(prob, cat, z["null_values"], 0.3, prior), rtol=1e-5, atol=1e-7)


def test_score_gather_peer_stores_single_gpu(tm):
    """tm_encode_score_gather: the scorer writes every score to `out` and to each peer segment (here: other buffers of the same GPU);
    results identical to the plain call.  The two-GPU version over symmetric memory is tests/test_gpu_multi.py."""
    rng = np.random.default_rng(31)
    B, W, D, Ed, Nn, Ne = 77, 30, 32, 32, 120, 600
    nfeat = rng.standard_normal((Nn, D)).astype(np.float32); efeat = rng.standard_normal((Ne, Ed)).astype(np.float32)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "t", 40, 64, device="cuda:0", null_model={}).cuda().eval()
    dev = torch.device("cuda:0")
    nodes = torch.as_tensor(rng.integers(0, Nn, (B, W, 6)).astype(np.int32)).to(dev)
    eidx = torch.as_tensor(rng.integers(0, Ne, (B, W, 3)).astype(np.int32)).to(dev)
    t = torch.as_tensor(np.sort(rng.integers(1e8, 1.1e8, (B, W, 3)), -1).astype(np.float32)).to(dev)
    cat = torch.as_tensor(rng.integers(0, 12, (B, W)).astype(np.uint8)).to(dev)
    cut = (t[:, :, 2].max(1).values + 5).contiguous()
    eid = torch.as_tensor(rng.integers(0, 4, (B, W, 3, 3)).astype(np.float32)).to(dev)
    ref = m.score_device(nodes, eidx, t, cat, cut, eid, group=B)
    world, rank = 4, 2
    bufs = [torch.full((world, B, W), float("nan"), device=dev) for _ in range(world)]
    from tempme_b200.dist import segment_offsets
    peers = segment_offsets(B, W, world, rank, [b.data_ptr() for b in bufs])
    assert len(peers) == world - 1
    out = m.score_device(nodes, eidx, t, cat, cut, eid, group=B, out=bufs[rank][rank], peer_ptrs=peers)
    torch.cuda.synchronize()
    assert out.data_ptr() == bufs[rank][rank].data_ptr()
    for b in bufs:
        assert torch.equal(b[rank], ref)
        assert torch.isnan(b[[r for r in range(world) if r != rank]]).all()
    with pytest.raises(Exception):
        m.score_device(nodes, eidx, t, cat, cut, eid, group=B, peer_ptrs=[bufs[0].data_ptr()] * 8)


def test_get_next_step_time_cut_golden(tm, golden):
    """get_next_step(e_idx_l=None) (utils/graph.py:308-333): prefixes of [root, neighbour] cut by time, against the unmodified reference."""
    g, z = golden("rand_small"), golden("nextstep_time")
    f = finder_of(tm, g)
    out = f.get_next_step(z["nbr"], z["cut"], int(z["N2"]), int(z["degree"]), e_idx_l=None, source_id=z["roots"], seed=int(z["seed"]))
    for a, name in zip(out, ("o_src", "o_tgt", "o_eidx", "o_ts")):
        ref = z[name]
        assert a.shape == ref.shape and a.dtype == ref.dtype and (a == ref).all(), name


@pytest.mark.parametrize("D,Ed", [(32, 32), (172, 172), (172, 1), (100, 7), (64, 32)])
def test_edge_projection_mode_equals_plain_mode(tm, D, Ed):
    """The scorer with lin_event's edge columns applied once per edge id (projected table, default) and with the raw feature rows
    (TEMPME_EDGE_PROJECTION=0 / edge_projection = False): same scores to fp32 round-off; the table follows weight updates."""
    rng = np.random.default_rng(D + Ed)
    src, dst, eidx, ts = synth_graph(17, 300, 20000, 10 ** 6)
    f = tm.NeighborFinder.from_events(300, src, dst, eidx, ts)
    q = np.arange(15000, 15300)
    sub = f.find_k_hop_device(1, src[q], ts[q], 10, eidx[q], seed=3)
    nodes, we, wt, anony, cat = f.find_k_walks_device(10, src[q], 3, sub, seed=4)
    eid = tm.edge_identity_device(we)
    nfeat = rng.standard_normal((300, D)).astype(np.float32); efeat = rng.standard_normal((20001, Ed)).astype(np.float32)
    nfeat[0] = 0; efeat[0] = 0
    torch.manual_seed(1)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "unit", 40, 64, device="cuda", null_model={}).cuda().eval()
    cut = torch.as_tensor(ts[q].astype(np.float32)).cuda()
    assert m.edge_projection
    a = m.score_device(nodes, we, wt, cat, cut, eid, group=100).clone()
    assert m._proj is not None and tuple(m._proj.shape) == (20001, D)
    m.edge_projection = False
    b = m.score_device(nodes, we, wt, cat, cut, eid, group=100).clone()
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-6, atol=0)
    m.edge_projection = True
    with torch.no_grad():
        m.event_conv.lin_event.weight.mul_(1.5)             # a weight update: the packed blob and the projected table must follow
    c = m.score_device(nodes, we, wt, cat, cut, eid, group=100).clone()
    m.edge_projection = False
    d = m.score_device(nodes, we, wt, cat, cut, eid, group=100).clone()
    assert float((c - a).abs().max()) > 1e-4
    np.testing.assert_allclose(c.cpu().numpy(), d.cpu().numpy(), rtol=2e-6, atol=0)


@pytest.mark.parametrize("D,Ed,hid,proj,N2", [(32, 32, 64, True, 3), (32, 32, 64, False, 3), (172, 172, 64, True, 5), (172, 172, 64, False, 2),
                                              (100, 7, 32, True, 3), (32, 32, 32, True, 4), (64, 32, 64, True, 3),
                                              (30, 6, 64, True, 3), (33, 5, 64, False, 3)])     # row sizes that are not multiples of 16 bytes: plain-load gathers
def test_walk_group_mode_equals_plain_mode_and_oracle(tm, orc, D, Ed, hid, proj, N2):
    """tm_encoder_desc.walk_fanout: the N2 walks of a first-hop slot share their event next to the root, evaluated once per group.  Same
    scores as the per-walk evaluation (fp32 round-off of one reassociated sum) and as the oracle; the hint is only a hint: with the
    walks of every root shuffled (groups no longer uniform, or only by chance) every tile falls back and the scores do not change."""
    from oracle import encoder as orc_enc
    rng = np.random.default_rng(D + Ed + N2)
    src, dst, eidx, ts = synth_graph(23, 300, 20000, 10 ** 6)
    f = tm.NeighborFinder.from_events(300, src, dst, eidx, ts)
    q = np.arange(15000, 15450)                       # 450 roots x 10 x N2 walks: the last tile of slots is ragged
    n = 10
    sub = f.find_k_hop_device(1, src[q], ts[q], n, eidx[q], seed=3)
    nodes, we, wt, anony, cat = f.find_k_walks_device(n, src[q], N2, sub, seed=4)
    eid = tm.edge_identity_device(we)
    nfeat = rng.standard_normal((300, D)).astype(np.float32); efeat = rng.standard_normal((20001, Ed)).astype(np.float32)
    nfeat[0] = 0; efeat[0] = 0
    torch.manual_seed(2)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "unit", 40, hid, device="cuda", null_model={}).cuda().eval()
    m.edge_projection = proj
    cut = torch.as_tensor(ts[q].astype(np.float32)).cuda()
    assert m.detect_fanout(we, nodes) % N2 == 0
    plain = m.score_device(nodes, we, wt, cat, cut, eid, group=100).clone()
    grouped = m.score_device(nodes, we, wt, cat, cut, eid, group=100, fanout=N2).clone()
    np.testing.assert_allclose(grouped.cpu().numpy(), plain.cpu().numpy(), rtol=2e-6, atol=0)
    p = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    for s in range(0, len(q), 100):
        sl = slice(s, s + 100)
        walks = (nodes[sl].cpu().numpy(), we[sl].cpu().numpy(), wt[sl].cpu().numpy(), cat[sl].cpu().numpy(), None)
        ref = orc_enc.forward(p, nfeat, efeat, walks, ts[q][sl], eid[sl].cpu().numpy())
        np.testing.assert_allclose(grouped[sl].cpu().numpy(), ref[..., 0], rtol=1e-5, atol=0)
    # a wrong hint: shuffle the walks of every root (the same permutation for all per-walk tensors)
    W = n * N2
    perm = torch.stack([torch.randperm(W, generator=torch.Generator().manual_seed(i)) for i in range(len(q))]).cuda()
    take = lambda x: torch.gather(x, 1, perm.view(len(q), W, *([1] * (x.dim() - 2))).expand_as(x)).contiguous()
    nodes2, we2, wt2, cat2, eid2 = take(nodes), take(we), take(wt), take(cat), take(eid)
    plain2 = m.score_device(nodes2, we2, wt2, cat2, cut, eid2, group=100).clone()
    hinted2 = m.score_device(nodes2, we2, wt2, cat2, cut, eid2, group=100, fanout=N2).clone()
    np.testing.assert_allclose(plain2.cpu().numpy(), take(plain).cpu().numpy(), rtol=2e-6, atol=0)
    assert torch.equal(hinted2, plain2) or np.allclose(hinted2.cpu().numpy(), plain2.cpu().numpy(), rtol=2e-6, atol=0)
    # half of the roots shuffled: shared and repeated tiles in one launch
    half = len(q) // 2
    mix = lambda a, b: torch.cat([a[:half], b[half:]]).contiguous()
    mixed = m.score_device(mix(nodes, nodes2), mix(we, we2), mix(wt, wt2), mix(cat, cat2), cut, mix(eid, eid2), group=100, fanout=N2)
    np.testing.assert_allclose(mixed.cpu().numpy(), mix(plain, plain2).cpu().numpy(), rtol=2e-6, atol=0)
    # the enhance path's hidden-vector output goes through the same kernel
    emb_a = m.enhance_predict_walks((nodes, we, wt, cat, None), ts[q], eid)
    m._with_fanout = lambda desc, fanout: desc           # per-walk evaluation
    emb_b = m.enhance_predict_walks((nodes, we, wt, cat, None), ts[q], eid)
    scale = float(emb_b.abs().max())
    assert float((emb_a - emb_b).abs().max()) <= 2e-6 * scale


def test_walk_group_mode_many_tiles_per_cta(tm):
    """More tiles of walk groups than resident CTAs (the next tile's first pass is prefetched behind the last sub-tile's rounds)."""
    rng = np.random.default_rng(5)
    src, dst, eidx, ts = synth_graph(29, 500, 40000, 10 ** 6)
    f = tm.NeighborFinder.from_events(500, src, dst, eidx, ts)
    q = rng.integers(20000, 40000, 4000)
    q.sort()
    n, N2 = 20, 3
    sub = f.find_k_hop_device(1, src[q], ts[q], n, eidx[q], seed=3)
    nodes, we, wt, anony, cat = f.find_k_walks_device(n, src[q], N2, sub, seed=4)      # 240,000 motifs = 80,000 slots = 625 tiles
    eid = tm.edge_identity_device(we)
    nfeat = rng.standard_normal((500, 32)).astype(np.float32); efeat = rng.standard_normal((40001, 32)).astype(np.float32)
    torch.manual_seed(3)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "unit", 40, 64, device="cuda", null_model={}).cuda().eval()
    cut = torch.as_tensor(ts[q].astype(np.float32)).cuda()
    plain = m.score_device(nodes, we, wt, cat, cut, eid, group=100).clone()
    grouped = m.score_device(nodes, we, wt, cat, cut, eid, group=100, fanout=N2).clone()
    np.testing.assert_allclose(grouped.cpu().numpy(), plain.cpu().numpy(), rtol=2e-6, atol=0)


def test_three_hops_one_call_vs_oracle_and_hop_by_hop(tm, orc):
    """find_k_hop(3) through tm_sample_khop (one C call) equals the oracle and the hop-by-hop route (tm_sample_hop per level), with a shard
    offset and both cut modes; an empty batch returns empty records."""
    src, dst, eidx, ts = synth_graph(31, 200, 8000, 10 ** 5)
    f = tm.NeighborFinder.from_events(200, src, dst, eidx, ts)
    og = orc.OracleGraph.from_events(200, src, dst, eidx, ts)
    q = np.arange(6000, 6040)
    n = 4
    for ee in (eidx[q], None):
        sub = f.find_k_hop(3, src[q], ts[q], n, ee, seed=5, row_offset=11)
        osub = og.find_k_hop(3, src[q], ts[q], n, ee, seed=5, row_offset=11)
        for a, b in zip(sub, osub):
            assert len(a) == 3
            for x, y in zip(a, b):
                assert x.shape == y.shape and (x == y).all()
        # hop by hop: the same draws (stage = level, row = row_offset * n^level + i)
        x0 = f.sample_hop_device(src[q], None if ee is not None else ts[q], n, ee, seed=5, stage=0, row_offset=11)
        x1 = f.sample_hop_device(x0[0].reshape(-1), None, n, x0[1].reshape(-1), seed=5, stage=1, row_offset=11 * n)
        x2 = f.sample_hop_device(x1[0].reshape(-1), None, n, x1[1].reshape(-1), seed=5, stage=2, row_offset=11 * n * n)
        for lvl, x in enumerate((x0, x1, x2)):
            for arr, ref in zip(x, (sub[0][lvl], sub[1][lvl], sub[2][lvl])):
                assert (arr.cpu().numpy().reshape(ref.shape) == ref).all()
    empty = f.find_k_hop(2, np.zeros(0, np.int64), np.zeros(0), n, None)
    assert all(len(r) == 2 and r[0].shape == (0, n) and r[1].shape == (0, n * n) for r in empty)


def test_edge_identity_bytes_equal_floats_and_score_alike(tm):
    """tm_edge_identity_u8: the same counts as bytes; the scorer reads either form (tm_encoder_desc.edge_identity_u8) -- bit-identical scores,
    with and without walk groups."""
    rng = np.random.default_rng(9)
    src, dst, eidx, ts = synth_graph(37, 300, 20000, 10 ** 6)
    f = tm.NeighborFinder.from_events(300, src, dst, eidx, ts)
    q = np.arange(15000, 15300)
    sub = f.find_k_hop_device(1, src[q], ts[q], 10, eidx[q], seed=3)
    nodes, we, wt, anony, cat = f.find_k_walks_device(10, src[q], 3, sub, seed=4)
    eid_f = tm.edge_identity_device(we)
    eid_b = tm.edge_identity_device(we, u8=True)
    assert eid_b.dtype == torch.uint8 and tuple(eid_b.shape) == (300, 30, 3, 4) and torch.equal(eid_b[..., :3].float(), eid_f) and int(eid_b[..., 3].max()) == 0
    nfeat = rng.standard_normal((300, 32)).astype(np.float32); efeat = rng.standard_normal((20001, 32)).astype(np.float32)
    torch.manual_seed(4)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "unit", 40, 64, device="cuda", null_model={}).cuda().eval()
    cut = torch.as_tensor(ts[q].astype(np.float32)).cuda()
    for fan in (None, 3):
        a = m.score_device(nodes, we, wt, cat, cut, eid_f, group=100, fanout=fan).clone()
        b = m.score_device(nodes, we, wt, cat, cut, eid_b, group=100, fanout=fan).clone()
        assert torch.equal(a, b)
    with pytest.raises(NotImplementedError):
        tm.edge_identity_device(torch.zeros((1, 256, 3), dtype=torch.int32, device="cuda"), u8=True)


@pytest.mark.parametrize("B,n,N2", [(1, 2, 3), (3, 5, 2), (7, 20, 5)])
def test_walk_group_mode_tiny_batches(tm, B, n, N2):
    """A handful of roots (one partial tile of slots) through the walk-group kernel == the per-walk evaluation."""
    rng = np.random.default_rng(B)
    src, dst, eidx, ts = synth_graph(41, 120, 6000, 10 ** 5)
    f = tm.NeighborFinder.from_events(120, src, dst, eidx, ts)
    q = np.arange(5000, 5000 + B)
    sub = f.find_k_hop_device(1, src[q], ts[q], n, eidx[q], seed=3)
    nodes, we, wt, anony, cat = f.find_k_walks_device(n, src[q], N2, sub, seed=4)
    eid = tm.edge_identity_device(we)
    nfeat = rng.standard_normal((120, 32)).astype(np.float32); efeat = rng.standard_normal((6001, 32)).astype(np.float32)
    torch.manual_seed(6)
    m = tm.TempME(_Base(nfeat, efeat), "tgn", "unit", 40, 64, device="cuda", null_model={}).cuda().eval()
    cut = torch.as_tensor(ts[q].astype(np.float32)).cuda()
    plain = m.score_device(nodes, we, wt, cat, cut, eid, group=100).clone()
    grouped = m.score_device(nodes, we, wt, cat, cut, eid, group=100, fanout=N2).clone()
    np.testing.assert_allclose(grouped.cpu().numpy(), plain.cpu().numpy(), rtol=2e-6, atol=0)
    # a hint that does not divide W is ignored
    odd = m.score_device(nodes, we, wt, cat, cut, eid, group=100, fanout=7 if (n * N2) % 7 else 9).clone()
    assert torch.equal(odd, plain)

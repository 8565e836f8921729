"""CPU: pin the oracle (oracle/) to the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Bit-exact for all integer/index outputs; float32 timestamps
are exact copies, so they are compared with == as well."""
import numpy as np
import pytest

import oracle
from oracle import encoder as orc_enc

ROOTS = ("src", "tgt", "bgd")


def graph_of(g, upto=None):
    s = slice(None, upto)
    return oracle.OracleGraph.from_events(int(g["n_nodes"]), g["src"][s].astype(np.int32), g["dst"][s].astype(np.int32),
                                          g["eidx"][s], g["ts"][s].astype(np.float64))


def check_sub(g, prefix, sub):
    for name, rec in zip(("node", "eidx", "ts"), sub):
        for l, a in enumerate(rec):
            ref = g[f"{prefix}_hop{l}_{name}"]
            assert a.shape == ref.shape
            assert (a == ref).all(), f"{prefix} hop{l} {name}"


def check_walks(g, prefix, walks):
    for name, a in zip(("nodes", "eidx", "t", "anony"), walks):
        ref = g[f"{prefix}_w_{name}"]
        assert a.shape == ref.shape
        assert (a == ref).all(), f"{prefix} walks {name}"


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert [hex(x) for x in oracle.philox4x32_10((0, 0), (0, 0, 0, 0))] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in oracle.philox4x32_10((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF,) * 4)] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in oracle.philox4x32_10((0xA4093822, 0x299F31D0), (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344))] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_draw_index_matches_pure_python():
    import refshim
    rng = np.random.default_rng(0)
    for _ in range(200):
        seed, row = int(rng.integers(0, 2**63)), int(rng.integers(0, 2**40))
        stage, slot, L = int(rng.integers(0, 18)), int(rng.integers(0, 100)), int(rng.integers(1, 2**31))
        assert oracle.draw_index(seed, stage, row, slot, L) == refshim.draw_index(seed, stage, row, slot, L)


def test_tie_star(golden):
    g = golden("tie_star")
    og = graph_of(g)
    off, nbr, e, t = og.export()
    assert (off == g["off"]).all() and (nbr == g["nbr"]).all() and (e == g["e"]).all() and (t == g["t"]).all()
    for v, ee, val in g["dict"]:
        assert og.dict_get(v, ee) == val
    # {1:0, 2:1, 3:1, 4:1, 5:4, 6:5} on the hub
    assert [og.dict_get(1, k) for k in range(1, 7)] == [0, 1, 1, 1, 4, 5]
    for v, ee, n in g["fb_eidx"]:
        assert og.find_before_batch([v], None, [ee])[1][0] == n
    for v, tt, n in g["fb_time"]:
        assert og.find_before_batch([int(v)], [tt], None)[1][0] == int(n)


@pytest.mark.parametrize("name", ["rand_small", "rand_bigts", "uslegis"])
def test_csr_hops_walks(golden, name):
    g = golden(name)
    og = graph_of(g)
    if "off" in g:
        off, nbr, e, t = og.export()
        assert (off == g["off"]).all() and (nbr == g["nbr"]).all() and (e == g["e"]).all() and (t == g["t"]).all()
        for v, ee, val in g["dict"]:
            assert og.dict_get(v, ee) == val
    q, n, N2, seed = g["q"], int(g["n"]), int(g["N2"]), int(g["base_seed"])
    ts = g["ts"].astype(np.float64)
    call = 0
    for r in ROOTS:
        roots = {"src": g["src"][q], "tgt": g["dst"][q], "bgd": g["fake"]}[r]
        e = None if r == "bgd" else g["eidx"][q]
        sub = og.find_k_hop(2, roots, ts[q], n, e, seed=seed + call)
        check_sub(g, r, sub)
        walks = og.sample_walks(roots, sub[0][0], sub[1][0], sub[2][0], N2, seed=seed + call + 1)
        check_walks(g, r, walks)
        call += 2
        if f"{r}_edge_identity" in g:
            assert (oracle.edge_identity(walks[1]) == g[f"{r}_edge_identity"]).all()
        if f"{r}_cat" in g:
            cat, _ = oracle.class_ids_prep(walks[3])
            assert (cat == g[f"{r}_cat"]).all()


def test_marginal_frequency(golden):
    g = golden("uslegis")
    tot = np.zeros(12, np.int64)
    for r in ROOTS:
        tot += oracle.class_ids_prep(g[f"{r}_w_anony"])[1]
    freq = tot / tot.sum()      # data_preprocess.py:191-192
    for r in ROOTS:
        assert np.allclose(freq[g[f"{r}_cat"]], g[f"{r}_marginal"], rtol=0, atol=1e-15)


def test_shard_offset_and_wide_fanout(golden):
    g = golden("rand_small")
    og = graph_of(g)
    q, n, N2 = g["q"], int(g["n"]), int(g["N2"])
    ts = g["ts"]
    sub = og.find_k_hop(2, g["src"][q][24:], ts[q][24:], n, g["eidx"][q][24:], seed=77, row_offset=24)
    check_sub(g, "shard1", sub)
    walks = og.sample_walks(g["src"][q][24:], sub[0][0], sub[1][0], sub[2][0], N2, seed=78, row_offset=24)
    check_walks(g, "shard1", walks)
    # ... and a shard reproduces the corresponding rows of the unsharded call (sharding invariance)
    assert (sub[0][1] == g["src_hop1_node"][24:]).all() and (walks[0] == g["src_w_nodes"][24:]).all()
    sub = og.find_k_hop(1, g["dst"][q], ts[q], 40, g["eidx"][q], seed=5)
    check_sub(g, "wide", sub)
    check_walks(g, "wide", og.sample_walks(g["dst"][q], sub[0][0], sub[1][0], sub[2][0], 1, seed=6))
    assert (oracle.edge_identity(g["src_w_eidx"]) == g["src_edge_identity"]).all()


def test_train_finder_missing_eidx(golden):
    g = golden("rand_bigts")
    og = graph_of(g, int(g["n_train"]))
    q, n, N2 = g["q"], int(g["n"]), int(g["N2"])
    assert int(g["train_raises"]) == 1
    with pytest.raises(IndexError):
        og.find_before_batch(g["src"][q][:1], g["ts"][q][:1], g["eidx"][q][:1])
    sub = og.find_k_hop(2, g["src"][q], g["ts"][q], n, None, seed=9 + 6)
    check_sub(g, "train", sub)
    check_walks(g, "train", og.sample_walks(g["src"][q], sub[0][0], sub[1][0], sub[2][0], N2, seed=9 + 7))


def test_null_model_distribution(golden):
    g = golden("nullmodel")
    og = graph_of(g)
    n, seed = int(g["n"]), int(g["base_seed"])
    ti = g["test_idx"]
    src, dst, ts, e = g["src"][ti], g["dst"][ti], g["ts"][ti].astype(np.float64), g["eidx"][ti]
    hist = np.zeros(12, np.int64)
    call = 0
    for k in range(50):                                    # utils/null_model.py:99-118
        s = slice(10 * k, 10 * k + 10)
        subs = []
        for roots, ee in ((src[s], e[s]), (dst[s], e[s]), (g["fakes"][k], None)):
            subs.append(og.find_k_hop(2, roots, ts[s], n, ee, seed=seed + call)); call += 1
        for roots, sub in zip((src[s], dst[s], g["fakes"][k]), subs):
            w = og.sample_walks(roots, sub[0][0], sub[1][0], sub[2][0], 1, seed=seed + call); call += 1
            hist += oracle.class_hist_null(w[3])
    assert hist.sum() == 500 * 3 * n
    assert np.array_equal(hist / (500 * 3 * n), g["dist"])


@pytest.mark.parametrize("tag", ["d172", "d32", "d32_plainattn", "d32_nocat", "d32_hid32"])
def test_encoder_oracle(golden, tag):
    g = golden("encoder_" + tag)
    p = {k[2:]: v for k, v in g.items() if k.startswith("p:")}
    walks = (g["w_nodes"], g["w_eidx"], g["w_t"], g["w_cat"], None)
    out = orc_enc.forward(p, g["node_feat"], g["edge_feat"], walks, g["cut_time"], g["edge_identity"],
                          use_temporal=bool(g["use_temporal"]), if_cat=bool(g["if_cat"]) if "if_cat" in g else True)
    assert out.shape == g["score"].shape and out.dtype == np.float32
    np.testing.assert_allclose(out, g["score"], rtol=1e-5, atol=0)   # the north_star tolerance
    if tag == "d172":
        # float64 arbiter.  Only meaningful while |dt * freq| is small: with ~1e8 timestamps the float32
        # rounding of dt*freq+phase moves the cosine argument by whole radians (that float32 value IS the
        # reference semantics, so the fp32 comparison above is the contract there).
        out64 = orc_enc.forward(p, g["node_feat"], g["edge_feat"], walks, g["cut_time"], g["edge_identity"],
                                dtype=np.float64, use_temporal=bool(g["use_temporal"]))
        np.testing.assert_allclose(out64, g["score"], rtol=1e-5, atol=0)


@pytest.mark.parametrize("tag", ["d32", "d172"])
def test_edge_importance_oracle(golden, tag):
    """oracle.encoder.edge_importance == the reference's retrieve_edge_imp_node (eval mode) on its own scores."""
    from oracle import encoder as enc
    z = golden(f"edgeimp_{tag}")
    p = {k[2:]: z[k] for k in z if k.startswith("p:")}
    sub = ([z["h0_node"], z["h1_node"]], [z["h0_eidx"], z["h1_eidx"]], None)
    walks = (None, z["w_eidx"], z["w_t"], None, None)
    for key, dep in (("dep", True), ("nodep", False)):
        i0, i1 = enc.edge_importance(p, z["edge_feat"], sub, z[f"{key}_score"], walks, use_dependency=dep)
        np.testing.assert_allclose(i0, z[f"{key}_imp0"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(i1, z[f"{key}_imp1"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("tag", ["d32", "d32_hid32"])      # hid32: enhance_main.py's own defaults (--hid_dim 32 --out_dim 32)
def test_enhance_path_oracle(golden, tag):
    """oracle.encoder enhance path == the reference's compute_walk_importance / enhance_predict_walks / enhance_predict_agg (eval)."""
    from oracle import encoder as enc
    z = golden("enhance_" + tag)
    p = {k[2:]: z[k] for k in z if k.startswith("p:")}
    ws = {pre: (z[f"{pre}_nodes"], z[f"{pre}_eidx"], z[f"{pre}_t"], z[f"{pre}_cat"], None) for pre in ("src", "tgt")}
    w = enc.walk_importance(ws["src"][2], ws["src"][0], z["cut_time"], z["node_degree"])
    np.testing.assert_allclose(w, z["w_src"], rtol=1e-5, atol=1e-7)
    emb = {pre: enc.enhance_predict_walks(p, z["node_feat"], z["edge_feat"], ws[pre], z["cut_time"], z[f"{pre}_ei"], z["node_degree"]) for pre in ws}
    np.testing.assert_allclose(emb["src"], z["emb_src"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(emb["tgt"], z["emb_tgt"], rtol=2e-5, atol=2e-5)
    pos = enc.affinity_score(p, np.concatenate([emb["src"], z["src_gat"]], -1), np.concatenate([emb["tgt"], z["tgt_gat"]], -1))
    neg = enc.affinity_score(p, np.concatenate([emb["src"], z["src_gat"]], -1), np.concatenate([emb["src"], z["bgd_gat"]], -1))
    np.testing.assert_allclose(pos, z["pos"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(neg, z["neg"], rtol=1e-4, atol=1e-4)


def test_kl_loss_oracle(golden):
    """oracle.encoder.kl_loss == the reference's TempME.kl_loss for both priors (absent classes, clamped 0 / 1 scores, one root)."""
    from oracle import encoder as enc
    z = golden("kl_loss")
    for name in ("us", "few", "one"):
        for prior in ("empirical", "uniform"):
            for target in (0.3, 0.05):
                v = enc.kl_loss(z[f"{name}_prob"], z[f"{name}_cat"], z["null_values"], target, prior)
                np.testing.assert_allclose(v, float(z[f"{name}_{prior}_{target}"]), rtol=1e-5, atol=1e-7)

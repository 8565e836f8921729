"""CPU, build container only: differential test of the oracle against the LIVE reference
(imported from /root/reference, randomness replaced by the draw contract via tests/refshim.py)
on freshly generated random multigraphs.  Skipped where the reference is absent (the GPU box)."""
import os

import numpy as np
import pytest

import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def rg():
    return refshim.import_reference()


@pytest.mark.parametrize("seed,N,E,tmax,n,N2,lo", [
    (11, 12, 150, 8, 5, 2, 0),       # tiny, dense ties, node 0 is a real node, self-loops
    (12, 80, 1500, 400, 9, 3, 1),    # sparser ties, ids start at 1 (node 0 = padding only)
    (13, 30, 900, 3, 33, 1, 0),      # three distinct timestamps, fan-out above one warp
    (14, 200, 600, 10**6, 4, 4, 1),  # almost no ties, many empty prefixes
])
def test_random_graph_matches_reference(rg, seed, N, E, tmax, n, N2, lo):
    import oracle
    rng = np.random.default_rng(seed)
    src = rng.integers(lo, N, E); dst = rng.integers(lo, N, E)
    ts = np.sort(rng.integers(0, tmax, E)).astype(np.float64)
    eidx = rng.permutation(E) + 1
    nf = rg.NeighborFinder(refshim.adj_list_from_events(N, src, dst, eidx, ts))
    og = oracle.OracleGraph.from_events(N, src, dst, eidx, ts)
    off, nbr, e, t = og.export()
    assert (off == nf.off_set_l).all() and (nbr == nf.node_idx_l).all() and (e == nf.edge_idx_l).all() and (t == nf.node_ts_l).all()
    for v in range(N):
        for ee, val in nf.nodeedge2idx[v].items():
            assert og.dict_get(v, ee) == val
    q = rng.choice(np.arange(E // 2, E), 20, replace=False)
    fake = rng.integers(lo, N, len(q))
    shim = refshim.DrawShim(base_seed=seed * 1000)
    call = 0
    with shim.patched(rg):
        for roots, ee in ((src[q], eidx[q]), (dst[q], eidx[q]), (fake, None)):
            sub = nf.find_k_hop(2, roots, ts[q], n, e_idx_l=ee)
            walks = nf.find_k_walks(n, roots, N2, sub)
            osub = og.find_k_hop(2, roots, ts[q], n, ee, seed=seed * 1000 + call)
            owalks = og.sample_walks(roots, osub[0][0], osub[1][0], osub[2][0], N2, seed=seed * 1000 + call + 1)
            call += 2
            for a, b in zip(sub, osub):
                for x, y in zip(a, b):
                    assert x.dtype == y.dtype and (x == y).all()
            for x, y in zip(walks, owalks):
                assert x.shape == y.shape and (x == y).all()
            assert walks[2].dtype == owalks[2].dtype == np.float32


def test_pack_container_and_loader_match_the_reference_loader(tmp_path):
    """A pack written by tempme_b200.h5min and opened with its reader, pushed through the REFERENCE's own
    utils/batch_loader.load_subgraph_margin / get_item, equals compat.load_subgraph_margin on the same file."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("ref_batch_loader", os.path.join(refshim.REF, "utils", "batch_loader.py"))
    bl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bl)
    from tempme_b200 import compat, h5min
    rng = np.random.default_rng(0)
    n, Q, W = 5, 9, 15
    arrays = {f"subgraph_{r}_{l}": rng.integers(0, 50, (Q, 3 * n ** (l + 1))).astype(np.float64) for r in ("src", "tgt", "bgd") for l in (0, 1)}
    arrays.update({f"walks_{r}_new": rng.integers(0, 9, (Q, W, 14)).astype(np.float64) for r in ("src", "tgt", "bgd")})
    arrays["dst_fake"] = rng.integers(1, 50, Q).astype(np.int64)
    path = str(tmp_path / "unit_test_cat.h5")
    h5min.write(path, dict(arrays))
    f = h5min.read(path)
    args = types.SimpleNamespace(n_degree=n)
    ref = bl.load_subgraph_margin(args, f)
    ours = compat.load_subgraph_margin(args, f)

    def same(a, b):
        if isinstance(a, (tuple, list)):
            assert len(a) == len(b)
            for x, y in zip(a, b):
                same(x, y)
        else:
            assert a.dtype == b.dtype and a.shape == b.shape and (a == b).all()
    same(ref, ours)
    same(bl.get_item(ref, np.arange(2, 6)), bl.get_item(ours, np.arange(2, 6)))

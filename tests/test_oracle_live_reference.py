"""CPU, build container only: differential test of the oracle against the LIVE reference
(imported from /root/reference, randomness replaced by the draw contract via tests/refshim.py)
on freshly generated random multigraphs.  Skipped where the reference is absent (the GPU box)."""
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

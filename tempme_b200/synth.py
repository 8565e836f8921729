"""Synthetic temporal graphs of the shapes BASELINE.json names.  All draws come
from numpy.random.Generator(PCG64(20240 + cfg)); there is no network for the real datasets."""
from __future__ import annotations

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# cfg -> (n first-hop fan-out, N2 second-step fan-out, node/time dim D, edge dim Ed)
SHAPES = {
    "cfg1": dict(n=30, N2=3, D=172, Ed=1, desc="bundled ml_uslegis_sampled (224 nodes, 8,832 events), TGAT base"),
    "cfg2": dict(n=30, N2=1, D=32, Ed=32, desc="synthetic Enron-shaped (184 nodes, 125,235 events), TGN base, 30 walks/query"),
    "cfg3": dict(n=20, N2=3, D=172, Ed=172, desc="synthetic Wikipedia-shaped bipartite (9,227 nodes, 157,474 events), GraphMixer base"),
    "cfg4": dict(n=20, N2=5, D=172, Ed=172, desc="synthetic Reddit-shaped (10,984 nodes, 672,447 events), 100 walks/query"),
    "cfg5": dict(n=20, N2=3, D=32, Ed=32, desc="synthetic power-law (1M nodes, 100M events)"),
}


def _zipf_choice(rng, ids, size, a):
    p = 1.0 / np.arange(1, len(ids) + 1, dtype=np.float64) ** a
    p /= p.sum()
    return ids[rng.choice(len(ids), size=size, p=p)]


def make_graph(cfg: str, scale: float = 1.0):
    """Returns dict(n_nodes, src, dst, eidx, ts) with events in chronological (CSV) order, eidx = 1..E."""
    rng = np.random.Generator(np.random.PCG64(20240 + int(cfg[-1])))
    if cfg == "cfg1":
        g = np.load(os.path.join(HERE, "..", "tests", "golden", "uslegis.npz"))
        return dict(n_nodes=int(g["n_nodes"]), src=g["src"].astype(np.int64), dst=g["dst"].astype(np.int64),
                    eidx=g["eidx"].astype(np.int64), ts=g["ts"].astype(np.float64))
    if cfg == "cfg2":
        N, E = 184, int(125_235 * scale)
        ids = np.arange(1, N + 1)
        src = _zipf_choice(rng, ids, E, 1.0); dst = _zipf_choice(rng, ids, E, 1.0)
        loop = src == dst
        dst[loop] = (dst[loop] % N) + 1                       # no self loops
        ts = np.sort(rng.integers(100_000_000, 110_000_000, E)).astype(np.float64)
        dup = rng.random(E) < 0.3                              # ~30 % duplicate timestamps
        ts[1:][dup[1:]] = ts[:-1][dup[1:]]
        ts = np.maximum.accumulate(ts)
        n_nodes = N + 1
    elif cfg in ("cfg3", "cfg4"):
        users, items, E, zu, zi = (8227, 1000, 157_474, 0.8, 1.1) if cfg == "cfg3" else (10_000, 984, 672_447, 0.8, 1.1)
        E = int(E * scale)
        src = _zipf_choice(rng, np.arange(1, users + 1), E, zu)
        dst = _zipf_choice(rng, np.arange(users + 1, users + items + 1), E, zi)
        ts = np.sort(rng.integers(0, 2_680_000, E)).astype(np.float64)
        n_nodes = users + items + 1
    elif cfg == "cfg5":
        N, E = int(1_000_000 * scale), int(100_000_000 * scale)
        # configuration-model endpoints with power-law weights w_r ~ r^-0.8 (degree exponent ~2.25), drawn by
        # inverse CDF; the largest hub receives ~1 % of the 2E endpoints (~2M entries at full scale)
        def endpoints(size):
            s_, top = 0.8, float(N + 1) ** 0.2
            x = (1.0 + rng.random(size) * (top - 1.0)) ** (1.0 / (1.0 - s_))
            return np.minimum(x.astype(np.int64), N)
        src = endpoints(E)
        dst = endpoints(E)
        loop = src == dst
        dst[loop] = (dst[loop] % N) + 1
        ts = np.sort(rng.random(E) * 1e9)
        dup = rng.random(E) < 0.05
        ts[1:][dup[1:]] = ts[:-1][dup[1:]]
        ts = np.maximum.accumulate(ts)
        n_nodes = N + 1
    else:
        raise ValueError(cfg)
    return dict(n_nodes=n_nodes, src=src.astype(np.int64), dst=dst.astype(np.int64),
                eidx=np.arange(1, len(src) + 1, dtype=np.int64), ts=ts)


def make_features(cfg: str, n_nodes: int, n_events: int, device=None):
    """Base-model feature tables (row 0 = padding): edge feats N(0,1); node feats N(0,1), except the
    Wikipedia/Reddit-shaped configs whose real node features are all-zero."""
    import torch
    sh = SHAPES[cfg]
    gen = torch.Generator(device="cpu").manual_seed(20240 + int(cfg[-1]))
    big = n_events * sh["Ed"] > 200_000_000
    if big and device is not None:     # cfg5: 12.8 GB -- generate on the device
        dgen = torch.Generator(device=device).manual_seed(20240 + int(cfg[-1]))
        efeat = torch.randn((n_events + 1, sh["Ed"]), generator=dgen, device=device)
    else:
        efeat = torch.randn((n_events + 1, sh["Ed"]), generator=gen)
    nfeat = torch.zeros((n_nodes, sh["D"])) if cfg in ("cfg3", "cfg4") else torch.randn((n_nodes, sh["D"]), generator=gen)
    efeat[0] = 0
    nfeat[0] = 0
    return nfeat, efeat


def make_queries(graph, rng, count):
    """Query events drawn from the test split (ts > 85th percentile), in chronological order, with a random
    fake destination per event (RandEdgeSampler semantics, utils/batch_loader.py:39-42).  The split and the destination
    pool are computed once per graph (cached in the dict)."""
    ts = graph["ts"]
    if "_pool" not in graph:
        graph["_pool"] = np.nonzero(ts > np.quantile(ts, 0.85))[0]
        graph["_dst_pool"] = np.flatnonzero(np.bincount(graph["dst"], minlength=int(graph["n_nodes"])))      # == np.unique(dst)
    pool, dst_pool = graph["_pool"], graph["_dst_pool"]
    q = np.sort(rng.choice(pool, size=count, replace=count > len(pool)))
    fake = dst_pool[rng.integers(0, len(dst_pool), count)]
    return graph["src"][q], graph["dst"][q], fake, ts[q], graph["eidx"][q]


def share_graph(cfg: str, scale: float, local_rank: int, world: int, barrier, tag: str = ""):
    """One graph per NODE instead of one per rank: local rank 0 generates it and writes the event arrays to /dev/shm, the other
    ranks map them read-only after `barrier()`.  Returns (graph, cleanup) -- call cleanup() on every rank after a later barrier."""
    if world <= 1:
        return make_graph(cfg, scale), (lambda: None)
    base = os.path.join("/dev/shm", f"tempme_b200_{cfg}_{scale}_{tag}")
    names = ("src", "dst", "ts")
    if local_rank == 0:
        g = make_graph(cfg, scale)
        try:
            for k in names:
                np.save(f"{base}_{k}.npy", g[k].astype(np.int32) if k != "ts" else g[k])
            with open(base + "_meta", "w") as fh:
                fh.write(str(int(g["n_nodes"])))
        except OSError:
            pass
    barrier()
    if local_rank != 0:
        try:
            g = dict(n_nodes=int(open(base + "_meta").read()))
            for k in names:
                g[k] = np.load(f"{base}_{k}.npy", mmap_mode="r")
            g["eidx"] = np.arange(1, len(g["src"]) + 1, dtype=np.int64)
        except OSError:
            g = make_graph(cfg, scale)          # no shared memory file system: every rank generates (same seed, same graph)

    def cleanup():
        if local_rank == 0:
            for k in names:
                try:
                    os.unlink(f"{base}_{k}.npy")
                except OSError:
                    pass
            try:
                os.unlink(base + "_meta")
            except OSError:
                pass
    return g, cleanup

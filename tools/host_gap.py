#!/usr/bin/env python
"""tools/host_gap.py [--workload cfg5]: host-side and device-side time of every call of one bench step (synchronised around each call),
to find host stalls between kernels."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tempme_b200 as tm
from tempme_b200 import synth
from tempme_b200.graph import edge_identity_device
from bench import random_params, default_events

ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="cfg5"); ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
sh = synth.SHAPES[args.workload]; n, N2, D, Ed = sh["n"], sh["N2"], sh["D"], sh["Ed"]
Q = default_events(args.workload) // 100 * 100
graph = synth.make_graph(args.workload, args.scale)
finder = tm.NeighborFinder.from_events(graph["n_nodes"], graph["src"], graph["dst"], graph["eidx"], graph["ts"], device=dev, seed=1234)
nfeat, efeat = synth.make_features(args.workload, graph["n_nodes"], len(graph["src"]), device=dev)

class Base:
    n_feat_th = nfeat.to(dev); e_feat_th = efeat.to(dev)
    node_raw_features = torch.nn.Embedding.from_pretrained(n_feat_th, padding_idx=0, freeze=True)
    edge_raw_features = torch.nn.Embedding.from_pretrained(e_feat_th, padding_idx=0, freeze=True)

model = tm.TempME(Base(), "tgn", args.workload, 40, 64, device=dev, null_model={}).to(dev).eval()
pipe = tm.MotifPipeline(finder, model, n, N2, group=100, seed=99)
rng = np.random.default_rng(1000)
qs = [pipe.stage_queries(*synth.make_queries(graph, rng, Q)) for _ in range(4)]
f = finder

def timed(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"  {name:28s} host {1e3 * (t1 - t0):8.3f} ms   until done {1e3 * (t2 - t0):8.3f} ms", flush=True)
    return r

for it, (roots, e, cut64) in enumerate(qs):
    print("step", it)
    h1 = timed("sample_hop", lambda: f.sample_hop_device(roots, cut64, n, e, seed=1, stage=0, row_offset=0))
    w = timed("find_k_walks", lambda: f.find_k_walks_device(n, roots, N2, ([h1[0]], [h1[1]], [h1[2]]), seed=2, row_offset=0, want_anony=False, want_cat=True,
                                                            hist_null=pipe.hist_null, hist_prep=pipe.hist_prep, scanned=pipe.scanned))
    nodes, eidx, t, _, cat = w
    eid = timed("edge_identity", lambda: edge_identity_device(eidx))
    c32 = timed("cut.to(float32)", lambda: cut64.to(torch.float32))
    timed("packed_weights", lambda: model.packed_weights())
    timed("_tables", lambda: model._tables())
    timed("score_device", lambda: model.score_device(nodes, eidx, t, cat, c32, eid, group=100))

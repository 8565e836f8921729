#!/usr/bin/env python
"""Measurement of SURVEY 8(f) row f1 on one B200: hop-2 subgraph sampling + retrieve_edge_imp_node (eval) for the
roots of one bench step (cfg2 by default).  Prints one JSON line: roots/s and motifs/s through the aggregation, the
gate kernel's tensor roofline and the aggregator's HBM roofline.  CUDA events on the launching stream, L2 flushed
between timed iterations, 3 warm-up iterations.

  python tools/bench_edge_imp.py [--workload cfg2] [--events 4000] [--steps 10]
"""
import argparse, ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tempme_b200 as tm
from tempme_b200 import synth
from bench import random_params, peaks

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2"); ap.add_argument("--events", type=int, default=4000)
ap.add_argument("--steps", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
sh = synth.SHAPES[args.workload]; n, N2, D, Ed = sh["n"], sh["N2"], sh["D"], sh["Ed"]; W = n * N2
graph = synth.make_graph(args.workload, 1.0)
finder = tm.NeighborFinder.from_events(graph["n_nodes"], graph["src"], graph["dst"], graph["eidx"], graph["ts"], device=dev, seed=1)
nfeat, efeat = synth.make_features(args.workload, graph["n_nodes"], len(graph["src"]), device=dev)

class Base:
    n_feat_th = nfeat.to(dev); e_feat_th = efeat.to(dev)
    node_raw_features = torch.nn.Embedding.from_pretrained(n_feat_th, padding_idx=0, freeze=True)
    edge_raw_features = torch.nn.Embedding.from_pretrained(e_feat_th, padding_idx=0, freeze=True)

model = tm.TempME(Base(), "tgn", args.workload, 40, 64, device=dev, null_model={}).to(dev).eval()
model.load_state_dict({k: torch.as_tensor(v) for k, v in random_params(D, Ed).items()}, strict=False)
pipe = tm.MotifPipeline(finder, model, n, N2, group=100, seed=5)
rng = np.random.default_rng(3)
Q = args.events // 100 * 100
roots, e, cut64, _ = pipe.stage_queries(*synth.make_queries(graph, rng, Q))
R = roots.numel()
scores, (nodes, eidx, t, cat, eid) = pipe.run_device(roots, e, cut64, want_walks=True)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ms = {"hop2": 0.0, "edge_imp": 0.0, "aggregate_only": 0.0}
nodep = tm.TempME(Base(), "tgn", args.workload, 40, 64, device=dev, null_model={}, use_dependency_aware_sampling=False).to(dev).eval()
for it in range(args.warmup + args.steps):
    flush.zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    sub = finder.find_k_hop_device(2, roots, None, n, e, seed=9)          # hop-1 + hop-2 slots (graph.py:233-262)
    ev[1].record()
    imp0, imp1 = model.edge_importance_device(scores, eidx, t, sub[0][0].view(R, n), sub[1][0].view(R, n), sub[0][1].view(R, n * n), sub[1][1].view(R, n * n))
    ev[2].record()
    nodep.edge_importance_device(scores, eidx, t, sub[0][0].view(R, n), sub[1][0].view(R, n), sub[0][1].view(R, n * n), sub[1][1].view(R, n * n))   # no gate: segmented max only
    ev[3].record(); ev[3].synchronize()
    if it >= args.warmup:
        ms["hop2"] += ev[0].elapsed_time(ev[1]); ms["edge_imp"] += ev[1].elapsed_time(ev[2]); ms["aggregate_only"] += ev[2].elapsed_time(ev[3])
ms = {k: v / args.steps for k, v in ms.items()}
M = R * W
gate_flops = M * 3 * 2.0 * ((Ed + D) * 64 + 64 * 32 + 32)
agg_bytes = R * (3 * W * 8.0 + (n + n * n) * 12.0)
pk = peaks()
print(json.dumps({"metric": "motif_to_edge_aggregation_motifs_per_sec", "value": M / (ms["edge_imp"] * 1e-3), "unit": "motifs/s",
                  "config": {"workload": args.workload, "roots": R, "walks_per_root": W, "hop_slots_per_root": n + n * n},
                  "ms": ms, "gate_algorithmic_tflops_if_whole_stage": gate_flops / (ms["edge_imp"] * 1e-3) / 1e12,
                  "aggregator_algorithmic_gbs_if_whole_stage": agg_bytes / (ms["edge_imp"] * 1e-3) / 1e9, "peaks": pk,
                  "launches": tm.launch_count(), "imp_mean": [float(imp0.mean()), float(imp1.mean())]}))

#!/usr/bin/env python
"""Measurement of SURVEY 8(f) row f3 on one B200: TempME.enhance_predict_walks (eval) for the roots of one bench step (cfg2 by default).
One JSON line: roots/s and motifs/s, stage times (scorer with hidden-vector output, walk importance, weighted reduction).
CUDA events on the launching stream, L2 flushed between timed iterations, 3 warm-up iterations."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tempme_b200 as tm
from tempme_b200 import synth
from bench import random_params

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2"); ap.add_argument("--events", type=int, default=16000)
ap.add_argument("--steps", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--hid", type=int, default=64, help="hid_dim: 64 (temp_exp_main.py) or 32 (enhance_main.py's default)")
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

model = tm.TempME(Base(), "tgn", args.workload, 40, args.hid, device=dev, null_model={}, batch_group=100).to(dev).eval()
if args.hid == 64:
    model.load_state_dict({k: torch.as_tensor(v) for k, v in random_params(D, Ed).items()}, strict=False)
model.node_degree = torch.as_tensor(np.bincount(np.concatenate([graph["src"], graph["dst"]]), minlength=graph["n_nodes"]).astype(np.float32)).to(dev)
pipe = tm.MotifPipeline(finder, model, n, N2, group=100, seed=5)
Q = args.events // 100 * 100
roots, e, cut64, _ = pipe.stage_queries(*synth.make_queries(graph, np.random.default_rng(3), Q))
R = roots.numel()
_, (nodes, eidx, t, cat, eid) = pipe.run_device(roots, e, cut64, want_walks=True)
cut = cut64.to(torch.float32)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ms = 0.0
for it in range(args.warmup + args.steps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    emb = model.enhance_predict_walks((nodes, eidx, t, cat.view(R, W, 1), None), cut, eid)
    b.record(); b.synchronize()
    if it >= args.warmup:
        ms += a.elapsed_time(b)
ms /= args.steps
print(json.dumps({"metric": "enhance_predict_walks_motifs_per_sec", "value": R * W / (ms * 1e-3), "unit": "motifs/s", "roots_per_sec": R / (ms * 1e-3),
                  "config": {"workload": args.workload, "roots": R, "walks_per_root": W, "hid_dim": args.hid}, "ms": ms, "launches": tm.launch_count(),
                  "emb_mean_abs": float(emb.abs().mean())}))

"""Drop-in wiring for the unmodified reference tree: the reference imports its graph / null-model code as
``utils`` and ``processed.utils`` (utils/__init__.py:1-3, processed/data_preprocess.py:16) and constructs the explainer by name.
``install()`` registers modules of those names backed by tempme_b200, so that ``from utils import NeighborFinder,
get_null_distribution, RandEdgeSampler`` (models/explainer.py:8, temp_exp_main.py) resolves to the B200 implementation."""
from __future__ import annotations

import sys
import types


def install(also_models: bool = False):
    """Registers ``utils`` / ``processed.utils`` (and with ``also_models`` a ``models`` module exposing TempME) in ``sys.modules``.
    Call before the reference's drivers import anything.  Returns the ``utils`` module."""
    import tempme_b200 as tm
    from tempme_b200 import null_model as nm

    u = types.ModuleType("utils")
    u.__doc__ = "tempme_b200 stand-in for the reference's utils package"
    for name in ("NeighborFinder", "get_null_distribution", "RandEdgeSampler"):
        setattr(u, name, getattr(tm, name))
    for name in ("load_data_shuffle", "pre_processing", "statistic", "degree_dict"):
        setattr(u, name, getattr(nm, name))
    u.load_subgraph_margin = load_subgraph_margin
    g = types.ModuleType("utils.graph"); g.NeighborFinder = tm.NeighborFinder
    n = types.ModuleType("utils.null_model")
    for name in ("get_null_distribution", "load_data_shuffle", "pre_processing", "statistic", "degree_dict"):
        setattr(n, name, getattr(nm, name))
    u.graph, u.null_model = g, n
    p = sys.modules.get("processed") or types.ModuleType("processed")
    p.utils = u
    sys.modules.update({"utils": u, "utils.graph": g, "utils.null_model": n, "processed": p, "processed.utils": u})
    if also_models:
        m = types.ModuleType("models")
        m.TempME = tm.TempME
        sys.modules["models"] = m
    return u


def load_subgraph_margin(args, file):
    """utils/batch_loader.py:120-201 for a pack opened with tempme_b200.load_pack or h5py: the 7-tuple
    (subgraph_src, subgraph_tgt, subgraph_bgd, walks_src, walks_tgt, walks_bgd, dst_fake) the drivers index with get_item."""
    n = args.n_degree

    def subgraph(root):
        recs = ([], [], [])
        for l, k in enumerate((n, n * n)):
            a = file[f"subgraph_{root}_{l}"][:]
            for j, r in enumerate(recs):
                r.append(a[:, j * k:(j + 1) * k])
        return recs

    def walks(root):
        w = file[f"walks_{root}_new"][:]
        return (w[:, :, :6].astype(int), w[:, :, 6:9].astype(int), w[:, :, 9:12], w[:, :, 12:13].astype(int), w[:, :, 13:14])

    return (subgraph("src"), subgraph("tgt"), subgraph("bgd"), walks("src"), walks("tgt"), walks("bgd"), file["dst_fake"][:])

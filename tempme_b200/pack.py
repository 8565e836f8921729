"""Offline explanation pack (reference processed/data_preprocess.py:99-145, 148-214, 345-419) built on the GPU.

The reference loops over the query events one by one (six sampler calls per event) and writes ``{data}_{MODE}.h5``, then
``{data}_{MODE}_cat.h5`` (walks with class id + dataset-wide class frequency) and ``{data}_{MODE}_edge.npy``.  ``build_pack`` makes the
same arrays with one batched call per root type; ``utils/batch_loader.load_subgraph_margin(args, file)`` (:120-201) only does
``file[name][:]``, so the returned dict can be passed to it directly.  ``save_pack`` writes the HDF5 container the reference opens
(through h5py when importable, else through the spec-level writer tempme_b200.h5min).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .graph import edge_identity_device

PACK_KEYS = ["subgraph_src_0", "subgraph_src_1", "subgraph_tgt_0", "subgraph_tgt_1", "subgraph_bgd_0", "subgraph_bgd_1",
             "walks_src_new", "walks_tgt_new", "walks_bgd_new", "dst_fake"]


def build_pack(finder, src, dst, ts, e_idx, dst_fake, n_degree, num_neighbors=3, seed=None):
    """-> (pack dict with PACK_KEYS, edge_load [3, Q, W, 3, 3] float64).

    src/dst/ts/e_idx: the query events (the reference passes all but the last event of the split, data_preprocess.py:106);
    dst_fake: the background roots (``RandEdgeSampler.sample``); source and target roots look their window up by e_idx, background
    roots by time (:114-122).  subgraph_*_l = [node | eidx | t] of hop l (:115-116); walks_*_new = [nodes 6 | eidx 3 | t 3 | class id |
    class frequency over all three walk sets] (:203-208); edge_load = new_edge_info of the three (:345-357)."""
    src, dst, dst_fake = np.asarray(src), np.asarray(dst), np.asarray(dst_fake)
    ts = np.asarray(ts, np.float64)
    Q, n, N2 = len(src), int(n_degree), int(num_neighbors)
    W = n * N2
    dev = finder.device
    hist = torch.zeros(12, dtype=torch.int64, device=dev)
    pack, per_root, edges = {}, {}, []
    for k, (name, roots, e) in enumerate((("src", src, e_idx), ("tgt", dst, e_idx), ("bgd", dst_fake, None))):
        # six top-level sampler calls, as the reference makes per event; with an explicit seed call c uses seed + c
        sub = finder.find_k_hop_device(2, roots, ts, n, e, seed=None if seed is None else seed + 2 * k)
        nodes, eidx, t, _, cat = finder.find_k_walks_device(n, roots, N2, sub, seed=None if seed is None else seed + 2 * k + 1,
                                                            want_anony=False, want_cat=True, hist_prep=hist)
        for l in range(2):
            pack[f"subgraph_{name}_{l}"] = np.concatenate([sub[0][l].cpu().numpy(), sub[1][l].cpu().numpy(), sub[2][l].cpu().numpy()], axis=-1).astype(np.float64)
        per_root[name] = (nodes.cpu().numpy(), eidx.cpu().numpy(), t.cpu().numpy(), cat.cpu().numpy())
        edges.append(edge_identity_device(eidx).cpu().numpy().astype(np.float64))
    freq = hist.cpu().numpy().astype(np.float64) / (Q * W * 3)               # :190-192
    for name, (nodes, eidx, t, cat) in per_root.items():
        pack[f"walks_{name}_new"] = np.concatenate([nodes.astype(np.float64), eidx.astype(np.float64), t.astype(np.float64),
                                                    cat[..., None].astype(np.float64), freq[cat][..., None]], axis=-1)
    pack["dst_fake"] = dst_fake
    return pack, np.stack(edges, axis=0)


def save_pack(pack, edge_load, directory, data, mode):
    """Writes ``{data}_{mode}_cat.h5`` and ``{data}_{mode}_edge.npy`` (the two files temp_exp_main.py:705-714 opens); returns the paths.
    The HDF5 file comes from h5py when it is importable, otherwise from tempme_b200.h5min (same container subset: version-0 superblock,
    one old-style root group, contiguous datasets), so a loadable ``_cat.h5`` is produced either way."""
    os.makedirs(directory, exist_ok=True)
    edge_path = os.path.join(directory, f"{data}_{mode}_edge.npy")
    np.save(edge_path, edge_load)
    path = os.path.join(directory, f"{data}_{mode}_cat.h5")
    try:
        import h5py
    except ImportError:
        from . import h5min
        h5min.write(path, {k: np.asarray(pack[k]) for k in PACK_KEYS})
        return path, edge_path
    with h5py.File(path, "w") as hf:
        for k in PACK_KEYS:
            hf.create_dataset(k, data=pack[k])
    return path, edge_path


def load_pack(path):
    """``{name: array}`` of a ``_cat.h5`` pack: what ``h5py.File(path)`` gives utils/batch_loader.load_subgraph_margin (:120-201), which only
    does ``file[name][:]``.  Uses h5py when importable, else the spec-following reader of tempme_b200.h5min."""
    try:
        import h5py
    except ImportError:
        from . import h5min
        return h5min.read(path)
    with h5py.File(path, "r") as hf:
        return {k: hf[k][:] for k in hf.keys()}

"""Null-model motif distribution with the reference's API (reference utils/null_model.py:10-128):
endpoint-shuffled graph -> 50 x 10 queries x 3 roots x ``degree`` walks -> 12-class frequencies.
Sampling, anonymisation and the histogram run on the GPU (tm_sample_hop / tm_sample_walks with the
fused class histogram); only the CSV load and the permutation are host work, as in the reference.
"""
from __future__ import annotations

import os
import os.path as osp
import random

import numpy as np
import torch

from .graph import NeighborFinder

degree_dict = {"wikipedia": 20, "reddit": 20, "uci": 30, "mooc": 60, "enron": 30, "enron_sampled": 30,
               "canparl": 30, "uslegis": 30, "uslegis_sampled": 30}


class RandEdgeSampler(object):
    """utils/batch_loader.py:32-42."""

    def __init__(self, src_list, dst_list):
        self.src_list = np.unique(np.concatenate(src_list))
        self.dst_list = np.unique(np.concatenate(dst_list))

    def sample(self, size):
        src_index = np.random.randint(0, len(self.src_list), size)
        dst_index = np.random.randint(0, len(self.dst_list), size)
        return self.src_list[src_index], self.dst_list[dst_index]


def _find_csv(data):
    name = 'ml_{}.csv'.format(data)
    roots = [os.environ.get("TEMPME_DATA_ROOT"), osp.join(os.getcwd(), "processed"), os.getcwd(),
             osp.join(osp.dirname(osp.realpath(__file__)), '..', 'processed')]
    for r in roots:
        if r and osp.exists(osp.join(r, name)):
            return osp.join(r, name)
    raise FileNotFoundError(f"{name} not found (set TEMPME_DATA_ROOT to the directory that holds it)")


def load_data_shuffle(mode, data):
    """utils/null_model.py:13-72: permute (src, dst, label) against (ts, e_idx), then the usual split."""
    import pandas as pd
    g_df = pd.read_csv(_find_csv(data))
    val_time, test_time = list(np.quantile(g_df.ts, [0.70, 0.85]))
    src_l, dst_l, e_idx_l, label_l, ts_l = g_df.u.values, g_df.i.values, g_df.idx.values, g_df.label.values, g_df.ts.values
    permutation = np.random.permutation(len(ts_l))
    src_l = np.array(src_l)[permutation]; dst_l = np.array(dst_l)[permutation]; label_l = np.array(label_l)[permutation]
    max_idx = max(src_l.max(), dst_l.max())
    random.seed(2023)
    total_node_set = set(np.unique(np.hstack([g_df.u.values, g_df.i.values])))
    temp_val = list(set(src_l[ts_l > val_time]).union(set(dst_l[ts_l > val_time])))
    mask_node_set = set(random.sample(temp_val, int(0.1 * len(total_node_set))))
    mask_src_flag = g_df.u.map(lambda x: x in mask_node_set).values
    mask_dst_flag = g_df.i.map(lambda x: x in mask_node_set).values
    none_node_flag = (1 - mask_src_flag) * (1 - mask_dst_flag)
    train = (ts_l <= val_time) * (none_node_flag > 0)
    val = (ts_l <= test_time) * (ts_l > val_time)
    test = ts_l > test_time
    train_rand_sampler = RandEdgeSampler((src_l[train],), (dst_l[train],))
    test_rand_sampler = RandEdgeSampler((src_l[train], src_l[val], src_l[test]), (dst_l[train], dst_l[val], dst_l[test]))
    if mode == "test":
        finder = NeighborFinder.from_events(max_idx + 1, src_l, dst_l, e_idx_l, ts_l)
        return test_rand_sampler, src_l[test], dst_l[test], ts_l[test], label_l[test], e_idx_l[test], finder
    finder = NeighborFinder.from_events(max_idx + 1, src_l[train], dst_l[train], e_idx_l[train], ts_l[train])
    return train_rand_sampler, src_l[train], dst_l[train], ts_l[train], label_l[train], e_idx_l[train], finder


def statistic(out_anony, sat, strint_rep=None):
    """utils/null_model.py:75-82 -- counts of the 12 anonymised classes, keys 1..12, on the GPU."""
    from .graph import class_hist_device
    a = torch.as_tensor(np.ascontiguousarray(out_anony)).to(torch.int32).cuda()
    hn, _, _, err = class_hist_device(a, want_cat=False)
    if int(err.item()):
        raise KeyError("anonymised row is not one of the 12 motif classes")
    for k, v in enumerate(hn.cpu().tolist()):
        sat[k + 1] = sat.get(k + 1, 0) + v
    return sat


def pre_processing(ngh_finder, sampler, src, dst, ts, val_e_idx_l, num_neighbors, fakes=None):
    """utils/null_model.py:86-121.  ``fakes`` (optional [50, 10]) replaces sampler.sample for replay."""
    degree = num_neighbors
    batch_size = 10
    total_sample = 50 * batch_size
    hist = torch.zeros(12, dtype=torch.int64, device=ngh_finder.device)
    for k in range(50):
        s = slice(k * batch_size, (k + 1) * batch_size)
        src_c, dst_c, ts_c = src[s], dst[s], ts[s]
        e_c = val_e_idx_l[s] if (val_e_idx_l is not None) else None
        if len(src_c) == 0:
            continue
        fake = fakes[k] if fakes is not None else sampler.sample(len(src_c))[1]
        # the reference samples 2 hops here but find_k_walks only reads hop 0 (graph.py:281)
        subs = [ngh_finder.find_k_hop_device(1, r, ts_c, degree, e_idx_l=e) for r, e in ((src_c, e_c), (dst_c, e_c), (fake, None))]
        for r, sub in zip((src_c, dst_c, fake), subs):
            ngh_finder.find_k_walks_device(degree, r, 1, sub, want_anony=False, want_cat=False, hist_null=hist)
    ngh_finder._raise_if_err("null model pre_processing")
    counts = hist.cpu().numpy()
    return {k + 1: counts[k] / (total_sample * 3 * degree) for k in range(12)}


def get_null_distribution(data_name):
    """utils/null_model.py:124-128."""
    num_neighbors = degree_dict[data_name]
    rand_sampler, src, dst, ts, _, e_idx, finder = load_data_shuffle(mode="test", data=data_name)
    return pre_processing(finder, rand_sampler, src, dst, ts, e_idx, num_neighbors)

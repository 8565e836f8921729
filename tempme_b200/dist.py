"""Query sharding over the GPUs of one node (one process per GPU, torch.distributed).

The path shards by *whole reference batches* (the temporal attention normalises by a batch-global std,
explainer.py:826-828); graph, feature tables and weights are replicated.  Draws are keyed by the global root
row, so every split produces the same walks and scores as one GPU.  The only exchanges are the 12-bin class
histogram all-reduce and the score gather (NCCL over NVLink on GPUs; gloo in the CPU tests of this logic).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_batches(n_batches: int, world: int, rank: int):
    """Contiguous range [b0, b1) of whole batches for `rank`; the first n_batches % world ranks get one more."""
    base, extra = divmod(n_batches, world)
    b0 = rank * base + min(rank, extra)
    return b0, b0 + base + (1 if rank < extra else 0)


class ShardedPipeline:
    """Runs a MotifPipeline-like object on this rank's share of the query events and assembles global results.

    `pipeline` needs: .group, .W, .hist_null (int64[12] tensor), .run_host(src, dst, fake, ts, eidx, row_offset) ->
    ndarray [3, q, W] (or .run_device + staging for CUDA pipelines)."""

    def __init__(self, pipeline, group=None):
        self.pipe = pipeline
        self.group = int(group or pipeline.group)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0

    def run(self, src, dst, fake, ts, eidx, gather=True):
        Q, g = len(src), self.group
        if Q % g:
            raise ValueError(f"{Q} query events are not a whole number of reference batches of {g}")
        b0, b1 = shard_batches(Q // g, self.world, self.rank)
        sl = slice(b0 * g, b1 * g)
        W = self.pipe.W
        if b1 > b0:
            local = self.pipe.run_host(src[sl], dst[sl], fake[sl], ts[sl], eidx[sl], row_offset=3 * g * b0)
        else:
            local = np.zeros((3, 0, W), np.float32)
        hist = self.pipe.hist_null.clone()
        if self.world > 1:
            dist.all_reduce(hist)                                   # the null-model / marginal histogram exchange
        if not gather:
            return local, hist
        if self.world == 1:
            return local, hist
        # score gather: shards differ by at most one batch -> pad to the largest, all_gather, trim
        dev = hist.device
        qmax = (-(-(Q // g) // self.world)) * g
        buf = torch.zeros((3, qmax, W), dtype=torch.float32, device=dev)
        buf[:, :local.shape[1]] = torch.as_tensor(local).to(dev)
        out = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(out, buf)
        parts = []
        for r, o in enumerate(out):
            a, b = shard_batches(Q // g, self.world, r)
            parts.append(o[:, :(b - a) * g])
        return torch.cat(parts, dim=1).cpu().numpy(), hist

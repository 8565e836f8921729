"""Query sharding over the GPUs of one node (one process per GPU, torch.distributed).

The path shards by *whole reference batches* (the temporal attention normalises by a batch-global std,
explainer.py:826-828); graph, feature tables and weights are replicated.  Draws are keyed by the global root
row, so every split produces the same walks and scores as one GPU.  The only exchanges are the 12-bin class
histogram all-reduce and the score gather (NCCL over NVLink on GPUs; gloo in the CPU tests of this logic).

ScoreExchange fuses the score gather into the scorer kernel: the gathered buffer lives in symmetric memory (every rank maps every
peer's copy over NVLink), the kernel stores each score into its rank's segment of every copy, and the histogram all-reduce that
follows on the stream orders the peers' reads after those stores.
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


def segment_offsets(rows_per_rank: int, W: int, world: int, rank: int, base_ptrs):
    """Device addresses of `rank`'s [rows_per_rank, W] float32 segment inside every OTHER rank's gathered buffer [world, rows_per_rank, W]
    (base_ptrs[p] = address of rank p's buffer as mapped on this GPU)."""
    seg = rank * rows_per_rank * W * 4
    return [int(base_ptrs[p]) + seg for p in range(world) if p != rank]


class ScoreExchange:
    """Gathered scores [world, rows_per_rank, W] in symmetric memory; `local` is this rank's segment (pass it as `out` of
    MotifPipeline.run_device / TempME.score_device) and `peer_ptrs` the same segment on the peers (pass as `peer_ptrs`).
    After the step's histogram all-reduce has completed on the stream, `gathered` holds every rank's scores.
    Before a rank overwrites its segment again, all ranks must have consumed the previous contents (any collective between the
    consumer and the next step orders that; bench.py's per-step token all-reduce does).
    Raises RuntimeError when symmetric memory is unavailable (callers fall back to dist.all_gather_into_tensor)."""

    def __init__(self, rows_per_rank: int, W: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("ScoreExchange: at most 8 ranks (one NVSwitch domain)")
        grp = group if group is not None else dist.group.WORLD
        try:
            if hasattr(symm, "is_symm_mem_enabled_for_group") and not symm.is_symm_mem_enabled_for_group(grp.group_name):
                symm.enable_symm_mem_for_group(grp.group_name)
        except Exception:       # newer torch enables it lazily in rendezvous
            pass
        self.gathered = symm.empty((self.world, rows_per_rank, W), dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.gathered, grp)
        ptrs = list(self.handle.buffer_ptrs)
        if len(ptrs) != self.world or int(ptrs[self.rank]) != self.gathered.data_ptr():
            raise RuntimeError("ScoreExchange: unexpected symmetric-memory mapping")
        self.local = self.gathered[self.rank]
        self.peer_ptrs = segment_offsets(rows_per_rank, W, self.world, self.rank, ptrs)


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

"""CPU, world_size 2 over gloo: the N>1 host logic (batch sharding, global row offsets, histogram all-reduce,
score gather).  The CUDA pipeline is replaced by a stand-in whose output is a pure function of the GLOBAL root row,
so any mistake in offsets / ordering / trimming changes the assembled result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tempme_b200.dist import ShardedPipeline, shard_batches


class FakePipeline:
    def __init__(self, group, W):
        self.group, self.W = group, W
        self.hist_null = torch.zeros(12, dtype=torch.int64)

    def run_host(self, src, dst, fake, ts, eidx, row_offset=0):
        q, g = len(src), self.group
        nb = q // g
        # batch-major global rows: [batch][3][group] (MotifPipeline.stage_queries)
        rows = row_offset + np.arange(3 * q).reshape(nb, 3, g)
        vals = rows[..., None] * 1000 + np.arange(self.W)            # f(global row, walk)
        roots = np.stack([src.reshape(nb, g), dst.reshape(nb, g), fake.reshape(nb, g)], axis=1)
        vals = vals + roots[..., None] * 0.5                          # ... and of the right root id
        for k in range(3 * q):
            self.hist_null[(row_offset + k) % 12] += self.W
        return vals.transpose(1, 0, 2, 3).reshape(3, q, self.W).astype(np.float32)


def _worker(rank, world, port, Q, group, W, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    src, dst, fake = rng.integers(1, 50, Q), rng.integers(1, 50, Q), rng.integers(1, 50, Q)
    ts, eidx = np.sort(rng.random(Q)), np.arange(Q)
    sp = ShardedPipeline(FakePipeline(group, W))
    scores, hist = sp.run(src, dst, fake, ts, eidx)
    ref_pipe = FakePipeline(group, W)
    ref = ref_pipe.run_host(src, dst, fake, ts, eidx, 0)
    ok = np.array_equal(scores, ref) and torch.equal(hist, ref_pipe.hist_null)
    t = torch.tensor([int(ok)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_batches", [4, 5, 1])
def test_sharded_pipeline_world2(n_batches):
    group, W = 6, 4
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_batches * group, group, W, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


def test_shard_batches_partition():
    for n in range(0, 40):
        for w in (1, 2, 3, 4, 8):
            cover = []
            for r in range(w):
                a, b = shard_batches(n, w, r)
                assert 0 <= a <= b <= n and (b - a) in (n // w, n // w + 1)
                cover += list(range(a, b))
            assert cover == list(range(n))


def test_score_exchange_segment_offsets():
    """ScoreExchange's address arithmetic: rank r writes its [rows, W] float32 segment at the same offset of every OTHER rank's
    gathered buffer [world, rows, W]; the segments of all ranks tile a buffer without overlap."""
    from tempme_b200.dist import segment_offsets
    rows, W, world = 300, 30, 8
    bases = [0x7F00_0000_0000 + p * 0x4000_0000 for p in range(world)]
    seen = {}
    for r in range(world):
        ptrs = segment_offsets(rows, W, world, r, bases)
        assert len(ptrs) == world - 1
        peers = [p for p in range(world) if p != r]
        for p, a in zip(peers, ptrs):
            off = a - bases[p]
            assert off == r * rows * W * 4 and off % 4 == 0
            seen.setdefault(p, []).append((off, off + rows * W * 4))
    for p, segs in seen.items():
        segs.sort()
        assert all(a[1] <= b[0] for a, b in zip(segs, segs[1:]))     # no overlap between the writers of one buffer
        assert segs[-1][1] <= world * rows * W * 4

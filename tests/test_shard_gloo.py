"""world_size-2 gloo test of the batch-shard host logic (no GPU): sharded result == unsharded result,
bit for bit, for even and ragged batch sizes."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raft_b200.shard import shard_range, sharded_forward


def test_shard_range_partitions_the_batch():
    for B in (0, 1, 2, 5, 8, 32, 33):
        for W in (1, 2, 3, 8):
            spans = [shard_range(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_forward(l, r):  # deterministic per-sample function standing in for the GPU engine
    return torch.stack([(l[..., :2] * 3 - r[..., 1:]).cumsum(1)[i] for i in range(l.shape[0])]) if l.shape[0] else \
        l.new_zeros((0,) + tuple(l.shape[1:3]) + (2,))


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    l = torch.randn(B, 6, 5, 3, generator=g)
    r = torch.randn(B, 6, 5, 3, generator=g)
    out = sharded_forward(_fake_forward, l, r)
    ref = _fake_forward(l, r)
    q.put((rank, bool(torch.equal(out, ref)), tuple(out.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5, 1])
def test_sharded_equals_unsharded_gloo(B):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(60) for p in ps]
    assert all(ok for _, ok, _ in res), res
    assert all(shape[0] == B for _, _, shape in res)

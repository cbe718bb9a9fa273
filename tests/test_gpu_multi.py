"""Multi-GPU evidence for SURVEY 8(e) (run with `gpurun --gpus 2`; skipped on a 1-GPU box):

  * the batch shard as specified: one process per GPU, NCCL all_gather of the [B/G,H,W,2] flows over NVLink,
    gathered result == single-GPU result BIT FOR BIT (even, ragged and B < world batches, pinned-host inputs);
  * two engines on two devices inside ONE process (per-device function attributes, csrc/common.cuh PerDeviceOnce)."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_two():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")


def test_two_devices_in_one_process(cuda):
    _need_two()
    from raft_b200 import synth
    from networks.RAFT import RAFT
    for small in (False, True):
        p = synth.make_weights(small)
        l, r = synth.make_batch(1, 96, 160)
        outs = []
        for d in ("cuda:0", "cuda:1", "cuda:0"):
            m = RAFT((96, 160, 3), SimpleNamespace(small=small), iters=4, device=d).load(p)
            o = m.forward(l, r)
            assert o.device == torch.device(d)
            outs.append(o.cpu())
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def _worker(rank, world, port, small, H, W, iters, batches, q):
    import torch.distributed as dist
    from raft_b200 import synth
    from raft_b200.shard import sharded_forward
    from networks.RAFT import RAFT
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        p = synth.make_weights(small)
        m = RAFT((H, W, 3), SimpleNamespace(small=small), iters=iters, device=dev).load(p)
        for B in batches:
            l, r = synth.make_batch(B, H, W)
            lh, rh = torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()  # the documented e2e path
            out = sharded_forward(m.forward, lh, rh, gather=True, device=dev)
            ok = True
            if rank == 0:
                ref = m.forward(lh, rh)  # the whole batch on one GPU
                ok = bool(torch.equal(out, ref))
            q.put((rank, B, ok, tuple(out.shape), str(out.device)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("small,H,W,iters", [(False, 96, 160, 4), (True, 128, 192, 3)])
def test_sharded_forward_nccl_equals_single_gpu_bit_exact(cuda, small, H, W, iters):
    _need_two()
    import torch.multiprocessing as mp
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "raft-tf_b200"), os.environ.get("PYTHONPATH", "")])
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    batches = (4, 3, 1)  # even, ragged, and an EMPTY shard on rank 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, small, H, W, iters, batches, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=600) for _ in range(2 * len(batches))]
    [p.join(120) for p in ps]
    assert all(p.exitcode == 0 for p in ps), [p.exitcode for p in ps]
    assert all(ok for _, _, ok, _, _ in res), res
    for rank, B, _, shape, devname in res:
        assert shape == (B, H, W, 2) and devname == f"cuda:{rank}", (rank, B, shape, devname)

"""Batch sharding across the GPUs of one box.

Frame pairs are independent units: no stage of the path mixes samples (instance norm is per sample,
BN uses frozen statistics), so rank g of G simply runs samples [g*B/G, (g+1)*B/G) with its own copy
of the weights.  There is NO data-path collective; the only communication is the final gather of the
[B/G,H,W,2] flows (NCCL all_gather over NVLink on GPUs; gloo in the CPU tests), off the per-iteration
path.  One process per GPU (torchrun)."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split; the first (batch % world) ranks get one extra sample."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_forward(forward: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], left: torch.Tensor,
                    right: torch.Tensor, gather: bool = True, device=None) -> torch.Tensor:
    """Run `forward` on this rank's slice of the batch; optionally all_gather the flows so that every rank
    returns the full [B,H,W,2] result in the original order.  Works with any initialised process group.

    `device`: where this rank's flows live (the engine's CUDA device).  Only needed for a rank whose shard is EMPTY
    (B < world): its placeholder must sit on the same kind of device as the other ranks' results, not on the (possibly
    host / pinned) input's device, or the NCCL all_gather would mix CPU and CUDA tensors."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    B = left.shape[0]
    lo, hi = shard_range(B, rank, world)
    if hi > lo:
        local = forward(left[lo:hi], right[lo:hi])
    else:
        if device is None:
            device = (torch.device("cuda", torch.cuda.current_device())
                      if dist.is_initialized() and dist.get_backend() == "nccl" else left.device)
        local = torch.zeros((0,) + tuple(left.shape[1:3]) + (2,), dtype=torch.float32, device=device)
    if world == 1 or not gather:
        return local
    sizes = [shard_range(B, r, world) for r in range(world)]
    maxn = max(h - l for l, h in sizes)
    pad = local.new_zeros((maxn,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous())
    return torch.cat([parts[r][:h - l] for r, (l, h) in enumerate(sizes)], 0)

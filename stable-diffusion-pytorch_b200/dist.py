"""Data-parallel sharding of a generation batch across the GPUs of one node.

Every image (its latent, its CFG pair, its context rows) is independent through all sampling steps
(models/diffusion.py:223-236 has no cross-sample op), so the path shards with NO collective inside
a step: rank r takes images [r*B/R, (r+1)*B/R), weights are replicated, the CFG pair stays on one
rank.  The only collective is ONE all-gather of the final fp32 latents per generation (NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``total`` images for ``rank`` (first ``total % world`` ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_inputs(latent: torch.Tensor, context: torch.Tensor, rank: int, world: int, do_cfg: bool = True):
    """Slice full-batch tensors for one rank.  ``latent`` (B,4,h,w); ``context`` (2B,S,D) ordered
    [uncond ; cond] when do_cfg (models/diffusion.py:196) else (B,S,D).  The full batch is generated
    from ONE seeded generator and then sliced, so initial latents are bit-identical to a 1-GPU run."""
    B = latent.shape[0]
    lo, hi = shard_range(B, rank, world)
    lat = latent[lo:hi]
    if do_cfg:
        if context.shape[0] != 2 * B:
            raise ValueError(f"CFG context must have 2*B={2 * B} rows, got {context.shape[0]}")
        ctx = torch.cat([context[lo:hi], context[B + lo:B + hi]], dim=0)      # keep the CFG pair on this rank
    else:
        ctx = context[lo:hi] if context.shape[0] == B else context
    return lat.contiguous(), ctx.contiguous()


def gather_latents(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank final latents into the full (B,4,h,w) batch — the path's only collective."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < maxn:                                   # ragged last ranks: pad to a common size
        pad = torch.cat([local, local.new_zeros((maxn - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = local.new_empty((world * maxn,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * maxn: r * maxn + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, 0)

"""Batch sharding over the GPUs of one NVSwitch box.

Samples are independent through encode, VQ and decode, so the batch is split contiguously over
ranks with replicated weights; the only data-path collective is ONE all-gather of the code
indices after encode (the reference's inference path has none at all: every dist.* call in
modules/codebook.py:96-118 is training-only).  Decode runs on the local shard.
One process per GPU; torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests) is the plumbing.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of the batch dim; the first (batch % world) ranks take one extra sample."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    s, e = shard_bounds(x.shape[0], rank, world)
    return x[s:e]


def all_gather_codes(local_codes: torch.Tensor, batch_total: int, group=None) -> torch.Tensor:
    """Every rank ends up with the full (B, T', h, w) int64 index tensor.

    local_codes: this rank's (b_local, T', h, w) LongTensor (b_local may be 0 or differ by one between
    ranks).  Codes travel as int32 (n_codes <= 2^31) in a single all_gather_into_tensor; ragged shards
    are padded to the largest shard and trimmed after the gather."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_codes
    per = (batch_total + world - 1) // world
    tail = tuple(local_codes.shape[1:])
    send = torch.zeros((per,) + tail, dtype=torch.int32, device=local_codes.device)
    send[: local_codes.shape[0]] = local_codes.to(torch.int32)
    recv = torch.empty((world * per,) + tail, dtype=torch.int32, device=local_codes.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if batch_total % world == 0:
        return recv.to(torch.int64)
    parts = []
    for r in range(world):
        s, e = shard_bounds(batch_total, r, world)
        parts.append(recv[r * per: r * per + (e - s)])
    return torch.cat(parts, dim=0).to(torch.int64)


@torch.no_grad()
def encode_sharded(model, x_full: torch.Tensor, is_image: bool, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Encode this rank's shard of ``x_full`` and all-gather the indices.
    Returns (all codes (B,T',h,w), local codes)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    s, e = shard_bounds(x_full.shape[0], rank, world)
    local = model.encode(x_full[s:e].to(model.device), is_image)
    return all_gather_codes(local, x_full.shape[0], group), local

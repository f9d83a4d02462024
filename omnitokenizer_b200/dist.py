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


class GatheredCodes:
    """Handle of an in-flight all-gather of code indices: the collective runs on NCCL's stream while the
    caller keeps enqueueing work (decode does not depend on other ranks' codes); ``wait()`` joins it and
    returns the full (B, T', h, w) int64 tensor."""

    def __init__(self, work, recv, batch_total, world, per):
        self._work, self._recv, self._B, self._world, self._per = work, recv, batch_total, world, per
        self._out = None

    def wait(self) -> torch.Tensor:
        if self._out is None:
            if self._work is not None:
                self._work.wait()          # current stream waits for the collective (no host sync on CUDA)
            recv = self._recv
            if self._world > 1 and self._B % self._world != 0:
                parts = []
                for r in range(self._world):
                    s, e = shard_bounds(self._B, r, self._world)
                    parts.append(recv[r * self._per: r * self._per + (e - s)])
                recv = torch.cat(parts, dim=0)
            self._out = recv.to(torch.int64)
        return self._out


def all_gather_codes_async(local_codes: torch.Tensor, batch_total: int, group=None) -> GatheredCodes:
    """Start the single collective of the path.  Codes travel as int32 (n_codes <= 2^31) in ONE
    all_gather_into_tensor; ragged shards are padded to the largest shard and trimmed in wait()."""
    world = dist.get_world_size(group)
    if world == 1:
        return GatheredCodes(None, local_codes, batch_total, 1, batch_total)
    per = (batch_total + world - 1) // world
    tail = tuple(local_codes.shape[1:])
    if local_codes.shape[0] == per:
        send = local_codes.to(torch.int32)
    else:
        send = torch.zeros((per,) + tail, dtype=torch.int32, device=local_codes.device)
        send[: local_codes.shape[0]] = local_codes.to(torch.int32)
    recv = torch.empty((world * per,) + tail, dtype=torch.int32, device=local_codes.device)
    work = dist.all_gather_into_tensor(recv, send, group=group, async_op=True)
    return GatheredCodes(work, recv, batch_total, world, per)


def all_gather_codes(local_codes: torch.Tensor, batch_total: int, group=None) -> torch.Tensor:
    """Every rank ends up with the full (B, T', h, w) int64 index tensor (blocking form).
    local_codes: this rank's (b_local, T', h, w) LongTensor (b_local may be 0 or differ by one between ranks)."""
    return all_gather_codes_async(local_codes, batch_total, group).wait()


@torch.no_grad()
def encode_sharded(model, x_full: torch.Tensor, is_image: bool, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Encode this rank's shard of ``x_full`` and all-gather the indices.
    Returns (all codes (B,T',h,w), local codes)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    s, e = shard_bounds(x_full.shape[0], rank, world)
    local = model.encode(x_full[s:e].to(model.device), is_image)
    return all_gather_codes(local, x_full.shape[0], group), local

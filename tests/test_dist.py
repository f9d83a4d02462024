"""world_size-2 gloo test of the N>1 host logic (shard bounds + the single all-gather of code indices)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from omnitokenizer_b200 import dist as od


def test_shard_bounds_cover_batch():
    for B in (1, 3, 8, 13):
        for world in (1, 2, 4, 8):
            spans = [od.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = (torch.arange(B * 2 * 3 * 3).reshape(B, 2, 3, 3) * 7) % 8192
    s, e = od.shard_bounds(B, rank, world)
    got = od.all_gather_codes(full[s:e].clone(), B)
    q.put((rank, bool(torch.equal(got, full)), got.dtype == torch.int64))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5, 1])
def test_all_gather_codes_gloo(B):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok and dt for _, ok, dt in res), res

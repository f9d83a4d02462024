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


class _StubModel:
    """encode() = a per-sample function of the input (what batch sharding relies on); empty shards return the
    right-shaped empty tensor exactly like OmniTokenizer_VQGAN._empty_encode."""
    device = torch.device("cpu")

    def encode(self, x, is_image):
        B = x.shape[0]
        Tp, h, w = 1 + (x.shape[2] - 1) // 4, x.shape[-2] // 8, x.shape[-1] // 8
        base = (x.sum(dim=(1, 2, 3, 4)) * 1000).long().abs() % 8192
        return base.view(B, 1, 1, 1).expand(B, Tp, h, w).clone()


def _worker_sharded(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.rand((B, 3, 5, 16, 16), generator=torch.Generator().manual_seed(5))
    model = _StubModel()
    allc, local = od.encode_sharded(model, x, False)
    want = model.encode(x, False)
    s, e = od.shard_bounds(B, rank, world)
    q.put((rank, bool(torch.equal(allc, want)), bool(torch.equal(local, want[s:e])), tuple(local.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [1, 3])
def test_encode_sharded_world_larger_than_batch(B):
    """B < world ("replicas only beyond B"): the surplus rank contributes an empty, right-shaped shard and every rank still
    ends up with the full index tensor."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(a and b for _, a, b, _ in res), res
    if B == 1:
        assert sorted(r[3][0] for r in res) == [0, 1]


def test_empty_encode_shapes():
    """OmniTokenizer_VQGAN.encode on an empty batch returns right-shaped empty results without touching the engine."""
    import omnitokenizer_b200 as ob
    m = ob.OmniTokenizer_VQGAN(ob.canonical_args())
    m.engine = lambda: None
    idx = m.encode(torch.zeros(0, 3, 17, 256, 256), False)
    assert tuple(idx.shape) == (0, 5, 32, 32) and idx.dtype == torch.int64
    emb, idx = m.encode(torch.zeros(0, 3, 64, 64), True, include_embeddings=True)
    assert tuple(emb.shape) == (0, 8, 1, 8, 8) and tuple(idx.shape) == (0, 1, 8, 8)
    mv = ob.OmniTokenizer_VQGAN(ob.canonical_args(["--use_vae"]))
    mv.engine = lambda: None
    assert tuple(mv.encode(torch.zeros(0, 3, 17, 128, 128), False).shape) == (0, 8, 5, 16, 16)

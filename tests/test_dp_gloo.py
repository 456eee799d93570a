"""CPU, world_size 2, gloo: the host-side data-parallel logic (sharding, max-over-ranks timing rule,
bucketed gradient all-reduce == single-process gradient of the whole batch)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from handwriting_line_generation_b200 import dp
    try:
        # sharding covers every line exactly once
        total = 37
        b, e = dp.shard_range(total, rank, world)
        cnt = torch.tensor([e - b])
        dist.all_reduce(cnt)
        assert cnt.item() == total
        # timing rule
        assert dp.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        # gradient all-reduce of a sharded batch == full-batch gradient (mean loss over the global batch)
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
        x, y = torch.randn(12, 8), torch.randn(12, 4)
        full = [p.clone() for p in torch.autograd.grad(((model(x) - y) ** 2).sum() / 12, list(model.parameters()))]
        xs, ys = dp.shard_batch(x, 0, rank, world), dp.shard_batch(y, 0, rank, world)
        (((model(xs) - ys) ** 2).sum() / 12 * world).backward()   # local sum / global count, pre-scaled by world
        dp.GradBuckets(model.parameters(), bucket_bytes=256).reduce()
        for p, g in zip(model.parameters(), full):
            assert torch.allclose(p.grad, g, atol=1e-6), (p.grad - g).abs().max()
        # overlapped reducer: hooks + flat bucket views; two steps, grads zeroed in between, one parameter unused
        torch.manual_seed(0)
        model2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
        unused = torch.nn.Parameter(torch.ones(5))
        red = dp.GradReducer(list(model2.parameters()) + [unused], bucket_bytes=300)
        assert len(red.buckets) > 1
        for _ in range(2):
            red.zero_grad()
            (((model2(xs) - ys) ** 2).sum() / 12 * world).backward()
            red.finish()
            for p, g in zip(model2.parameters(), full):
                assert torch.allclose(p.grad, g, atol=1e-6), (p.grad - g).abs().max()
                assert any(p.grad.data_ptr() == v.data_ptr() for b in red.buckets for v in b["views"])
            assert unused.grad.abs().max() == 0
        red.remove()
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_range_balanced():
    from handwriting_line_generation_b200 import dp
    for total in (0, 1, 7, 128):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1

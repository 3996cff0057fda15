"""Host-side logic of the multi-GPU paths on CPU: two ranks over gloo (rendezvous on 127.0.0.1).
Covers the news-table shard / all-gather assembly, the Partition_Sampler-compatible impression split and
the metric all-reduce of evaluate.py, and the data-parallel gradient mean the trainer relies on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from news_recommendation_mind_b200 import evaluate as ev


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_rows, H, n_impr, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. sharded news table: rank r "encodes" rows [lo, hi) (here: a deterministic function of the row id)
        lo, hi, per = ev.shard_bounds(n_rows, world, rank)
        shard = torch.zeros(per, H)
        ids = torch.arange(lo, hi, dtype=torch.float32)
        shard[: hi - lo] = ids[:, None] * 10 + torch.arange(H, dtype=torch.float32)[None, :]
        table = ev.gather_news_shards(shard, n_rows)
        exp = torch.arange(n_rows, dtype=torch.float32)[:, None] * 10 + torch.arange(H, dtype=torch.float32)[None, :]
        assert table.shape == (n_rows, H) and torch.equal(table, exp)
        # 2. impression partition + metric reduction
        i0, i1 = ev.partition_bounds(n_impr, world, rank)
        g = torch.Generator().manual_seed(7)
        per_impr = torch.rand(n_impr, 4, generator=g, dtype=torch.float64)
        mean = ev.reduce_metric_sums(per_impr[i0:i1])
        assert torch.allclose(mean, per_impr.mean(0), rtol=0, atol=1e-12)
        # 3. data-parallel gradient mean (what DDP does for the dense gradients, twotower.py:49-50)
        grad = torch.full((5,), float(rank + 1))
        dist.all_reduce(grad)
        grad /= world
        assert torch.allclose(grad, torch.full((5,), (world + 1) / 2))
        # 4. trainer.GradSync's flat path (every gradient that the encoder-backward hook did not take): SUM over ranks,
        #    gradients re-pointed at 16-byte aligned views of one flat buffer, 1/world left to the optimiser's grad_scale
        import types
        from news_recommendation_mind_b200 import trainer
        torch.manual_seed(3)
        net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))          # 35 + 5 + 15 + 3 parameters: odd sizes
        opt = types.SimpleNamespace(grad_scale=1.0)
        sync = trainer.GradSync(net, opt)
        assert opt.grad_scale == 1.0 / world
        for p0 in net.parameters():                                                      # rank 0's weights everywhere
            ref = p0.detach().clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(ref, p0.detach())
        xs = torch.full((4, 7), float(rank + 1))
        net(xs).sum().backward()
        local = [p0.grad.detach().clone() for p0 in net.parameters()]
        sync.finish()
        for p0, g0 in zip(net.parameters(), local):
            tot = g0.clone()
            dist.all_reduce(tot)
            assert torch.allclose(p0.grad, tot, rtol=1e-6, atol=1e-6)
            assert p0.grad.data_ptr() % 16 == 0 and p0.grad.is_contiguous() and p0.grad.shape == p0.shape
        sync.close()
        assert opt.grad_scale == 1.0
        out.put((rank, i0, i1, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,n_impr", [(11, 7), (8, 8), (3, 1)])
def test_two_rank_eval_plumbing(n_rows, n_impr):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rows, 6, n_impr, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(world))
    # partitions tile [0, n_impr) contiguously, remainder on the last rank (utils.py:267-283)
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == n_impr
    assert got[0][2] - got[0][1] == n_impr // world
    # shards tile [0, n_rows)
    assert got[0][3] == 0 and got[0][4] == got[1][3] and got[1][4] == n_rows


def test_partition_matches_reference_sampler_rule():
    for n in range(0, 40):
        for ws in (1, 2, 3, 8):
            spans = [ev.partition_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            assert all(b - a == n // ws for a, b in spans[:-1])

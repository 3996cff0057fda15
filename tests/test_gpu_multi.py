"""2-GPU NCCL tests (run with `gpurun --gpus 2`; skipped when fewer than 2 devices are visible): data-parallel
training with DDP over the CUDA model must equal single-GPU training on the concatenated batch (gradient MEAN
semantics of twotower.py:49-50), and the sharded evaluation must equal the single-GPU evaluation."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda:%d" % rank))
    try:
        from helpers import build_model, manager_for, random_batch
        from news_recommendation_mind_b200 import data, evaluate as ev, trainer
        B, C, S, L, E, H, V = 8, 5, 10, 32, 300, 150, 2000
        torch.manual_seed(0)
        man = manager_for("cnn", "lstm", C, S, L, E, H, 10, precision="fp32", device="cuda:%d" % rank)
        model = build_model(man, V)
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank], output_device=rank)
        gen = torch.Generator().manual_seed(123)
        full = random_batch(gen, B * world, C, S, L, V)
        mine = {k: v[rank * B:(rank + 1) * B] for k, v in full.items()}
        logp = ddp(mine)[0]
        torch.nn.NLLLoss()(logp, mine["label"].cuda()).backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
        if rank == 0:
            ref = build_model(manager_for("cnn", "lstm", C, S, L, E, H, 10, precision="fp32", device="cuda:0"), V,
                              {k: v.detach().cpu() for k, v in model.state_dict().items()})
            lp = ref(full)[0]
            torch.nn.NLLLoss()(lp, full["label"].cuda()).backward()
            worst = 0.0
            for k, p in ref.named_parameters():
                a, b = grads[k].double(), p.grad.detach().cpu().double()
                worst = max(worst, float((a - b).norm() / (b.norm() + 1e-30)))
            out.put(("ddp_grad_rel_err", worst))
        # trainer.GradSync (in-place all-reduce, table gradient started from inside the encoder backward, 1/world folded
        # into Adam) == DDP on the bf16 path, gradients and one optimiser step
        import copy
        torch.manual_seed(1)
        mb = build_model(manager_for("cnn", "lstm", C, S, L, E, H, 10, precision="bf16", device="cuda:%d" % rank), V)
        for p_ in mb.parameters():
            dist.broadcast(p_.data, src=0)
        mb2 = copy.deepcopy(mb)
        dd = torch.nn.parallel.DistributedDataParallel(mb, device_ids=[rank], output_device=rank)
        o1 = trainer.FusedAdam(dd, lr=1e-3, bert_lr=1e-4)
        trainer.train_step(dd, mine, o1)
        g_ddp = {k: p.grad.detach().clone() for k, p in mb.named_parameters()}
        o2 = trainer.FusedAdam(mb2, lr=1e-3, bert_lr=1e-4)
        sync = trainer.GradSync(mb2, o2)
        trainer.train_step(mb2, mine, o2, sync)
        sync.close()
        torch.cuda.synchronize()
        worst_g, worst_p = 0.0, 0.0
        for (k, p1), (_, p2) in zip(mb.named_parameters(), mb2.named_parameters()):
            a, b = (p2.grad.detach().double() / world), g_ddp[k].double()
            e_ = float((a - b).norm() / (b.norm() + 1e-30))
            if rank == 0:
                print("gradsync vs ddp %-44s %.3e" % (k, e_), flush=True)
            worst_g = max(worst_g, e_)
            worst_p = max(worst_p, float((p1.detach().double() - p2.detach().double()).norm() / (p1.detach().double().norm() + 1e-30)))
        out.put(("gradsync_rank%d" % rank, (worst_g, worst_p)))
        # sharded evaluation == single-rank evaluation
        news_ids, news_mask = data.make_news_table(300, L, seed=5)
        impr = data.make_eval_impressions(news_ids, news_mask, 41, S, seed=9)
        model.eval()
        table = ev.encode_all_news(model, news_ids, news_mask)
        got = ev.evaluate(model, news_ids, news_mask, impr)
        dist.destroy_process_group()
        single_table = ev.encode_all_news(model, news_ids, news_mask)          # no process group -> world 1
        single = ev.evaluate(model, news_ids, news_mask, impr)
        out.put(("eval_rank%d" % rank, (got == single, bool(torch.equal(table, single_table)))))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_ddp_and_sharded_eval():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = dict(out.get(timeout=5) for _ in range(5))
    assert res["ddp_grad_rel_err"] < 2e-4, res
    assert res["eval_rank0"] == (True, True) and res["eval_rank1"] == (True, True), res
    for r in range(world):
        assert res["gradsync_rank%d" % r][0] < 1e-5 and res["gradsync_rank%d" % r][1] < 1e-6, res

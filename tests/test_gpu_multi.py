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
        sync = trainer.GradSync(mb2, o2, prewarm=0)
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
        # GradSync when the table receives TWO gradients in one backward (encode_news and encode_user called separately, as
        # the reference's forward does, TwoTowerBaseModel.py:65-75): each contribution gets its own all-reduce and both are
        # summed into .grad -- against plain autograd accumulation + an explicit all-reduce of every gradient
        from news_recommendation_mind_b200.twotower import TwoTowerBaseModel
        from news_recommendation_mind_b200 import ops as _ops
        mb3, mb4 = copy.deepcopy(mb2), copy.deepcopy(mb2)
        for m_ in (mb3, mb4):
            m_.zero_grad(set_to_none=True)
        lp = TwoTowerBaseModel.forward(mb3, mine)[0]                      # two NewsCNN calls -> two table gradients
        _ops.NLLMean.apply(lp, mine["label"].cuda()).backward()
        for p_ in mb3.parameters():
            if p_.grad is not None:
                dist.all_reduce(p_.grad)
        o4 = trainer.FusedAdam(mb4, lr=1e-3, bert_lr=1e-4)
        sync4 = trainer.GradSync(mb4, o4, prewarm=0)
        lp = TwoTowerBaseModel.forward(mb4, mine)[0]
        _ops.NLLMean.apply(lp, mine["label"].cuda()).backward()
        sync4.finish()
        sync4.close()
        torch.cuda.synchronize()
        worst2 = 0.0
        for (k, p3), (_, p4) in zip(mb3.named_parameters(), mb4.named_parameters()):
            if p3.grad is None:
                continue
            worst2 = max(worst2, float((p4.grad.double() - p3.grad.double()).norm() / (p3.grad.double().norm() + 1e-30)))
        out.put(("gradsync_two_calls_rank%d" % rank, worst2))
        # the whole data-parallel step as ONE CUDA graph (NCCL all-reduces captured) == eager GradSync steps
        mb5, mb6 = copy.deepcopy(mb2), copy.deepcopy(mb2)
        o5, o6 = trainer.FusedAdam(mb5, lr=1e-3, bert_lr=1e-4), trainer.FusedAdam(mb6, lr=1e-3, bert_lr=1e-4)
        sync5 = trainer.GradSync(mb5, o5, prewarm=0)
        mine_d = {k: v.cuda() for k, v in mine.items()}
        gs = trainer.GraphStep(mb5, o5, mine_d, sync5)
        l5 = [float(gs(mine_d).detach()) for _ in range(3)]
        gs.close()                                                        # before any communicator teardown (GraphStep.close)
        sync5.close()
        sync6 = trainer.GradSync(mb6, o6, prewarm=0)
        o6.enable_device_step_scalars("cuda:%d" % rank)
        l6 = []
        for _ in range(3):
            o6.begin_step()
            l6.append(float(trainer.train_step(mb6, mine_d, o6, sync6)))
        sync6.close()
        torch.cuda.synchronize()
        exact = l5 == l6 and all(torch.equal(a_, b_) for a_, b_ in zip(mb5.parameters(), mb6.parameters()))
        worst5 = max(float((a_.double() - b_.double()).norm() / (b_.double().norm() + 1e-30)) for a_, b_ in zip(mb5.parameters(), mb6.parameters()))
        dl = max(abs(a_ - b_) / max(1.0, abs(b_)) for a_, b_ in zip(l5, l6))
        out.put(("graph_sync_rank%d" % rank, (dl, worst5, exact)))
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


def run(world, timeout=420):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
        if p.is_alive():
            p.kill()
        assert p.exitcode == 0, p.exitcode
    return dict(out.get(timeout=5) for _ in range(1 + 4 * world))


def check(res, world):
    assert res["ddp_grad_rel_err"] < 2e-4, res
    for r in range(world):
        assert res["eval_rank%d" % r] == (True, True), res
        assert res["gradsync_rank%d" % r][0] < 1e-5 and res["gradsync_rank%d" % r][1] < 1e-6, res
        assert res["gradsync_two_calls_rank%d" % r] < 1e-6, res
        assert res["graph_sync_rank%d" % r][0] < 1e-6 and res["graph_sync_rank%d" % r][1] < 1e-6, res


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_ddp_gradsync_graph_and_sharded_eval(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    check(run(world), world)


if __name__ == "__main__":       # python tests/test_gpu_multi.py <world>: prints the measured errors (log kept under profiles/)
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    r = run(w)
    for k in sorted(r):
        print(k, r[k])
    check(r, w)
    print("PASS world=%d" % w)

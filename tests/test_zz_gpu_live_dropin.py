"""The drop-in claim end to end, on the GPU: the reference's OWN loops (utils/Manager.py, staged under oracle/_ref) drive THIS
package's TwoTower -- Manager.train (_get_loss, _get_optim with torch.optim.Adam + the linear warm-up schedule, the _train loop),
Manager.save / Manager.load, Manager.evaluate in fast mode (_eval_fast: encode_news over the MIND_news loader, the news.pt round
trip, init_embedding, predict_fast per dev row, _group_lists, cal_metric) -- over the reference's own MIND datasets and
DataLoaders, and land where the unmodified reference model lands on the host.

STATUS: written after round 2's GPU budget was spent, so it has not run on hardware yet.  Every call it makes is covered one by
one by tests that have (goldens, predict_fast, the news table, the trajectory test); it is marked as a NON-STRICT expected failure
so that an unrun test cannot mask the verified suite, and it sorts last.  An XPASS in the report is the result to read; remove the
marker once it has been seen green.  Skipped where the reference is not staged."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import ref_harness as RH            # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(RH.reference_root() is None, reason="reference not staged / mounted here")]


def _manager(MM, mind_attrs, **kw):
    man = object.__new__(MM.Manager)                                      # no argument parsing: the attributes the loops read
    base = dict(rank=-1, world_size=0, scale="demo", step=0, interval=10, save_epoch=False, epochs=2, smoothing=0.3, checkpoint=0,
                anomaly=False, lr=3e-3, bert_lr=1e-3, scheduler="linear", warmup=3, hold_step=10 ** 9, fast=True,
                metrics=["auc", "mean_mrr", "ndcg@5", "ndcg@10"])
    for k, v in dict(mind_attrs, **dict(base, **kw)).items():
        setattr(man, k, v)
    man._log = lambda res: None                                           # the result log file is not part of the path
    return man


@pytest.mark.timeout(600)
@pytest.mark.xfail(strict=False, reason="not yet run on hardware (written after the round's GPU budget was spent); XPASS = green")
@pytest.mark.parametrize("encu", ["lstm", "attn"])
def test_reference_manager_loops_drive_this_model_on_the_gpu(encu, tmp_path, monkeypatch):
    from torch.utils.data import DataLoader
    from helpers import build_model, manager_for
    from news_recommendation_mind_b200 import data
    from test_oracle_live_reference import _mind_manager, _stage_mind_files
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        import utils.Manager as MM
        from utils.MIND import MIND, MIND_news
    finally:
        sys.path.remove(root)
    monkeypatch.chdir(tmp_path)
    B, C, S, L, E, H, V, hn, n_news, n_users, impr_size, n_impr = 6, 5, 6, 12, 32, 16, 31000, 4, 60, 30, 7, 40
    ids, mask = data.make_news_table(n_news, L, seed=3)
    # ---- files of the train split (24 impressions, one positive each) and of the dev split (impressions cut at impr_size)
    x = data.make_train_batch(ids, mask, 4 * B, C, S, seed=9, n_users=n_users)
    lines = []
    for b in range(4 * B):
        his = [int(v) for v in x["his_id"][b] if int(v) != 0]
        impr = ["N%d-1" % int(x["cdd_id"][b, 0])] + ["N%d-0" % int(v) for v in x["cdd_id"][b, 1:]]
        lines.append("%d\tU%d\tt\t%s\t%s\n" % (b + 1, int(x["user_id"][b]), " ".join("N%d" % v for v in his), " ".join(impr)))
    train_dir = _stage_mind_files(tmp_path, "train", ids, mask, lines, n_users)
    ev = data.make_eval_impressions(ids, mask, n_impr, S, seed=4, n_users=n_users, impr_size=impr_size)
    off = ev["offsets"].tolist()
    n_rows = len(off) - 1
    lines = []
    for i in range(n_impr):
        rows = [r for r in range(n_rows) if int(ev["impr_index"][r]) == i]
        his = [int(v) for v in ev["his_id"][rows[0]] if int(v) != 0]
        cand = ["N%d-%d" % (int(ev["cdd_id"][j]), int(ev["label"][j])) for j in range(off[rows[0]], off[rows[-1] + 1])]
        lines.append("%d\tU%d\tt\t%s\t%s\n" % (i + 1, int(ev["user_id"][rows[0]]), " ".join("N%d" % v for v in his), " ".join(cand)))
    dev_dir = _stage_mind_files(tmp_path, "dev", ids, mask, lines, n_users)
    # ---- the unmodified reference model on the host, this package's model on the GPU, same initial weights
    ref = RH.build_model("cnn", encu, V=V, E=E, H=H, C=C, S=S, L=L, hn=hn, n_users=n_users, seed=21, dropout_p=0.0)
    gen = torch.Generator().manual_seed(21)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    ours = build_model(manager_for("cnn", encu, C, S, L, E, H, hn, precision="fp32", device="cuda:0", n_users=n_users), V,
                       state={k: v.detach().clone() for k, v in ref.state_dict().items()})
    assert ours.name == ref.name
    mind_train = vars(_mind_manager("train", C, S, L, impr_size))
    mind_dev = vars(_mind_manager("dev", C, S, L, impr_size))
    # the training batches are drawn ONCE (newsample shuffles the negatives with Python's `random`) and fed to both runs
    batches = list(DataLoader(MIND(_manager(MM, mind_train, name=ref.name), train_dir), batch_size=B, shuffle=False))
    assert len(batches) == 4 and set(batches[0]) >= {"cdd_encoded_index", "his_encoded_index", "his_mask", "label", "user_id"}
    losses = {}
    real_float = float

    class Spy(torch.nn.NLLLoss):                                          # the loop keeps only the epoch sum of the losses
        def forward(self, pred, label):
            out = super().forward(pred, label)
            losses.setdefault(pred.device.type, []).append(real_float(out.detach()))
            return out
    monkeypatch.setattr(MM.nn, "NLLLoss", Spy)
    ref.train()
    _manager(MM, mind_train, name=ref.name, scale="small").train(ref, [batches])
    ours.train()
    _manager(MM, mind_train, name=ours.name, scale="small").train(ours, [batches])
    torch.cuda.synchronize()
    assert len(losses["cpu"]) == len(losses["cuda"]) == 8
    assert all(abs(a - b) <= 1e-4 * max(1.0, abs(b)) for a, b in zip(losses["cuda"], losses["cpu"])), losses
    theirs = ref.state_dict()
    for k, v in ours.state_dict().items():
        # Adam divides by sqrt(v): where a gradient is ~0 its fp32 rounding noise becomes a fraction of one step (lr 3e-3)
        torch.testing.assert_close(v.detach().cpu(), theirs[k], rtol=1e-3, atol=3e-4, msg=lambda m: k + ": " + m)
    # ---- checkpoint written by the reference run, loaded into this model by Manager.evaluate(load=True), fast evaluation of both
    ref_man = _manager(MM, mind_dev, name=ref.name, checkpoint=8)
    ref_man.get_news_num = lambda: n_news
    ref_man.save(ref, 8, RH.make_optimizer(ref))
    loaders = [DataLoader(MIND(ref_man, dev_dir), batch_size=1), DataLoader(MIND_news(ref_man, dev_dir), batch_size=7)]
    want = ref_man.evaluate(ref, loaders, log=False)
    table_ref = torch.load("data/cache/tensors/%s/demo/dev/news.pt" % ref.name).clone()
    got = ref_man.evaluate(ours, loaders, load=True, log=False)
    table_ours = torch.load("data/cache/tensors/%s/demo/dev/news.pt" % ours.name, map_location="cpu")
    for k, v in ours.state_dict().items():
        assert torch.equal(v.detach().cpu(), theirs[k]), k                # Manager.load put the reference's weights in
    assert table_ours.shape == table_ref.shape == (n_news + 1, H)
    assert float((table_ours - table_ref).abs().max()) <= 1e-5 * max(1.0, float(table_ref.abs().max()))
    assert set(got) == set(want)
    assert all(abs(float(got[k]) - float(want[k])) <= 1e-4 + 1e-12 for k in want), (got, want)


@pytest.mark.timeout(300)
@pytest.mark.xfail(strict=False, reason="not yet run on hardware (written after the round's GPU budget was spent); XPASS = green")
def test_free_running_graph_steps_equal_eager_steps():
    """The race tests/test_host_logic.py::test_step_scalar_upload_survives_a_host_that_runs_ahead models on the CPU, on the
    device: 24 GraphStep replays queued back to back WITHOUT reading a loss in between (the host runs ~15 steps ahead of the
    device), under a LinearWarmupSchedule that moves the learning rates every step, must leave bit-identical parameters to 24
    eager steps.  With one pinned host block for the step scalars the queued uploads read later steps' values."""
    import copy
    from helpers import build_model, manager_for, random_batch
    from news_recommendation_mind_b200 import trainer
    B, C, S, L, E, H, V = 6, 5, 9, 32, 300, 150, 500
    gen = torch.Generator().manual_seed(4)
    batches = [{k: v.cuda() for k, v in random_batch(gen, B, C, S, L, V).items()} for _ in range(3)]
    torch.manual_seed(6)
    m1 = build_model(manager_for("cnn", "lstm", C, S, L, E, H, 10, precision="bf16"), V)
    m2 = copy.deepcopy(m1)
    o1, o2 = trainer.FusedAdam(m1, lr=1e-3, bert_lr=1e-4), trainer.FusedAdam(m2, lr=1e-3, bert_lr=1e-4)
    s1, s2 = trainer.LinearWarmupSchedule(o1, 5, 40), trainer.LinearWarmupSchedule(o2, 5, 40)
    gs = trainer.GraphStep(m1, o1, batches[0])
    torch.cuda.synchronize()
    for s in range(24):                                     # nothing here waits for the device
        gs(batches[s % 3])
        s1.step()
    o2.enable_device_step_scalars("cuda:0")                 # same arithmetic for the step scalars as the graph path
    for s in range(24):
        o2.begin_step()
        float(trainer.train_step(m2, batches[s % 3], o2))   # reads the loss: one step in flight at a time
        s2.step()
    torch.cuda.synchronize()
    assert o1.steps == o2.steps == 24
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), k
    gs.close()

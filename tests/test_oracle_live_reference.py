"""Live cross-check of the oracle against the reference ITSELF (oracle/_ref staged from /root/reference, or the mount), on seeds
that are NOT in the committed golden files: guards against goldens that only pin the cases they were generated for.
Skipped where the reference is not available (it never enters the repository)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import ref_harness as RH            # noqa: E402
from oracle import twotower_oracle as O         # noqa: E402
from helpers import random_batch                # noqa: E402

pytestmark = pytest.mark.skipif(RH.reference_root() is None, reason="reference not staged / mounted here")


@pytest.mark.parametrize("encn,encu,seed", [("cnn", "lstm", 101), ("cnn", "gru", 102), ("cnn", "attn", 103), ("cnn", "avg", 104),
                                            ("cnn", "mha", 105), ("mha", "lstm", 106)])
def test_oracle_matches_the_live_reference_on_fresh_seeds(encn, encu, seed):
    B, C, S, L, E, H, V, hn = 5, 4, 7, 12, 40, 20, 300, 5
    torch.set_num_threads(4)
    model = RH.build_model(encn, encu, V=V, E=E, H=H, C=C, S=S, L=L, hn=hn, n_users=40, seed=seed, dropout_p=0.0)
    model.train()
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():                                   # non-trivial values everywhere (biases and the padding row included)
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    x = random_batch(gen, B, C, S, L, V)
    logp = model(x)[0]
    loss = torch.nn.NLLLoss()(logp, x["label"])
    loss.backward()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ours = O.forward(params, x, True, encoder_n=encn, encoder_u=encu, head_num=hn)
    O.nll_loss(ours, x["label"]).backward()
    assert torch.allclose(ours, logp.detach(), rtol=1e-5, atol=1e-6), float((ours - logp).abs().max())
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        g = params[name].grad
        assert g is not None, name
        scale = float(p.grad.abs().max()) + 1e-12
        assert float((g - p.grad).abs().max()) <= 2e-5 * scale + 1e-7, (name, float((g - p.grad).abs().max()), scale)


def test_metrics_oracle_matches_the_live_cal_metric():
    """cal_metric (Manager.py:1276-1344: sklearn AUC, mrr / dcg through np.argsort(score)[::-1]) on fresh impressions against
    oracle/metrics_oracle.py: all four metrics on tie-free scores, AUC also on heavily tied ones.

    MRR / nDCG of TIED scores are not defined by the reference: np.argsort's default sort is unstable (numpy >= 2 sorts float64
    with a SIMD network even for 5 elements -- probed on this build --, numpy 1.x insertion-sorts up to 16 elements and
    introsorts beyond), so the order among equal scores depends on the numpy build and the CPU.  The oracle and mr_rank_metrics
    use the numpy-1.x small-array behaviour everywhere (stable sort, reversed: the later position first); AUC is order free."""
    import numpy as np
    from oracle import metrics_oracle as M
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        from utils.Manager import cal_metric
    finally:
        sys.path.remove(root)
    rng = np.random.default_rng(2024)

    def impressions(count, lo, hi, tied):
        labels, preds = [], []
        for _ in range(count):
            n = int(rng.integers(lo, hi))
            y = (rng.random(n) < 0.12).astype(np.float64)
            y[int(rng.integers(0, n))] = 1.0
            if y.sum() == n:
                y[0] = 0.0
            p = np.round(rng.random(n) * 5) / 5 if tied else rng.permutation(n) / n + rng.random() * 1e-3
            labels.append(y.tolist())
            preds.append((1.0 / (1.0 + np.exp(-p))).tolist())
        return labels, preds
    labels, preds = impressions(300, 2, 300, tied=False)
    ref = cal_metric(labels, preds, ["auc", "mean_mrr", "ndcg@5", "ndcg@10"])
    ours = M.ranking_metrics(labels, preds)
    assert ours == {k: float(ref[k]) for k in ours}, (ours, ref)
    labels, preds = impressions(200, 2, 200, tied=True)
    assert M.ranking_metrics(labels, preds)["auc"] == float(cal_metric(labels, preds, ["auc"])["auc"])


def test_grouping_and_partition_match_the_live_reference_utils():
    """utils.utils._group_lists (split impressions merged by impr_index, utils.py:60-80) and Partition_Sampler (utils.py:267-283)
    themselves, against oracle/metrics_oracle.py AND the product's host logic (evaluate.group_rows / partition_bounds)."""
    import numpy as np
    from oracle import metrics_oracle as M
    from news_recommendation_mind_b200 import evaluate as ev
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        from utils.utils import _group_lists, Partition_Sampler
    finally:
        sys.path.remove(root)
    rng = np.random.default_rng(5)
    for trial in range(20):
        n_rows = int(rng.integers(1, 60))
        # chunks of one impression are usually adjacent (MIND.py:225-226) but the gathered rows of several ranks are not sorted
        idx = rng.integers(0, max(2, n_rows // 2), size=n_rows)
        if trial % 2 == 0:
            idx = np.sort(idx)
        sizes = rng.integers(1, 6, size=n_rows)
        labels = [rng.integers(0, 2, size=s).tolist() for s in sizes]
        preds = [rng.random(s).tolist() for s in sizes]
        ref_l, ref_p = _group_lists(idx.tolist(), labels, preds)
        our_l, our_p = M.group_by_impression(idx.tolist(), labels, preds)
        assert our_l == ref_l and our_p == ref_p
        # product: row permutation + group offsets reproduce the same concatenation
        order, goff = ev.group_rows(torch.as_tensor(idx))
        rows = list(range(n_rows)) if order is None else order.tolist()
        got_p = [sum((preds[r] for r in rows[int(goff[g]):int(goff[g + 1])]), []) for g in range(goff.numel() - 1)]
        assert got_p == ref_p
    for n, ws in ((10, 3), (7, 7), (1000, 8), (5, 2), (376000, 8)):
        for r in range(ws):
            s = Partition_Sampler(range(n), ws, r)
            assert (s.start, s.end) == M.partition_bounds(n, ws, r) == tuple(ev.partition_bounds(n, ws, r))


@pytest.mark.parametrize("encu", ["lstm", "gru", "attn", "lstur"])
def test_eval_paths_match_the_live_reference(encu):
    """Manager._eval_fast's calls on the reference itself (Manager.py:498-517): encode_news in eval mode -> the [N+1, H] table,
    predict_fast over that table (news_reprs installed as init_embedding does, minus the torch.load from data/cache), and the slow
    eval forward (sigmoid) -- against oracle.encode_news / predict_fast / forward(training=False)."""
    B, C, S, L, E, H, V, hn, n_news = 6, 9, 8, 10, 32, 16, 200, 4, 25
    seed = 300 + len(encu)
    model = RH.build_model("cnn", encu, V=V, E=E, H=H, C=C, S=S, L=L, hn=hn, n_users=40, seed=seed, dropout_p=0.0)
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    model.eval()
    x = random_batch(gen, B, C, S, L, V)
    x["cdd_id"] = torch.randint(0, n_news + 1, (B, C), generator=gen)
    kw = {}
    if encu == "lstur":
        keep = torch.ones(B, dtype=torch.long)             # eval: the user embedding is always kept (no Bernoulli mask at test time)
        model.encoderU.keep_user = keep
        kw["keep_user"] = keep
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    # the news table, row 0 = the empty article [CLS][SEP] (MIND.py:125-127)
    tok = torch.randint(1, V, (n_news + 1, L), generator=gen)
    ln = torch.randint(2, L + 1, (n_news + 1,), generator=gen)
    msk = (torch.arange(L)[None, :] < ln[:, None]).long()
    tok = tok * msk
    with torch.no_grad():
        model.init_encoding()
        table = model.encode_news({"cdd_encoded_index": tok.unsqueeze(1), "cdd_attn_mask": msk.unsqueeze(1)}).squeeze(-2)
        model.destroy_encoding()
        model.news_reprs = torch.nn.Embedding.from_pretrained(table)
        ref_fast = model.predict_fast(x)
        ref_slow = model(x)[0]
    ours_table = O.encode_news(params, tok.unsqueeze(1), msk.unsqueeze(1), "cnn").squeeze(-2)
    assert torch.allclose(ours_table, table, rtol=1e-5, atol=1e-6)
    ours_fast = O.predict_fast(params, table, x, encoder_n="cnn", encoder_u=encu, head_num=hn, **kw)
    ours_slow = O.forward(params, x, False, encoder_n="cnn", encoder_u=encu, head_num=hn, **kw)
    assert ref_fast.shape == ours_fast.shape == (B, C)
    assert torch.allclose(ours_fast, ref_fast, rtol=1e-5, atol=1e-6), float((ours_fast - ref_fast).abs().max())
    assert torch.allclose(ours_slow, ref_slow, rtol=1e-5, atol=1e-6), float((ours_slow - ref_slow).abs().max())


@pytest.mark.parametrize("encu", ["lstm", "lstur"])
def test_checkpoints_interchange_through_the_reference_save_and_load(encu, tmp_path, monkeypatch):
    """Manager.save / Manager.load THEMSELVES (Manager.py:288-343, unbound, over a stand-in `self` carrying the three attributes
    they read) move a `.model` file between the reference's TwoTower + torch.optim.Adam and this package's TwoTower + FusedAdam:
    reference -> file -> ours, ours -> file -> reference, and a DDP-written file (`module.` prefixes) into a bare model."""
    import types
    from helpers import build_model, manager_for
    from news_recommendation_mind_b200 import trainer
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        from utils.Manager import Manager
    finally:
        sys.path.remove(root)
    B, C, S, L, E, H, V, hn = 4, 3, 5, 8, 24, 12, 120, 4
    ref = RH.build_model("cnn", encu, V=V, E=E, H=H, C=C, S=S, L=L, hn=hn, n_users=40, seed=77, dropout_p=0.0)
    ref_opt = RH.make_optimizer(ref)
    gen = torch.Generator().manual_seed(77)
    ref.train()
    for _ in range(2):                                      # moments and step counts worth saving
        x = random_batch(gen, B, C, S, L, V)
        if encu == "lstur":
            ref.encoderU.keep_user = torch.ones(B, dtype=torch.long)
        RH.train_step(ref, ref_opt, x)
    ours = build_model(manager_for("cnn", encu, C, S, L, E, H, hn, device="cpu"), V)
    ours_opt = trainer.FusedAdam(ours, lr=1.0, bert_lr=1.0)
    me = types.SimpleNamespace(name=ours.name, scale="demo", world_size=0)
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/model_params/%s" % me.name)

    def same(a, b):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa)

    def same_opt(fused, adam):
        assert [g["lr"] for g in fused.param_groups] == [g["lr"] for g in adam.param_groups]
        fp = [p for g in fused.param_groups for p in g["params"]]
        ap = [p for g in adam.param_groups for p in g["params"]]
        assert len(fp) == len(ap)
        for p, q in zip(fp, ap):
            if q in adam.state:
                assert torch.equal(fused.state[p][0], adam.state[q]["exp_avg"]) and torch.equal(fused.state[p][1], adam.state[q]["exp_avg_sq"])
                assert float(adam.state[q]["step"]) == fused.steps
    # reference -> ours
    Manager.save(me, ref, 2, ref_opt)
    assert os.path.exists("data/model_params/%s/demo_step2.model" % me.name)
    Manager.load(me, ours, 2, ours_opt)
    same(ours, ref)
    same_opt(ours_opt, ref_opt)
    # ours -> reference (a fresh reference model and optimiser)
    with torch.no_grad():
        for p in ours.parameters():
            p.mul_(1.5)
    Manager.save(me, ours, 3, ours_opt)
    ref2 = RH.build_model("cnn", encu, V=V, E=E, H=H, C=C, S=S, L=L, hn=hn, n_users=40, seed=5, dropout_p=0.0)
    ref2_opt = RH.make_optimizer(ref2, lr=1.0, bert_lr=1.0)
    Manager.load(me, ref2, 3, ref2_opt)
    same(ours, ref2)
    same_opt(ours_opt, ref2_opt)
    # a file written from a DDP-wrapped model (keys prefixed `module.`) into a bare model, world_size <= 1 (Manager.py:324-332)
    wrapped = torch.nn.Module()
    wrapped.module = ref
    Manager.save(me, wrapped, 4, ref_opt)
    ours2 = build_model(manager_for("cnn", encu, C, S, L, E, H, hn, device="cpu"), V)
    Manager.load(me, ours2, 4, trainer.FusedAdam(ours2))
    same(ours2, ref)


def test_optimizer_groups_and_schedule_match_the_reference_get_optim():
    """Manager._get_optim ITSELF (Manager.py:389-422) over this package's TwoTower: the parameters it puts in the `bert` group, the
    order inside both groups (the state-dict indices of a checkpoint), and the learning rates of the linear warm-up schedule over a
    whole run -- against trainer.FusedAdam + trainer.LinearWarmupSchedule."""
    import types
    from helpers import build_model, manager_for
    from news_recommendation_mind_b200 import trainer
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        from utils.Manager import Manager
    finally:
        sys.path.remove(root)
    for encn, encu in (("cnn", "lstm"), ("mha", "lstur"), ("cnn", "mha")):
        ours = build_model(manager_for(encn, encu, 3, 5, 8, 24, 12, 4, device="cpu"), 120)
        me = types.SimpleNamespace(world_size=0, lr=1e-4, bert_lr=6e-6, scheduler="linear", warmup=4, epochs=3)
        ref_opt, ref_sched = Manager._get_optim(me, ours, 7)                # 7 steps per epoch x 3 epochs
        opt = trainer.FusedAdam(ours, lr=me.lr, bert_lr=me.bert_lr)
        sched = trainer.LinearWarmupSchedule(opt, me.warmup, 7 * me.epochs)
        assert [[id(p) for p in g["params"]] for g in opt.param_groups] == [[id(p) for p in g["params"]] for g in ref_opt.param_groups]
        assert len(opt.param_groups[1]["params"]) == 1                      # the word-embedding table alone
        for _ in range(7 * me.epochs + 2):
            got, want = [g["lr"] for g in opt.param_groups], [g["lr"] for g in ref_opt.param_groups]
            assert all(abs(a - b) <= 1e-12 * max(abs(b), 1e-30) + 1e-30 for a, b in zip(got, want)), (got, want)
            for p in ours.parameters():
                p.grad = torch.zeros_like(p)
            ref_opt.step()
            ref_sched.step()
            sched.step()


def _stage_mind_files(tmp_path, mode, news_ids, news_mask, lines, n_users):
    """the files utils/MIND.py reads, relative to the working directory: behaviors.tsv, the two id dictionaries and the tokenised
    news cache (the tokeniser itself needs the network; its output format is MIND.py:139-150: two [N+1, 512] integer arrays)"""
    import json
    import pickle
    d = tmp_path / "data" / "MIND" / ("MINDdemo_%s" % mode)
    d.mkdir(parents=True)
    (d / "behaviors.tsv").write_text("".join(lines))
    dic = tmp_path / "data" / "dictionaries"
    dic.mkdir(parents=True, exist_ok=True)
    (dic / ("nid2idx_demo_%s.json" % mode)).write_text(json.dumps({"N%d" % i: i for i in range(1, news_ids.shape[0])}))
    (dic / "uid2idx_demo.json").write_text(json.dumps({"U%d" % i: i for i in range(1, n_users + 1)}))
    nc = tmp_path / "data" / "cache" / "MIND" / "news" / "bert" / ("MINDdemo_%s" % mode)
    nc.mkdir(parents=True)
    with open(nc / "news.pkl", "wb") as f:
        pickle.dump({"encoded_news": news_ids.numpy().copy(), "attn_mask": news_mask.numpy().copy()}, f)
    return "data/MIND/MINDdemo_%s/" % mode


def _mind_manager(mode, C, S, L, impr_size):
    import types
    m = types.SimpleNamespace(his_size=S, impr_size=impr_size, signal_length=L, npratio=C - 1, shuffle_pos=False, descend_history=False,
                              bert="bert", mode=mode, rank=-1, world_size=0)
    m.get_bert_for_cache = lambda: "bert"
    m.get_special_token_id = lambda tok: {"[PAD]": 0, "[SEP]": 102}[tok]
    return m


def test_batches_follow_the_live_dataset_class(tmp_path, monkeypatch):
    """utils/MIND.py ITSELF -- MINDBaseDataset.__init__ (init_behaviors from a behaviors.tsv, the impr_size chunking of
    MIND.py:225-226, the news cache), MIND.__getitem__ (MIND.py:296-405) and the DataLoader's default collate -- over the files of
    the very samples data.make_train_batch / data.make_eval_impressions draw: same keys, dtypes, shapes and values (the negatives
    of a training sample up to the order random.sample leaves them in)."""
    import numpy as np
    from torch.utils.data import default_collate
    from news_recommendation_mind_b200 import data
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        from utils.MIND import MIND
    finally:
        sys.path.remove(root)
    monkeypatch.chdir(tmp_path)
    B, C, S, L, n_news, n_users = 24, 5, 6, 12, 60, 30
    ids, mask = data.make_news_table(n_news, L, seed=3)
    # ---------------------------------------------------------------- train
    x = data.make_train_batch(ids, mask, B, C, S, seed=9, n_users=n_users)
    # sample 0 becomes a user without history, as data.py lays one out (his_len 0: ids all 0 = the empty article, his_mask[0] = 1,
    # MIND.py:332-337); the log-normal history lengths of the synthetic data almost never draw one
    x["his_id"][0] = 0
    x["his_encoded_index"][0], x["his_attn_mask"][0] = ids[0], mask[0]
    x["his_mask"][0] = 0
    x["his_mask"][0, 0] = 1
    lens = x["his_mask"].squeeze(-1).sum(-1).long()
    lines = []
    for b in range(B):
        his = [int(v) for v in x["his_id"][b] if int(v) != 0]
        assert len(his) == (int(lens[b]) if int(x["his_id"][b, 0]) != 0 else 0)
        impr = ["N%d-1" % int(x["cdd_id"][b, 0])] + ["N%d-0" % int(v) for v in x["cdd_id"][b, 1:]]
        lines.append("%d\tU%d\tt\t%s\t%s\n" % (b + 1, int(x["user_id"][b]), " ".join("N%d" % v for v in his), " ".join(impr)))
    ds = MIND(_mind_manager("train", C, S, L, 0), _stage_mind_files(tmp_path, "train", ids, mask, lines, n_users))
    assert len(ds) == B
    ref = default_collate([ds[i] for i in range(B)])
    assert set(ref) == set(x)
    for k in x:
        assert ref[k].dtype == x[k].dtype and ref[k].shape == x[k].shape, (k, ref[k].dtype, x[k].dtype, ref[k].shape, x[k].shape)
    for k in ("user_id", "his_id", "his_encoded_index", "his_attn_mask", "his_mask", "cdd_mask", "label"):
        assert torch.equal(ref[k], x[k]), k
    assert torch.equal(ref["cdd_id"][:, 0], x["cdd_id"][:, 0])
    assert torch.equal(ref["cdd_id"][:, 1:].sort(dim=1).values, x["cdd_id"][:, 1:].sort(dim=1).values)
    assert torch.equal(ref["cdd_encoded_index"], ids[ref["cdd_id"]]) and torch.equal(ref["cdd_attn_mask"], mask[ref["cdd_id"]])
    # ---------------------------------------------------------------- dev, impressions longer than impr_size cut into chunks
    impr_size, n_impr = 7, 40
    ev = data.make_eval_impressions(ids, mask, n_impr, S, seed=4, n_users=n_users, impr_size=impr_size)
    off = ev["offsets"].tolist()
    n_rows = len(off) - 1
    assert n_rows > n_impr                                      # some impressions were cut
    lines = []
    for i in range(n_impr):
        rows = [r for r in range(n_rows) if int(ev["impr_index"][r]) == i]
        assert rows == list(range(rows[0], rows[-1] + 1))      # chunks of one impression are adjacent
        his = [int(v) for v in ev["his_id"][rows[0]] if int(v) != 0]
        cand = ["N%d-%d" % (int(ev["cdd_id"][j]), int(ev["label"][j])) for j in range(off[rows[0]], off[rows[-1] + 1])]
        lines.append("%d\tU%d\tt\t%s\t%s\n" % (i + 1, int(ev["user_id"][rows[0]]), " ".join("N%d" % v for v in his), " ".join(cand)))
    ds = MIND(_mind_manager("dev", C, S, L, impr_size), _stage_mind_files(tmp_path, "dev", ids, mask, lines, n_users))
    assert len(ds) == n_rows
    for r in range(n_rows):
        got = default_collate([ds[r]])                          # the dev loader's batch size is 1 (Manager.py:228-238)
        assert int(got["impr_index"]) == int(ev["impr_index"][r]) + 1          # 1-based in the reference; only equality is used
        assert int(got["user_id"]) == int(ev["user_id"][r])
        assert got["cdd_id"][0].tolist() == ev["cdd_id"][off[r]:off[r + 1]].tolist()
        assert got["label"][0].tolist() == [int(v) for v in ev["label"][off[r]:off[r + 1]]]
        for k in ("his_id", "his_encoded_index", "his_attn_mask", "his_mask"):
            assert got[k].dtype == ev[k].dtype and torch.equal(got[k][0], ev[k][r]), (k, r)
        assert torch.equal(got["cdd_encoded_index"][0], ids[got["cdd_id"][0]])


@pytest.mark.parametrize("encu", ["lstm", "attn"])
def test_training_trajectory_matches_the_live_training_loop(encu, tmp_path, monkeypatch):
    """Manager.train ITSELF (Manager.py:700-722: _get_loss, _get_optim with the linear warm-up schedule, then the _train loop of
    Manager.py:586-688) drives the reference model over two epochs of four batches; the oracle's train_step (forward, NLLLoss,
    backward, two-group Adam) with the schedule's learning rates must land on the same losses and the same parameters."""
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        import utils.Manager as MM
    finally:
        sys.path.remove(root)
    monkeypatch.chdir(tmp_path)
    B, C, S, L, E, H, V, hn = 6, 4, 5, 10, 32, 16, 150, 4
    model = RH.build_model("cnn", encu, V=V, E=E, H=H, C=C, S=S, L=L, hn=hn, n_users=40, seed=55, dropout_p=0.0)
    gen = torch.Generator().manual_seed(55)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    batches = [random_batch(gen, B, C, S, L, V) for _ in range(4)]
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    man = object.__new__(MM.Manager)                                      # no argument parsing: the attributes the loop reads
    for k, v in dict(rank=-1, world_size=0, name="live", scale="small", step=0, interval=10, save_epoch=False, epochs=2, smoothing=0.3,
                     checkpoint=0, anomaly=False, lr=3e-3, bert_lr=1e-3, scheduler="linear", warmup=3, hold_step=10 ** 9).items():
        setattr(man, k, v)
    man._log = lambda res: None                                           # the result log file is not part of the path
    losses = []
    real_float = float

    class Spy(torch.nn.NLLLoss):                                          # the loop keeps only the epoch sum of the losses
        def forward(self, pred, label):
            out = super().forward(pred, label)
            losses.append(real_float(out.detach()))
            return out
    monkeypatch.setattr(MM.nn, "NLLLoss", Spy)
    model.train()
    man.train(model, [batches])
    assert len(losses) == 8
    total, warm = 8, 3
    state, ours = {}, []
    for s in range(total):
        f = s / warm if s < warm else max(0.0, (total - s) / (total - warm))          # get_linear_schedule_with_warmup
        loss, _ = O.train_step(params, state, batches[s % 4], s + 1, lr=3e-3 * f, bert_lr=1e-3 * f, encoder_n="cnn", encoder_u=encu,
                               head_num=hn)
        ours.append(loss)
    assert all(abs(a - b) <= 1e-5 * max(1.0, abs(b)) for a, b in zip(ours, losses)), (ours, losses)
    for k, ref in model.state_dict().items():
        torch.testing.assert_close(params[k], ref, rtol=1e-4, atol=1e-4, msg=lambda m: k + ": " + m)


@pytest.mark.parametrize("encu", ["lstm", "attn"])
def test_fast_evaluation_matches_the_live_manager_evaluate(encu, tmp_path, monkeypatch):
    """Manager.evaluate ITSELF in fast mode (Manager.py:545-584 -> _eval_fast :473-541 -> cal_metric :1276-1344) over the reference's
    own datasets and DataLoaders (MIND dev with impressions cut at impr_size, MIND_news), on the reference model: the metrics
    dictionary against the oracle pipeline the GPU tests are held to (encode_news table incl. row 0 = the encoded empty article,
    predict_fast per chunk, chunks of one impression merged, AUC / MRR / nDCG@5/10 rounded to 4 decimals)."""
    from torch.utils.data import DataLoader
    from news_recommendation_mind_b200 import data
    from oracle import metrics_oracle as M
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        import utils.Manager as MM
        from utils.MIND import MIND, MIND_news
    finally:
        sys.path.remove(root)
    monkeypatch.chdir(tmp_path)
    C, S, L, E, H, V, hn, n_news, n_users, impr_size, n_impr = 5, 6, 12, 32, 16, 31000, 4, 60, 30, 7, 40
    ids, mask = data.make_news_table(n_news, L, seed=3)
    ev = data.make_eval_impressions(ids, mask, n_impr, S, seed=4, n_users=n_users, impr_size=impr_size)
    off = ev["offsets"].tolist()
    n_rows = len(off) - 1
    lines = []
    for i in range(n_impr):
        rows = [r for r in range(n_rows) if int(ev["impr_index"][r]) == i]
        his = [int(v) for v in ev["his_id"][rows[0]] if int(v) != 0]
        cand = ["N%d-%d" % (int(ev["cdd_id"][j]), int(ev["label"][j])) for j in range(off[rows[0]], off[rows[-1] + 1])]
        lines.append("%d\tU%d\tt\t%s\t%s\n" % (i + 1, int(ev["user_id"][rows[0]]), " ".join("N%d" % v for v in his), " ".join(cand)))
    directory = _stage_mind_files(tmp_path, "dev", ids, mask, lines, n_users)
    model = RH.build_model("cnn", encu, V=V, E=E, H=H, C=C, S=S, L=L, hn=hn, n_users=n_users, seed=21, dropout_p=0.0)
    gen = torch.Generator().manual_seed(21)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    man = object.__new__(MM.Manager)
    dm = _mind_manager("dev", C, S, L, impr_size)
    for k, v in dict(vars(dm), name=model.name, scale="demo", fast=True, smoothing=0.3, checkpoint=0,
                     metrics=["auc", "mean_mrr", "ndcg@5", "ndcg@10"]).items():
        setattr(man, k, v)
    man.get_news_num = lambda: n_news
    man._log = lambda res: None
    loaders = [DataLoader(MIND(man, directory), batch_size=1), DataLoader(MIND_news(man, directory), batch_size=7)]
    ref = man.evaluate(model, loaders, log=False)
    # ---- the oracle pipeline
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        table = O.encode_news(params, ids.unsqueeze(1), mask.unsqueeze(1), "cnn").squeeze(-2)
        assert float(table[0].abs().max()) > 0                              # MIND_news starts at index 0: row 0 IS encoded
        saved = torch.load("data/cache/tensors/%s/demo/dev/news.pt" % model.name)
        assert torch.allclose(table, saved, rtol=1e-5, atol=1e-6)
        idx, labels, preds = [], [], []
        for r in range(n_rows):
            x = {"cdd_id": ev["cdd_id"][off[r]:off[r + 1]].unsqueeze(0), "his_encoded_index": ev["his_encoded_index"][r:r + 1],
                 "his_attn_mask": ev["his_attn_mask"][r:r + 1], "his_mask": ev["his_mask"][r:r + 1], "user_id": ev["user_id"][r:r + 1]}
            p = O.predict_fast(params, table, x, encoder_n="cnn", encoder_u=encu, head_num=hn)
            idx.append(int(ev["impr_index"][r]))
            preds.append(p[0].tolist())
            labels.append(ev["label"][off[r]:off[r + 1]].tolist())
    gl, gp = M.group_by_impression(idx, labels, preds)
    assert len(gl) == n_impr < n_rows
    ours = M.ranking_metrics(gl, gp)
    assert ours == {k: float(ref[k]) for k in ours}, (ours, ref)


def test_prediction_file_matches_the_live_manager_test(tmp_path, monkeypatch):
    """Manager.test ITSELF (Manager.py:815-850: load the checkpoint, _test_fast over the MIND test split -- whose history arrives
    REVERSED under the default descend_history=False, MIND.py:423-426 --, merge the chunks, write prediction.txt with scipy's ordinal
    ranks) against the oracle's predictions ranked by metrics_oracle.ordinal_rank (the rule mr_rank_metrics implements bit for bit)
    and written by evaluate.write_predictions.  The only textual difference allowed is the `.0` scipy >= 1.10 appends to the ranks
    (the reference's pinned scipy prints integers, and so does the MIND submission format)."""
    import numpy as np
    from torch.utils.data import DataLoader
    from news_recommendation_mind_b200 import data, evaluate as evl
    from oracle import metrics_oracle as M
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        import utils.Manager as MM
        from utils.MIND import MIND, MIND_news
    finally:
        sys.path.remove(root)
    monkeypatch.chdir(tmp_path)
    C, S, L, E, H, V, hn, n_news, n_users, impr_size, n_impr = 5, 6, 12, 32, 16, 31000, 4, 60, 30, 7, 25
    ids, mask = data.make_news_table(n_news, L, seed=13)
    ev = data.make_eval_impressions(ids, mask, n_impr, S, seed=14, n_users=n_users, impr_size=impr_size)
    off = ev["offsets"].tolist()
    n_rows = len(off) - 1
    lines = []
    for i in range(n_impr):
        rows = [r for r in range(n_rows) if int(ev["impr_index"][r]) == i]
        his = [int(v) for v in ev["his_id"][rows[0]] if int(v) != 0]
        cand = ["N%d" % int(ev["cdd_id"][j]) for j in range(off[rows[0]], off[rows[-1] + 1])]          # the test split has no labels
        lines.append("%d\tU%d\tt\t%s\t%s\n" % (i + 1, int(ev["user_id"][rows[0]]), " ".join("N%d" % v for v in his), " ".join(cand)))
    directory = _stage_mind_files(tmp_path, "test", ids, mask, lines, n_users)
    model = RH.build_model("cnn", "lstm", V=V, E=E, H=H, C=C, S=S, L=L, hn=hn, n_users=n_users, seed=31, dropout_p=0.0)
    model.mode = "test"                                         # TwoTowerBaseModel.py:12 under `-m test`: where init_embedding looks for news.pt
    gen = torch.Generator().manual_seed(31)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    man = object.__new__(MM.Manager)
    for k, v in dict(vars(_mind_manager("test", C, S, L, impr_size)), name=model.name, scale="demo", fast=True, smoothing=0.3,
                     checkpoint=5).items():
        setattr(man, k, v)
    man.get_news_num = lambda: n_news
    os.makedirs("data/model_params/%s" % model.name)
    man.save(model, 5, RH.make_optimizer(model))
    loaders = [DataLoader(MIND(man, directory), batch_size=1), DataLoader(MIND_news(man, directory), batch_size=7)]
    man.test(model, loaders)
    ref_text = open("data/results/%s/demo_step5/prediction.txt" % model.name).read()
    # ---- ours
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        table = O.encode_news(params, ids.unsqueeze(1), mask.unsqueeze(1), "cnn").squeeze(-2)
        idx, preds = [], []
        for r in range(n_rows):
            n_his = int((ev["his_id"][r] != 0).sum())
            order = list(range(n_his))[::-1] + list(range(n_his, S))                                     # MIND.py:423-426
            x = {"cdd_id": ev["cdd_id"][off[r]:off[r + 1]].unsqueeze(0), "his_encoded_index": ev["his_encoded_index"][r:r + 1][:, order],
                 "his_attn_mask": ev["his_attn_mask"][r:r + 1][:, order], "his_mask": ev["his_mask"][r:r + 1], "user_id": ev["user_id"][r:r + 1]}
            idx.append(int(ev["impr_index"][r]))
            preds.append(O.predict_fast(params, table, x, encoder_n="cnn", encoder_u="lstm", head_num=hn)[0].tolist())
    merged = M.group_by_impression(idx, preds)[0]
    ranks = np.concatenate([M.ordinal_rank(np.asarray(p)) for p in merged])
    offsets = np.concatenate([[0], np.cumsum([len(p) for p in merged])])
    assert evl.write_predictions("ours.txt", torch.from_numpy(ranks.astype(np.int32)), torch.from_numpy(offsets)) == n_impr
    assert open("ours.txt").read() == ref_text.replace(".0", "")
    assert len(ref_text.splitlines()) == n_impr < n_rows

"""CPU tests of the host-side pieces that need no GPU: the synthetic MIND-shaped generator (SURVEY.md 8d input
contract), the optimiser's learning-rate groups (Manager._get_optim, utils/Manager.py:389-413), padding helpers,
and the staging-buffer compatibility check of the training loop."""
import re
import types

import numpy as np
import pytest
import torch

from news_recommendation_mind_b200 import data, ops, trainer
from oracle import metrics_oracle as MO


def test_news_table_follows_the_reference_layout():
    """MIND.py:103-127: row 0 is the empty article [CLS] [SEP] PAD..., every title starts with [CLS]=101, ends with
    [SEP]=102 at its last valid position, PAD=0 afterwards; the attention mask marks exactly the valid prefix."""
    L = 32
    ids, mask = data.make_news_table(300, L, seed=3)
    assert ids.dtype == torch.int64 and mask.dtype == torch.int64 and ids.shape == (301, L) and mask.shape == (301, L)
    assert ids[0, :2].tolist() == [101, 102] and int(ids[0, 2:].abs().sum()) == 0 and mask[0].tolist() == [1, 1] + [0] * (L - 2)
    ln = mask.sum(1)
    assert int(ln.min()) >= 2 and int(ln.max()) <= L
    pos = torch.arange(L)[None, :]
    assert torch.equal(mask, (pos < ln[:, None]).long())                      # valid prefix
    assert bool((ids[:, 0] == 101).all())
    assert bool((ids.gather(1, (ln - 1)[:, None]).squeeze(1) == 102).all())   # [SEP] closes every title
    assert int((ids * (1 - mask)).abs().sum()) == 0                           # PAD after the title
    assert int(ids.max()) < 30522 and int(ids.min()) >= 0


def test_train_batch_schema_matches_the_reference_collate():
    """MIND.__getitem__ + default collate (MIND.py:352-363): field names, dtypes and shapes; history right padded
    with news 0 and his_mask = 1 on the first max(len, 1) slots; label = index of the positive (always 0)."""
    B, C, S, L = 16, 5, 50, 32
    ids, mask = data.make_news_table(500, L, seed=1)
    x = data.make_train_batch(ids, mask, B, C, S, seed=7)
    assert x["cdd_encoded_index"].shape == (B, C, L) and x["his_encoded_index"].shape == (B, S, L)
    assert x["cdd_attn_mask"].shape == (B, C, L) and x["his_attn_mask"].shape == (B, S, L)
    for k in ("cdd_encoded_index", "his_encoded_index", "cdd_attn_mask", "his_attn_mask", "user_id", "cdd_id", "his_id", "label"):
        assert x[k].dtype == torch.int64, k
    assert x["his_mask"].dtype == torch.float64 and x["his_mask"].shape == (B, S, 1)
    assert x["label"].shape == (B,) and int(x["label"].abs().sum()) == 0
    hm = x["his_mask"].squeeze(-1)
    ln = hm.sum(1).long()
    assert int(ln.min()) >= 1
    assert torch.equal(hm, (torch.arange(S)[None, :] < ln[:, None]).double())  # a prefix of ones
    his_id = x["his_id"]
    assert int((his_id * (1 - hm).long()).abs().sum()) == 0                    # padded slots hold news 0
    assert torch.equal(x["his_encoded_index"], ids[his_id]) and torch.equal(x["cdd_encoded_index"], ids[x["cdd_id"]])
    assert int(x["cdd_id"].min()) >= 1                                         # candidates are real news
    y = data.make_train_batch(ids, mask, B, C, S, seed=7)
    assert all(torch.equal(x[k], y[k]) for k in x)                             # seeded: reproducible


def test_eval_impressions_are_well_formed():
    L, S = 32, 50
    ids, mask = data.make_news_table(400, L, seed=2)
    imp = data.make_eval_impressions(ids, mask, 25, S, seed=5)
    off = imp["offsets"]
    assert off[0] == 0 and off.numel() == 26 and off[-1] == imp["cdd_id"].numel() == imp["label"].numel()
    n = np.diff(np.asarray(off))
    assert n.min() >= 2 and n.max() <= 300
    lab = np.asarray(imp["label"])
    for i in range(25):
        seg = lab[off[i]:off[i + 1]]
        assert seg.max() == 1 and seg.min() == 0                               # AUC defined: one positive, one negative at least
        c = np.asarray(imp["cdd_id"][off[i]:off[i + 1]])
        assert len(set(c.tolist())) == len(c) and c.min() >= 1                 # distinct candidates


def test_fused_adam_groups_follow_the_bert_regex():
    """Manager._get_optim: parameters whose NAME matches the regex `bert` get bert_lr, the others lr."""
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embedding = torch.nn.Module()
            self.embedding.bert_word_embedding = torch.nn.Embedding(7, 4)
            self.encoderN = torch.nn.Linear(4, 3)
            self.encoderU = torch.nn.Linear(3, 3)
    m = M()
    opt = trainer.FusedAdam(m, lr=1e-4, bert_lr=6e-6)
    base, bert = opt.param_groups
    names = dict(m.named_parameters())
    assert base["lr"] == 1e-4 and bert["lr"] == 6e-6
    assert [id(p) for p in bert["params"]] == [id(p) for n, p in names.items() if re.search("bert", n)]
    assert len(bert["params"]) == 1 and len(base["params"]) == 4
    for g in opt.param_groups:
        for p in g["params"]:
            p.grad = torch.ones_like(p)
    opt.zero_grad(set_to_none=False)
    assert all(float(p.grad.abs().sum()) == 0 for p in m.parameters())
    opt.zero_grad()
    assert all(p.grad is None for p in m.parameters())


def test_padding_helpers_and_flags():
    assert ops.pad_to(150, 16) == 160 and ops.pad_to(160, 16) == 160 and ops.pad_to(300, 64) == 320 and ops.pad_to(30522, 128) == 30592
    assert ops.PRECISIONS["bf16"] != ops.PRECISIONS["fp32"] and ops.PRECISIONS["f32"] == ops.PRECISIONS["fp32"]
    assert isinstance(ops.GROUPED_TABLE_GRAD, bool)


def test_prefetcher_buffer_compatibility_check():
    pf = types.SimpleNamespace()
    fits = trainer.BatchPrefetcher._fits
    a = {"ids": torch.zeros(4, 3, dtype=torch.int64), "mask": torch.zeros(4, 1, dtype=torch.float64), "tag": "x"}
    buf = {"ids": torch.empty(4, 3, dtype=torch.int64), "mask": torch.empty(4, 1, dtype=torch.float64)}
    assert fits(pf, buf, a)
    assert not fits(pf, None, a)
    assert not fits(pf, buf, dict(a, ids=torch.zeros(5, 3, dtype=torch.int64)))          # another batch size
    assert not fits(pf, buf, dict(a, mask=torch.zeros(4, 1, dtype=torch.float32)))        # another dtype
    assert not fits(pf, buf, dict(a, extra=torch.zeros(1)))                               # a tensor the buffers do not hold


def test_adam_device_block_follows_the_schedule():
    """FusedAdam's CUDA-graph mode keeps every step-dependent scalar in one device block per launch
    {1/bc1, 1/sqrt(bc2), grad_scale, -, lr[...]}: the host mirror must follow a LinearWarmupSchedule and grad_scale
    (the captured launch froze lr at capture time)."""
    m = torch.nn.Sequential(torch.nn.Linear(3, 2), torch.nn.Linear(2, 1))
    opt = trainer.FusedAdam(m, lr=1e-2, bert_lr=1e-3)
    opt.dyn = torch.zeros(trainer.FusedAdam.DYN_BLOCK)             # CPU stand-ins for the device block and its pinned mirror
    opt._dyn_host = torch.zeros(trainer.FusedAdam.DYN_BLOCK)
    sched = trainer.LinearWarmupSchedule(opt, 2, 6)
    opt.grad_scale = 0.25
    opt.begin_step()
    assert opt.steps == 1 and opt.dyn[2] == 0.25 and torch.all(opt.dyn[4:8] == 0)          # factor 0 at step 0
    assert abs(float(opt.dyn[0]) - 1.0 / (1 - 0.9)) < 1e-5 and abs(float(opt.dyn[1]) - 1.0 / (1 - 0.999) ** 0.5) < 1e-3
    sched.step()
    opt.begin_step()
    assert torch.allclose(opt.dyn[4:8], torch.full((4,), 5e-3)) and opt.steps == 2
    opt.param_groups[0]["lr"] = 7e-3                               # e.g. load_state_dict
    opt.begin_step()
    assert torch.allclose(opt.dyn[4:8], torch.full((4,), 7e-3))


def test_group_rows_matches_group_lists():
    """evaluate.group_rows / reorder_rows == utils/utils.py:60-80 (_group_lists): rows with one impr_index are concatenated in
    arrival order, groups in first-appearance order."""
    from news_recommendation_mind_b200 import evaluate as ev
    idx = [3, 5, 3, 7, 5, 3]
    cols = [[1., 0.], [0.], [0., 1., 1.], [1.], [1., 0.], [0.]]
    exp = MO.group_by_impression(idx, cols)[0]
    order, g_off = ev.group_rows(torch.tensor(idx))
    assert order is not None and order.tolist() == [0, 2, 5, 1, 4, 3] and g_off.tolist() == [0, 3, 5, 6]
    cnt = torch.tensor([len(c) for c in cols])
    impr = {"offsets": torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(cnt, 0)]),
            "label": torch.tensor([v for c in cols for v in c]), "cdd_id": torch.arange(int(cnt.sum())),
            "impr_index": torch.tensor(idx), "user_id": torch.arange(6)}
    r = ev.reorder_rows(impr, order)
    off = r["offsets"][g_off]
    got = [r["label"][a:b].tolist() for a, b in zip(off[:-1].tolist(), off[1:].tolist())]
    assert got == exp and r["user_id"].tolist() == [0, 2, 5, 1, 4, 3]
    # adjacent chunks (the usual case): no permutation needed
    order, g_off = ev.group_rows(torch.tensor([4, 4, 9, 2, 2, 2]))
    assert order is None and g_off.tolist() == [0, 2, 3, 6]
    order, g_off = ev.group_rows(torch.arange(5))
    assert order is None and g_off.tolist() == [0, 1, 2, 3, 4, 5]
    # partition: no group is cut; rank spans tile the rows
    g_off = torch.tensor([0, 2, 3, 6, 7, 11])
    spans = [ev.group_partition_bounds(g_off, 3, r) for r in range(3)]
    assert spans[0][0] == 0 and spans[-1][1] == 11 and all(spans[i][1] == spans[i + 1][0] for i in range(2))
    assert all(s[0] in g_off.tolist() and s[1] in g_off.tolist() for s in spans)


def test_mrr_ndcg_tie_rule_is_reversed_stable_argsort():
    """Manager.py:1216,1269: order = np.argsort(score)[::-1]; with a stable sort equal scores come out in DESCENDING position."""
    assert MO.mrr(np.array([1, 0]), np.array([.5, .5])) == 0.5
    assert MO.mrr(np.array([0, 1]), np.array([.5, .5])) == 1.0
    y, s = np.array([1., 0., 0., 1.]), np.array([.2, .7, .7, .2])
    order = np.argsort(s, kind="stable")[::-1]
    assert order.tolist() == [2, 1, 3, 0]
    exp = np.sum((2 ** y[order[:2]] - 1) / np.log2(np.arange(2) + 2)) / 1.0
    assert abs(MO.dcg(y, s, 2) - exp) < 1e-15
    assert MO.ordinal_rank(s).tolist() == [3, 1, 2, 4]             # prediction.txt keeps scipy's ordinal rule (ascending position)


def test_id_only_batch_is_the_same_sample():
    ids, mask = data.make_news_table(300, 32, seed=3)
    a = data.make_train_batch(ids, mask, 8, 5, 12, seed=7)
    b = data.make_train_batch(ids, mask, 8, 5, 12, seed=7, id_only=True)
    assert torch.equal(a["cdd_id"], b["cdd_id"].long()) and torch.equal(a["his_id"], b["his_id"].long())
    assert torch.equal(a["his_mask"], b["his_mask"].double()) and torch.equal(a["user_id"], b["user_id"])
    assert torch.equal(ids[b["his_id"].long()], a["his_encoded_index"])
    nbytes = sum(v.numel() * v.element_size() for v in b.values())
    assert nbytes < sum(v.numel() * v.element_size() for v in a.values()) // 50


def test_prediction_file_matches_the_reference_writer(tmp_path):
    """Manager.test (utils/Manager.py:842-850): `index [ranks]` lines, ranks = scipy rankdata(1 - p, 'ordinal')."""
    import scipy.stats as ss
    from news_recommendation_mind_b200 import evaluate as ev
    rng = np.random.RandomState(0)
    preds = [rng.rand(n).round(2) for n in (5, 2, 17, 1, 9)]                  # rounded: ties occur
    offsets = np.concatenate([[0], np.cumsum([len(p) for p in preds])])
    ranks = np.concatenate([ss.rankdata(1 - p, method="ordinal") for p in preds]).astype(np.int32)
    path = tmp_path / "prediction.txt"
    n = ev.write_predictions(str(path), torch.from_numpy(ranks), torch.from_numpy(offsets))
    assert n == 5
    # (the reference's pinned environment -- scipy of the torch-1.9 era -- returns integer ordinal ranks, printed as `3`; scipy
    #  >= 1.10 returns float64 and the same code would print `3.0`.  The MIND submission format and this writer use integers.)
    exp = "".join(str(i + 1) + " [" + ",".join(str(int(r)) for r in ss.rankdata(1 - np.asarray(p), method="ordinal")) + "]\n"
                  for i, p in enumerate(preds))
    assert path.read_text() == exp


def test_fused_adam_state_dict_interchanges_with_torch_adam():
    """Manager.save / load (utils/Manager.py:289-343) go through optimizer.state_dict() / load_state_dict(): the layout
    is torch.optim.Adam's, so a checkpoint of the reference's optimiser resumes here and the other way round."""
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embedding = torch.nn.Module()
            self.embedding.bert_word_embedding = torch.nn.Embedding(7, 4)
            self.encoderN = torch.nn.Linear(4, 3)
    torch.manual_seed(0)
    m = M()
    base = [p for n, p in m.named_parameters() if not re.search("bert", n)]
    bert = [p for n, p in m.named_parameters() if re.search("bert", n)]
    ref = torch.optim.Adam([{"params": base, "lr": 1e-4}, {"params": bert, "lr": 6e-6}])      # Manager._get_optim
    for _ in range(3):
        ref.zero_grad()
        (m.encoderN(m.embedding.bert_word_embedding(torch.tensor([1, 2, 3]))).sum() ** 2).backward()
        ref.step()
    ours = trainer.FusedAdam(m, lr=9.0, bert_lr=9.0)
    ours.load_state_dict(ref.state_dict())
    assert ours.steps == 3 and [g["lr"] for g in ours.param_groups] == [1e-4, 6e-6]
    for p in base + bert:
        assert torch.equal(ours.state[p][0], ref.state[p]["exp_avg"]) and torch.equal(ours.state[p][1], ref.state[p]["exp_avg_sq"])
    sd = ours.state_dict()
    assert [g["params"] for g in sd["param_groups"]] == [g["params"] for g in ref.state_dict()["param_groups"]]
    fresh = torch.optim.Adam([{"params": base, "lr": 1.0}, {"params": bert, "lr": 1.0}])
    fresh.load_state_dict(sd)                                                                   # torch accepts our layout
    for p in base + bert:
        assert torch.equal(fresh.state[p]["exp_avg"], ref.state[p]["exp_avg"])
        assert float(fresh.state[p]["step"]) == 3.0
    assert [g["lr"] for g in fresh.param_groups] == [1e-4, 6e-6]
    with pytest.raises(ValueError):
        ours.load_state_dict({"state": {}, "param_groups": sd["param_groups"][:1]})


def test_linear_warmup_schedule_matches_transformers():
    from transformers import get_linear_schedule_with_warmup
    w = [torch.nn.Parameter(torch.zeros(2)), torch.nn.Parameter(torch.zeros(3))]
    ref_opt = torch.optim.Adam([{"params": [w[0]], "lr": 1e-4}, {"params": [w[1]], "lr": 6e-6}])
    ref = get_linear_schedule_with_warmup(ref_opt, num_warmup_steps=5, num_training_steps=23)
    ours_opt = types.SimpleNamespace(param_groups=[{"lr": 1e-4}, {"lr": 6e-6}])
    ours = trainer.LinearWarmupSchedule(ours_opt, 5, 23)
    for _ in range(30):
        assert [g["lr"] for g in ours_opt.param_groups] == pytest.approx([g["lr"] for g in ref_opt.param_groups], rel=1e-12, abs=0)
        ref_opt.step()
        ref.step()
        ours.step()


def test_dedup_plan_reconstructs_the_slots():
    ids, mask = data.make_news_table(300, 32, seed=3)
    b = data.make_train_batch(ids, mask, 8, 5, 12, seed=7, id_only=True, dedup_capacity=128)
    flat = torch.cat([b["cdd_id"].reshape(-1), b["his_id"].reshape(-1)]).long()
    assert torch.equal(b["uniq_id"].long()[b["uniq_inverse"].long()], flat)
    n = int(torch.unique(flat).numel())
    assert b["uniq_id"].numel() == 128 and torch.all(b["uniq_id"][n:] == 0) and torch.all(b["uniq_id"][1:n] > b["uniq_id"][:n - 1])
    assert data.dedup_plan(b["cdd_id"], b["his_id"], n - 1) is None and data.dedup_plan(b["cdd_id"], b["his_id"])[2] == n


def test_step_scalar_upload_survives_a_host_that_runs_ahead(monkeypatch):
    """FusedAdam's graph mode uploads the step scalars with an asynchronous copy from pinned memory: the DMA reads the host block
    when the copy EXECUTES.  Model of that: a device queue that snapshots the source only when it drains (at an event
    synchronisation or at the end).  20 steps queued without a single synchronisation must each see their OWN bias corrections and
    learning rate -- true with the ring of event-guarded blocks, false with one block (the bug this guards against)."""
    queue, seen = [], []

    class Dev:                                                      # stands in for the CUDA tensor `dyn` and its stream
        device = "fake"

        def copy_(self, h, non_blocking=False):
            queue.append(("copy", h))                               # a reference: nothing is read yet

    def drain(upto=None):
        while queue:
            item = queue.pop(0)
            if item[0] == "copy":
                seen.append(item[1].clone())                        # the DMA runs now
            elif item[1] is upto:
                return

    class Ev:
        def record(self, stream):
            queue.append(("event", self))

        def synchronize(self):
            if any(it[0] == "event" and it[1] is self for it in queue):
                drain(self)

    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: None)
    m = torch.nn.Linear(3, 2)

    def run(slots):
        queue.clear()
        seen.clear()
        opt = trainer.FusedAdam(m, lr=1e-2, bert_lr=1e-3)
        opt.dyn = Dev()
        n = trainer.FusedAdam.DYN_BLOCK
        if slots:
            opt._dyn_ring, opt._dyn_done, opt._dyn_slot = [torch.zeros(n) for _ in range(slots)], [None] * slots, 0
        else:
            opt._dyn_host = torch.zeros(n)                          # the round-1/2 layout: one host block
        sched = trainer.LinearWarmupSchedule(opt, 5, 40)
        for _ in range(20):
            opt.begin_step()
            sched.step()
        drain()
        return [(float(h[0]), float(h[4])) for h in seen]

    want = [(1.0 / (1.0 - 0.9 ** s), 1e-2 * min(s - 1, 5) / 5 if s - 1 < 5 else 1e-2 * (40 - (s - 1)) / 35) for s in range(1, 21)]
    got = run(trainer.FusedAdam.DYN_SLOTS)
    assert len(got) == 20
    for (a, b), (c, d) in zip(got, want):
        assert abs(a - c) <= 1e-6 * c and abs(b - d) <= 1e-9 + 1e-6 * d, (got, want)
    stale = run(0)
    assert len(set(stale)) == 1 and stale[0] != got[0]              # one block: every queued copy read the LAST step's values

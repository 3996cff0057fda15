"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and against the golden
vectors generated from the reference.  Tolerances: fp32 mode 1e-5 relative on logits (north_star);
gradients a little looser because reduction orders differ; integer work bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import build_model, manager_for, model_from_golden, rel_err, random_batch as _random_batch
from oracle import metrics_oracle as MO
from oracle import twotower_oracle as O

pytestmark = pytest.mark.gpu

CASES = [("tt_cnn_lstm", "cnn", "lstm"), ("tt_cnn_gru", "cnn", "gru"), ("tt_cnn_attn", "cnn", "attn"),
         ("tt_cnn_avg", "cnn", "avg"), ("tt_cnn_mha", "cnn", "mha"), ("tt_mha_lstm", "mha", "lstm"),
         ("tt_mha_lstur", "mha", "lstur"), ("tt_cnn_lstur", "cnn", "lstur"), ("tt_cnn_lstm_L32", "cnn", "lstm")]


def _inject_dropout(model, g):
    ex = g.get("extra", {})
    if "drop_keep_cdd" not in ex:
        return
    enc = model.encoderN
    calls = {"n": 0}
    orig = enc.forward

    def fwd(emb, mask=None):
        enc.keep_override = ex["drop_keep_cdd"] if calls["n"] % 2 == 0 else ex["drop_keep_his"]
        calls["n"] += 1
        return orig(emb, mask)
    enc.forward = fwd


@pytest.mark.parametrize("name,encn,encu", CASES)
def test_golden_fp32(name, encn, encu):
    g = load_golden(name)
    model = model_from_golden(g, encn, encu, "fp32")
    _inject_dropout(model, g)
    model.train()
    logp = model(g["x"])[0]
    assert rel_err(logp, g["train_logp"]) < 1e-5
    loss = torch.nn.NLLLoss()(logp, g["x"]["label"].cuda())
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * max(1.0, abs(float(g["loss"])))
    for k, p in model.named_parameters():
        if k in g["grads"]:
            assert p.grad is not None, k
            assert rel_err(p.grad, g["grads"][k]) < 2e-4, (k, rel_err(p.grad, g["grads"][k]))
    # padding row of the token table gets exactly zero gradient
    assert float(model.embedding.weight.grad[0].abs().max()) == 0.0
    model.eval()
    with torch.no_grad():
        prob = model(g["x"])[0]
        assert rel_err(prob, g["eval_prob"]) < 1e-5
        assert rel_err(model.encode_news(g["x"]), g["cdd_repr"]) < 1e-5
        assert rel_err(model.encode_user(g["x"])[0], g["user_repr"]) < 1e-5


def test_cnn_module_standalone_fp32():
    g = load_golden("module_cnn")
    man = manager_for("cnn", "lstm", 3, 3, 10, 16, 12, 4, precision="fp32")
    import news_recommendation_mind_b200 as mr
    enc = mr.CNN_Encoder(man)
    enc.load_state_dict(g["params"])
    enc = enc.cuda()
    emb = g["emb"].cuda().requires_grad_(True)
    c, news = enc(emb, g["mask"].cuda())
    assert rel_err(c, g["c"]) < 1e-5 and rel_err(news, g["news"]) < 1e-5
    assert float(news[0, 0].abs().max()) == 0.0                      # all-masked title -> exact zeros
    (news * g["wn"].cuda()).sum().backward()
    assert rel_err(emb.grad, g["d_emb"]) < 1e-4
    for k, p in enc.named_parameters():
        assert rel_err(p.grad, g["grads"][k]) < 1e-4, k


@pytest.mark.parametrize("encu", ["lstm", "gru"])
def test_mind_small_shape_fp32_vs_oracle(encu):
    """MIND-small shape of BASELINE config 1 (title 32, his 50, npratio 4, 300d/150) at a batch the
    oracle finishes in seconds."""
    gen = torch.Generator().manual_seed(42)
    torch.manual_seed(42)
    B, C, S, L, E, H, V = 8, 5, 50, 32, 300, 150, 3000
    man = manager_for("cnn", encu, C, S, L, E, H, 10, precision="fp32")
    model = build_model(man, V)
    with torch.no_grad():
        model.embedding.weight.normal_(0, 0.3)
    x = _random_batch(gen, B, C, S, L, V)
    params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ref = O.forward(params, x, True, encoder_n="cnn", encoder_u=encu)
    O.nll_loss(ref, x["label"]).backward()
    model.train()
    logp = model(x)[0]
    torch.nn.NLLLoss()(logp, x["label"].cuda()).backward()
    assert rel_err(logp, ref) < 1e-5
    for k, p in model.named_parameters():
        assert rel_err(p.grad, params[k].grad) < 5e-4, (k, rel_err(p.grad, params[k].grad))


def test_embed_grad_bit_reproducible_and_exact_order():
    from news_recommendation_mind_b200 import ops
    gen = torch.Generator().manual_seed(3)
    T, E, V = 20000, 300, 1000
    ids = torch.randint(0, V, (T,), generator=gen)
    ids[:3000] = 7                                       # a long segment (multi-chunk path)
    d = torch.randn(T, E, generator=gen)
    a = ops.embed_grad(ids.cuda(), d.cuda(), V, E, 0)
    b = ops.embed_grad(ids.cuda().int(), d.cuda(), V, E, 0)
    assert torch.equal(a, b)                             # int32 / int64 ids, run-to-run: identical bits
    ref = torch.zeros(V, E, dtype=torch.float64).index_add_(0, ids, d.double())
    ref[0] = 0
    assert float(a[0].abs().max()) == 0.0
    assert rel_err(a, ref) < 1e-6
    # rows that never occur are exact zeros
    unused = torch.ones(V, dtype=torch.bool); unused[ids] = False
    assert float(a[unused.cuda()].abs().max()) == 0.0 if unused.any() else True


def test_embedding_gather_bit_exact():
    from news_recommendation_mind_b200 import ops
    gen = torch.Generator().manual_seed(5)
    table = torch.randn(500, 300, generator=gen)
    ids = torch.randint(0, 500, (7, 3, 32), generator=gen)
    out = ops.EmbeddingGather.apply(ids.cuda(), table.cuda(), 0)
    assert torch.equal(out.cpu(), table[ids])


def test_rank_metrics_match_reference_bit_exact_ranks():
    from news_recommendation_mind_b200 import ops
    g = load_golden("metrics")
    metrics, rank = ops.rank_metrics(g["preds"].float().cuda(), g["labels"].cuda(), g["offsets"].cuda(), want_rank=True)
    # ranks: feed the SAME float32 score bits to the oracle
    offs = g["offsets"].numpy()
    p32 = g["preds"].float().numpy()
    lab = g["labels"].numpy()
    exp_rank = np.concatenate([MO.ordinal_rank(p32[a:b]) for a, b in zip(offs[:-1], offs[1:])])
    assert np.array_equal(rank.cpu().numpy(), exp_rank)
    res = MO.ranking_metrics([lab[a:b] for a, b in zip(offs[:-1], offs[1:])], [p32[a:b] for a, b in zip(offs[:-1], offs[1:])])
    got = metrics.mean(0).cpu().numpy()
    assert [round(float(v), 4) for v in got] == [res["auc"], res["mean_mrr"], res["ndcg@5"], res["ndcg@10"]]
    np.testing.assert_array_equal(np.round(got, 4), g["result"].numpy())    # = reference cal_metric output


def test_rank_metrics_ties_and_large_impression():
    from news_recommendation_mind_b200 import ops
    rng = np.random.default_rng(0)
    ns = [2, 5, 300, 2500]
    labels, preds = [], []
    for n in ns:
        y = (rng.random(n) < 0.2).astype(np.float32); y[0] = 1; y[1] = 0
        p = np.round(rng.random(n), 2).astype(np.float32)           # plenty of ties
        labels.append(y); preds.append(p)
    offs = np.cumsum([0] + ns)
    m, rank = ops.rank_metrics(torch.from_numpy(np.concatenate(preds)).cuda(), torch.from_numpy(np.concatenate(labels)).cuda(),
                               torch.from_numpy(offs).cuda(), want_rank=True)
    exp_rank = np.concatenate([MO.ordinal_rank(p) for p in preds])
    assert np.array_equal(rank.cpu().numpy(), exp_rank)
    for i, (y, p) in enumerate(zip(labels, preds)):
        exp = [MO.auc(y, p), MO.mrr(y, p), MO.ndcg(y, p, 5), MO.ndcg(y, p, 10)]
        np.testing.assert_allclose(m[i].cpu().numpy(), exp, rtol=1e-12, atol=1e-12)


def test_adam_matches_oracle():
    from news_recommendation_mind_b200 import ops
    gen = torch.Generator().manual_seed(9)
    p = torch.randn(1000, 30, generator=gen); g = torch.randn(1000, 30, generator=gen) * 1e-3
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    pc, mc, vc = p.cuda().clone(), m.cuda().clone(), v.cuda().clone()
    shadow = torch.zeros(1000, 64, dtype=torch.bfloat16, device="cuda")
    for step in range(1, 4):
        O.adam_step(p, g * step, m, v, step, 1e-2)
        ops.adam_step(pc, (g * step).cuda(), mc, vc, step, 1e-2, shadow=shadow)
    assert rel_err(pc, p) < 1e-6 and rel_err(mc, m) < 1e-6 and rel_err(vc, v) < 1e-6
    assert torch.equal(shadow[:, :30], pc.to(torch.bfloat16)) and float(shadow[:, 30:].abs().max()) == 0.0


def test_fused_multi_tensor_adam_matches_oracle():
    """trainer.FusedAdam.step = one mr_adam_step_multi launch for all parameters, two LR groups (Manager.py:389-413),
    vector (n % 4 == 0) and scalar tails, bf16 shadow of the table refreshed in the same launch."""
    import types
    import torch.nn as nn
    from news_recommendation_mind_b200 import trainer
    gen = torch.Generator().manual_seed(11)
    shapes = {"embedding.bert_word_embedding.weight": (257, 300), "encoderN.cnn.weight": (150, 300, 3), "encoderN.cnn.bias": (150,),
              "encoderN.odd": (7, 3), "encoderU.w": (601,)}

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = nn.ParameterDict({k.replace(".", "_"): nn.Parameter(torch.randn(*sh, generator=gen)) for k, sh in shapes.items()})

        def named_parameters(self, *a, **k):
            for (name, _), p in zip(shapes.items(), self.ps.values()):
                yield name, p
    mod = M().cuda()
    opt = trainer.FusedAdam(mod, lr=1e-2, bert_lr=3e-3)
    shadow = torch.zeros(257, 320, dtype=torch.bfloat16, device="cuda")
    table = dict(mod.named_parameters())["embedding.bert_word_embedding.weight"]
    opt.embedding = types.SimpleNamespace(weight=table, shadow_bf16=lambda: shadow, mark_shadow_fresh=lambda s: None)
    opt._want_shadow = True
    ref = {n: (p.detach().cpu().clone(), torch.zeros(p.shape), torch.zeros(p.shape)) for n, p in mod.named_parameters()}
    for step in range(1, 4):
        for n, p in mod.named_parameters():
            g = torch.randn(p.shape, generator=gen) * 1e-2
            p.grad = g.cuda()
            O.adam_step(ref[n][0], g, ref[n][1], ref[n][2], step, 3e-3 if "bert" in n else 1e-2)
        opt.step()
    for n, p in mod.named_parameters():
        assert rel_err(p, ref[n][0]) < 1e-6, n
        assert rel_err(opt.state[p][0], ref[n][1]) < 1e-6 and rel_err(opt.state[p][1], ref[n][2]) < 1e-6, n
    assert torch.equal(shadow[:, :300], table.detach().to(torch.bfloat16)) and float(shadow[:, 300:].abs().max()) == 0.0


def test_predict_fast_and_news_table():
    g = load_golden("tt_cnn_lstm")
    model = model_from_golden(g, "cnn", "lstm", "fp32").eval()
    B, C, S, L, E, H, V, hn = [int(v) for v in g["meta"]]
    n_news = 30
    gen = torch.Generator().manual_seed(1)
    table = torch.randn(n_news, H, generator=gen)
    model.init_embedding(table)
    with torch.no_grad():
        got = model.predict_fast(g["x"])
    params = {k: v for k, v in g["params"].items()}
    exp = O.predict_fast(params, table, g["x"], encoder_n="cnn", encoder_u="lstm")
    assert rel_err(got, exp) < 1e-5
    model.destroy_embedding()


def test_bad_shapes_raise():
    import news_recommendation_mind_b200 as mr
    man = manager_for("cnn", "lstm", 3, 3, 10, 16, 12, 4, precision="fp32")
    enc = mr.CNN_Encoder(man).cuda()
    with pytest.raises(ValueError):
        enc(torch.zeros(2, 3, 10, 17, device="cuda"))
    with pytest.raises(AssertionError):
        mr.MHA_Encoder(manager_for("mha", "lstm", 3, 3, 10, 300, 150, 12))      # 150 % 12 != 0, as the reference asserts


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fast_evaluation_metrics_match_oracle(precision):
    """Manager._eval_fast semantics (Manager.py:473-541): news table from encode_news over the whole news set,
    per-impression predict_fast, then cal_metric.  AUC / MRR / nDCG@5 / nDCG@10 must equal the oracle's to 4
    decimals (north_star); in fp32 the scores themselves agree to 1e-5."""
    from news_recommendation_mind_b200 import data, evaluate as ev
    torch.manual_seed(11)
    C, S, L, E, H, V = 5, 12, 32, 300, 150, 30522
    man = manager_for("cnn", "lstm", C, S, L, E, H, 10, precision=precision)
    model = build_model(man, V).eval()
    with torch.no_grad():
        model.embedding.weight.normal_(0, 0.3)
    news_ids, news_mask = data.make_news_table(400, L, seed=5)
    impr = data.make_eval_impressions(news_ids, news_mask, 60, S, seed=9)
    got = ev.evaluate(model, news_ids, news_mask, impr)
    # oracle: same pipeline on the CPU
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        table = O.encode_news(params, news_ids.unsqueeze(1), news_mask.unsqueeze(1), "cnn").squeeze(1)
        offs = impr["offsets"].numpy()
        labels, preds = [], []
        for i in range(len(offs) - 1):
            x = {"cdd_id": impr["cdd_id"][offs[i]:offs[i + 1]].unsqueeze(0), "his_encoded_index": impr["his_encoded_index"][i:i + 1],
                 "his_attn_mask": impr["his_attn_mask"][i:i + 1], "his_mask": impr["his_mask"][i:i + 1],
                 "user_id": impr["user_id"][i:i + 1]}
            preds.append(O.predict_fast(params, table, x, encoder_n="cnn", encoder_u="lstm").squeeze(0).numpy())
            labels.append(impr["label"][offs[i]:offs[i + 1]].numpy())
    exp = MO.ranking_metrics(labels, preds)
    print(precision, got, exp)
    if precision == "fp32":
        assert got == {k: exp[k] for k in got}
    else:   # bf16 scores differ in the 3rd significant digit; the rank-based means move by at most a few 1e-3
        for k in got:
            assert abs(got[k] - exp[k]) < 5e-3, (k, got[k], exp[k])


@pytest.mark.parametrize("name,encn,encu,B,C,S,L,precision,tol_logit,tol_grad", [
    # BASELINE configs[2]: CNN news encoder + MHA user encoder, title 32 / his 50 / npratio 4, 300d -> 150, 10 heads
    ("config3_fp32", "cnn", "mha", 12, 5, 50, 32, "fp32", 1e-5, 5e-4),
    ("config3_bf16", "cnn", "mha", 12, 5, 50, 32, "bf16", 1e-3, None),
    # BASELINE configs[4]: MHA news encoder + LSTUR user encoder, title 48 / his 100 / npratio 9
    ("config5_fp32", "mha", "lstur", 3, 10, 100, 48, "fp32", 1e-5, 5e-4),
    ("config5_bf16", "mha", "lstur", 3, 10, 100, 48, "bf16", 1e-3, None),
])
def test_baseline_config_shapes_vs_oracle(name, encn, encu, B, C, S, L, precision, tol_logit, tol_grad):
    """The other BASELINE.json configurations at their real title / history / candidate sizes (small batch so that
    the CPU oracle finishes in seconds): training logits, loss gradients (fp32) and evaluation probabilities against
    the oracle; north_star bounds 1e-5 (fp32) / 1e-3 (bf16) relative on the logits.  Dropout is switched off
    (dropout_p = 0) and the LSTUR Bernoulli draw is injected, so the comparison is deterministic."""
    E, H, V, hn, n_users = 300, 150, 3000, 10, 40
    import zlib
    gen = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)      # (str hash() is salted per process: not a seed)
    torch.manual_seed(17)
    man = manager_for(encn, encu, C, S, L, E, H, hn, precision=precision, n_users=n_users, dropout_p=0.0)
    model = build_model(man, V)
    with torch.no_grad():
        model.embedding.weight.normal_(0, 0.3)
    x = _random_batch(gen, B, C, S, L, V, n_users=n_users)
    keep = None
    if encu == "lstur":
        keep = torch.bernoulli(torch.full((B,), 0.5), generator=gen)
        model.encoderU.keep_user = keep
    params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    kw = dict(encoder_n=encn, encoder_u=encu, head_num=hn, keep_user=keep, dropout_p=0.0)
    ref = O.forward(params, x, True, **kw)
    O.nll_loss(ref, x["label"]).backward()
    model.train()
    logp = model(x)[0]
    torch.nn.NLLLoss()(logp, x["label"].cuda()).backward()
    e = rel_err(logp, ref)
    print("%s: train logits rel err %.3e" % (name, e))
    assert e < tol_logit, e
    if tol_grad is not None:
        # gradients: against a float64 run of the oracle; an fp32 implementation is allowed tol_grad, or -- for the
        # ill-conditioned ones (the pooling queries: softmax backward sums to zero over 50..100 items) -- three times
        # the error the fp32 oracle itself makes against float64
        p64 = {k: v.detach().double().requires_grad_(True) for k, v in params.items()}
        x64 = {k: (v.double() if v.is_floating_point() else v) for k, v in x.items()}
        O.nll_loss(O.forward(p64, x64, True, **kw), x["label"]).backward()
        for k, p in model.named_parameters():
            if p.grad is None or p64[k].grad is None:
                continue
            ge = rel_err(p.grad, p64[k].grad)
            oe = rel_err(params[k].grad, p64[k].grad)
            print("  %-40s cuda %.2e   fp32 oracle %.2e" % (k, ge, oe))
            assert ge < max(tol_grad, 3 * oe), (k, ge, oe)
    assert float(model.embedding.weight.grad[0].abs().max()) == 0.0
    model.eval()
    with torch.no_grad():
        prob = model(x)[0]
        eref = O.forward(params, x, False, **kw)
    assert rel_err(prob, eref) < tol_logit

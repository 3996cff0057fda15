"""Pins the oracle (oracle/*.py) to the real reference: every case below was
produced by oracle/make_golden.py running /root/reference itself."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO
from oracle import twotower_oracle as O
from conftest import load_golden

CASES = [("tt_cnn_lstm", "cnn", "lstm"), ("tt_cnn_gru", "cnn", "gru"), ("tt_cnn_attn", "cnn", "attn"),
         ("tt_cnn_avg", "cnn", "avg"), ("tt_cnn_mha", "cnn", "mha"), ("tt_mha_lstm", "mha", "lstm"),
         ("tt_mha_lstur", "mha", "lstur"), ("tt_cnn_lstur", "cnn", "lstur"),
         ("tt_cnn_lstm_L32", "cnn", "lstm")]


def _kw(g, encn, encu):
    hn = int(g["meta"][7])
    kw = dict(encoder_n=encn, encoder_u=encu, head_num=hn)
    ex = g.get("extra", {})
    if "keep_user" in ex:
        kw["keep_user"] = ex["keep_user"]
    return kw, ex


@pytest.mark.parametrize("name,encn,encu", CASES)
def test_forward_and_grads_match_reference(name, encn, encu):
    g = load_golden(name)
    kw, ex = _kw(g, encn, encu)
    params = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    tkw = dict(kw)
    if encn == "mha":
        tkw.update(drop_keep_cdd=ex["drop_keep_cdd"], drop_keep_his=ex["drop_keep_his"], dropout_p=0.2)
    logp = O.forward(params, g["x"], True, **tkw)
    torch.testing.assert_close(logp, g["train_logp"], rtol=1e-5, atol=1e-6)
    loss = O.nll_loss(logp, g["x"]["label"])
    torch.testing.assert_close(loss, g["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    for k, ref in g["grads"].items():
        assert params[k].grad is not None, k
        torch.testing.assert_close(params[k].grad, ref, rtol=2e-4, atol=2e-6, msg=lambda m: k + ": " + m)
    # parameters the reference leaves without a gradient must have none here either
    for k, p in params.items():
        if k not in g["grads"]:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    with torch.no_grad():
        prob = O.forward(params, g["x"], False, **kw)
        torch.testing.assert_close(prob, g["eval_prob"], rtol=1e-5, atol=1e-6)
        cdd = O.encode_news(params, g["x"]["cdd_encoded_index"], g["x"]["cdd_attn_mask"], encn, kw["head_num"])
        torch.testing.assert_close(cdd, g["cdd_repr"], rtol=1e-5, atol=1e-6)
        ukw = {k: v for k, v in kw.items()}
        user = O.encode_user(params, g["x"], **ukw)
        torch.testing.assert_close(user, g["user_repr"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,encn,encu", [("cfg_cnn_lstm", "cnn", "lstm"), ("cfg_cnn_mha", "cnn", "mha"),
                                            ("cfg_mha_lstur", "mha", "lstur")])
def test_baseline_config_sizes_match_reference(name, encn, encu):
    """The oracle against the REAL reference at the sizes of BASELINE.json's configurations (title 32 / his 50 / npratio 4
    and title 48 / his 100 / npratio 9; 300d -> 150, 10 heads; small batch and vocabulary).  The weights are rebuilt from
    numpy (oracle/make_golden.py::np_params) from the (name, shape) list in the fixture; the fixture holds the reference's
    outputs and, per parameter, sum / sum|.| / L2 norm / first 8 entries of its gradient."""
    from oracle.make_golden import grad_stats, np_params
    g = load_golden(name)
    B, C, S, L, E, H, V, hn, seed = [int(v) for v in g["meta"]]
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    shapes = {str(n): tuple(int(v) for v in shp[:nd]) for n, shp, nd in zip(z["param_names"], z["param_shapes"], z["param_ndim"])}
    params = {k: v.clone().requires_grad_(True) for k, v in np_params(shapes, seed).items()}
    kw = dict(encoder_n=encn, encoder_u=encu, head_num=hn)
    if "keep_user" in g.get("extra", {}):
        kw["keep_user"] = g["extra"]["keep_user"]
    logp = O.forward(params, g["x"], True, dropout_p=0.0, **kw)
    torch.testing.assert_close(logp, g["train_logp"], rtol=1e-5, atol=1e-6)
    loss = O.nll_loss(logp, g["x"]["label"])
    torch.testing.assert_close(loss, g["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    assert len(g["grad_stats"]) >= 6
    for k, ref in g["grad_stats"].items():
        assert params[k].grad is not None, k
        got = grad_stats(params[k].grad)
        scale = float(ref[2])                                     # L2 norm of the reference gradient
        assert float((got - ref).abs().max()) <= 2e-4 * max(scale, float(ref[1])) + 1e-7, (k, got[:3], ref[:3])
    for k, p in params.items():
        if k not in g["grad_stats"]:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
    with torch.no_grad():
        torch.testing.assert_close(O.forward(params, g["x"], False, **kw), g["eval_prob"], rtol=1e-5, atol=1e-6)
        cdd = O.encode_news(params, g["x"]["cdd_encoded_index"], g["x"]["cdd_attn_mask"], encn, hn)
        torch.testing.assert_close(cdd, g["cdd_repr"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(O.encode_user(params, g["x"], **kw), g["user_repr"], rtol=1e-5, atol=1e-6)


def test_padding_row_gets_no_gradient():
    g = load_golden("tt_cnn_lstm")
    assert float(g["grads"]["embedding.bert_word_embedding.weight"][0].abs().max()) == 0.0
    assert float(g["params"]["embedding.bert_word_embedding.weight"][0].abs().max()) > 0.0


def test_adam_trajectory_matches_reference():
    g = load_golden("tt_cnn_lstm")
    params = {k: v.clone() for k, v in g["params"].items()}
    state = {}
    losses = []
    for s in range(len(g["traj_losses"])):
        xb = {k.split("/", 1)[1]: v for k, v in g["traj_batches"].items() if k.startswith("step%d/" % s)}
        loss, _ = O.train_step(params, state, xb, s + 1, lr=1e-2, bert_lr=3e-3, encoder_n="cnn", encoder_u="lstm")
        losses.append(loss)
    np.testing.assert_allclose(losses, g["traj_losses"].numpy(), rtol=1e-5)
    for k, ref in g["traj_params"].items():
        # Adam divides by sqrt(v): where a gradient is ~0 its fp32 rounding noise is amplified to a
        # fraction of lr (1e-2 here), hence the absolute tolerance of a few % of one step
        torch.testing.assert_close(params[k], ref, rtol=1e-4, atol=3e-4, msg=lambda m: k + ": " + m)


def test_cnn_module_matches_reference():
    g = load_golden("module_cnn")
    emb = g["emb"].clone().requires_grad_(True)
    P = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    c, news = O.cnn_news_encoder(emb, g["mask"], P["cnn.weight"], P["cnn.bias"], P["wordQueryProject.weight"],
                                 P["wordQueryProject.bias"], P["query_words"])
    torch.testing.assert_close(c, g["c"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(news, g["news"], rtol=1e-5, atol=1e-6)
    assert float(news[0, 0].abs().max()) == 0.0          # all-masked title -> exact zeros
    (news * g["wn"]).sum().backward()
    torch.testing.assert_close(emb.grad, g["d_emb"], rtol=1e-4, atol=1e-6)
    for k, ref in g["grads"].items():
        torch.testing.assert_close(P[k].grad, ref, rtol=1e-4, atol=1e-6, msg=lambda m: k + ": " + m)


def test_masked_softmax_known_answers():
    g = load_golden("xsoftmax")
    s = g["s"].clone().requires_grad_(True)
    p = O.masked_softmax(s, g["m"])
    torch.testing.assert_close(p, g["p"], rtol=1e-6, atol=1e-7)
    assert float(p[1].abs().max()) == 0.0
    (p * g["w"]).sum().backward()
    torch.testing.assert_close(s.grad, g["ds"], rtol=1e-5, atol=1e-7)
    assert float(s.grad[1].abs().max()) == 0.0


def test_metrics_match_reference_cal_metric():
    g = load_golden("metrics")
    offs = g["offsets"].numpy()
    labels = [g["labels"].numpy()[a:b] for a, b in zip(offs[:-1], offs[1:])]
    preds = [g["preds"].numpy()[a:b] for a, b in zip(offs[:-1], offs[1:])]
    res = MO.ranking_metrics(labels, preds)
    got = np.array([res["auc"], res["mean_mrr"], res["ndcg@5"], res["ndcg@10"]])
    np.testing.assert_array_equal(got, g["result"].numpy())
    ranks = np.concatenate([MO.ordinal_rank(p) for p in preds])
    np.testing.assert_array_equal(ranks, g["ranks"].numpy())           # bit-exact impression ranking


def test_auc_known_answer_with_tie():
    # cal_metric([[1,0,0,1,0]], [[.9,.1,.9,.3,.2]]) -> auc 0.75
    assert MO.auc(np.array([1, 0, 0, 1, 0]), np.array([.9, .1, .9, .3, .2])) == 0.75


def test_auc_matches_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(0)
    for _ in range(20):
        n = int(rng.integers(2, 50))
        y = np.zeros(n); y[: max(1, n // 4)] = 1; rng.shuffle(y)
        if y.sum() == n:
            y[0] = 0
        s = np.round(rng.random(n), 1)                                   # many ties
        assert abs(MO.auc(y, s) - sk.roc_auc_score(y, s)) < 1e-12


def test_group_and_partition():
    lab, pred = MO.group_by_impression([3, 5, 3, 7], [[1, 0], [0], [0, 1], [1]], [[.1, .2], [.3], [.4, .5], [.6]])
    assert lab == [[1, 0, 0, 1], [0], [1]] and pred == [[.1, .2, .4, .5], [.3], [.6]]
    assert [MO.partition_bounds(10, 4, r) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 10)]

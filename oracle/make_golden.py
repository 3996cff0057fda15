"""Generate tests/golden/*.npz by running the REAL reference (read-only mount at
/root/reference) on seeded inputs.  Test infrastructure only; run in the build
container (the GPU box has no /root/reference):

    python oracle/make_golden.py

Harness shims -- none of them edits the reference:
  1. models.Modules.Attention._softmax_backward_data is re-bound to the torch-2.x
     signature (the reference calls the torch-1.9 one, Attention.py:79).
  2. BERT_Embedding.__init__ downloads bert-base-uncased (BERT.py:16-19); the
     instance is built with nn.Module.__init__ only and given
     bert_word_embedding = nn.Embedding(V, E, padding_idx=0), the same attribute,
     class and forward().
  3. Intended-semantics adapters for code that is broken as shipped:
     MHA_User_Encoder gets his_mask^T for the pooling step (MHA.py:71 passes the
     un-transposed mask); LSTUR_User_Encoder is called with user_index= and an
     injected Bernoulli draw (twotower.py:44 / TwoTower.py:47 / RNN.py:100-101).
  4. nn.Dropout in MHA_Encoder is replaced by a fixed keep-mask so the draw is
     reproducible (MHA.py:19,37).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("MIND_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF)
    import models.Modules.Attention as A  # noqa
    A._softmax_backward_data = lambda g, o, d, s: torch._softmax_backward_data(g, o, d, s.dtype)
    from models.TwoTower import TwoTower
    from models.Embeddings.BERT import BERT_Embedding
    from models.Encoders.CNN import CNN_Encoder
    from models.Encoders.MHA import MHA_Encoder, MHA_User_Encoder
    from models.Encoders.RNN import RNN_User_Encoder, LSTUR_User_Encoder
    from models.Encoders.Pooling import Attention_Pooling, Average_Pooling
    return types.SimpleNamespace(**locals())


def fake_manager(**kw):
    m = types.SimpleNamespace(scale="demo", mode="train", cdd_size=5, impr_size=2000, batch_size_news=500,
                              his_size=50, signal_length=32, device="cpu", bert_dim=300, hidden_dim=150,
                              head_num=10, dropout_p=0.2, descend_history=False, encoderN="cnn",
                              encoderU="lstm", n_users=40)
    m.__dict__.update(kw)
    m.get_user_num = lambda: m.n_users
    return m


class FixedDropout(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.p = p
        self.keep = None

    def forward(self, x):
        if self.keep is None or not self.training:
            return x
        return x * self.keep.reshape(x.shape).to(x.dtype) / (1.0 - self.p)


def build_reference_model(R, man, V):
    emb = R.BERT_Embedding.__new__(R.BERT_Embedding)
    nn.Module.__init__(emb)
    emb.hidden_dim = man.bert_dim
    emb.bert_word_embedding = nn.Embedding(V, man.bert_dim, padding_idx=0)
    with torch.no_grad():
        emb.bert_word_embedding.weight.normal_(0, 0.3)      # row 0 kept non-zero on purpose
    encN = {"cnn": R.CNN_Encoder, "mha": R.MHA_Encoder}[man.encoderN](man)
    if man.encoderN == "mha":
        encN.dropOut = FixedDropout(man.dropout_p)
    u = man.encoderU
    if u in ("lstm", "gru"):
        encU = R.RNN_User_Encoder(man)
    elif u == "attn":
        encU = R.Attention_Pooling(man)
    elif u == "avg":
        encU = R.Average_Pooling(man)
    elif u == "mha":
        inner = R.MHA_User_Encoder(man)

        class MHAUserIntended(nn.Module):            # shim 3
            def __init__(self):
                super().__init__()
                self.mha = inner.mha
                self.query_news = inner.query_news
                self.layerNorm = inner.layerNorm

            def forward(self, news_repr, his_mask=None, **kw):
                from models.Modules.Attention import get_attn_mask, scaled_dp_attention
                hm = his_mask.to(news_repr.device)
                h = self.mha(news_repr, get_attn_mask(hm.squeeze(-1)))
                return scaled_dp_attention(self.query_news, h, h, hm.transpose(-1, -2))
        encU = MHAUserIntended()
    elif u == "lstur":
        inner = R.LSTUR_User_Encoder(man)
        with torch.no_grad():
            inner.userEmbedding.weight[0].zero_()

        class LSTURIntended(nn.Module):              # shim 3
            def __init__(self):
                super().__init__()
                self.rnn = inner.rnn
                self.userEmbedding = inner.userEmbedding
                self.keep_user = None

            def forward(self, news_repr, his_mask=None, user_id=None, **kw):
                B = news_repr.size(0)
                masked = self.keep_user.to(torch.long) * user_id
                h0 = self.userEmbedding(masked).unsqueeze(0)
                c0 = torch.zeros(1, B, news_repr.size(-1))
                _, st = self.rnn(news_repr.flip(dims=[1]), (h0, c0))
                return st[0].transpose(0, 1)
        encU = LSTURIntended()
    else:
        raise ValueError(u)
    model = R.TwoTower(man, emb, encN, encU)
    return model


def synth_batch(g, B, C, S, L, V, n_users, n_news=30):
    def titles(n):
        ids = torch.zeros(n, L, dtype=torch.int64)
        mask = torch.zeros(n, L, dtype=torch.int64)
        for i in range(n):
            ln = int(torch.randint(2, L + 1, (1,), generator=g))
            ids[i, :ln] = torch.randint(1, V, (ln,), generator=g)
            mask[i, :ln] = 1
        return ids, mask
    cid, cm = titles(B * C)
    hid, hm = titles(B * S)
    his_mask = torch.zeros(B, S, 1, dtype=torch.float64)
    for b in range(B):
        ln = int(torch.randint(0, S + 1, (1,), generator=g)) if b else 0     # sample 0: empty history
        if ln == 0:
            his_mask[b, 0] = 1                                                 # MIND.py:334-335
            hid.view(B, S, L)[b] = 0
            hm.view(B, S, L)[b] = 0
            hid.view(B, S, L)[b, :, 0] = 1
            hid.view(B, S, L)[b, :, 1] = 2
            hm.view(B, S, L)[b, :, :2] = 1
        else:
            his_mask[b, :ln] = 1
    return {
        "cdd_encoded_index": cid.view(B, C, L), "cdd_attn_mask": cm.view(B, C, L),
        "his_encoded_index": hid.view(B, S, L), "his_attn_mask": hm.view(B, S, L),
        "his_mask": his_mask, "user_id": torch.randint(1, n_users + 1, (B,), generator=g),
        "cdd_id": torch.randint(1, n_news, (B, C), generator=g),
        "his_id": torch.randint(1, n_news, (B, S), generator=g),
        "label": torch.randint(0, C, (B,), generator=g),
    }


def dump(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    flat = {}
    for k, v in arrays.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[k + "/" + kk] = vv.detach().cpu().numpy() if torch.is_tensor(vv) else np.asarray(vv)
        else:
            flat[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **flat)
    print("wrote", name, sum(a.size for a in flat.values()), "elements")


def model_case(R, name, encN, encU, B, C, S, L, E, H, V, hn, seed, adam_steps=0):
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    man = fake_manager(encoderN=encN, encoderU=encU, cdd_size=C, his_size=S, signal_length=L, bert_dim=E,
                       hidden_dim=H, head_num=hn)
    model = build_reference_model(R, man, V)
    x = synth_batch(g, B, C, S, L, V, man.n_users)
    extra = {}
    if encU == "lstur":
        keep = torch.randint(0, 2, (B,), generator=g)
        model.encoderU.keep_user = keep
        extra["keep_user"] = keep
    if encN == "mha":
        kc = (torch.rand(B * C, L, H, generator=g) >= man.dropout_p)
        kh = (torch.rand(B * S, L, H, generator=g) >= man.dropout_p)
        extra["drop_keep_cdd"], extra["drop_keep_his"] = kc, kh
        calls = {"n": 0}
        drop = model.encoderN.dropOut
        orig = drop.forward

        def fwd(t):                                   # first call = candidates, second = history
            drop.keep = kc if calls["n"] % 2 == 0 else kh
            calls["n"] += 1
            return orig(t)
        drop.forward = fwd
    params0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    # train-mode forward + grads (Manager.py:636-644)
    model.train()
    logp = model(x)[0]
    loss = nn.NLLLoss()(logp, x["label"])
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    # eval-mode forward (sigmoid)
    model.eval()
    with torch.no_grad():
        prob = model(x)[0]
        cdd_repr = model.encode_news(x)
        user_repr = model.encode_user(x)[0]
    out = dict(x=x, params=params0, grads=grads, extra=extra, train_logp=logp.detach(), loss=loss.detach(),
               eval_prob=prob, cdd_repr=cdd_repr, user_repr=user_repr,
               meta=np.array([B, C, S, L, E, H, V, hn], dtype=np.int64))
    if adam_steps:
        # trajectory: Manager._get_optim groups (Manager.py:396-413) lr 1e-2 / bert_lr 3e-3 (large so
        # that the steps are visible in fp32), fresh batches each step
        model.train()
        base, bert = [], []
        for n_, p in model.named_parameters():
            (bert if "bert" in n_ else base).append(p)
        opt = torch.optim.Adam([{"params": base, "lr": 1e-2}, {"params": bert, "lr": 3e-3}])
        for p in model.parameters():
            p.grad = None
        losses = []
        batches = {}
        for s in range(adam_steps):
            xb = synth_batch(g, B, C, S, L, V, man.n_users)
            for k, v in xb.items():
                batches["step%d/%s" % (s, k)] = v
            opt.zero_grad(set_to_none=True)
            l = nn.NLLLoss()(model(xb)[0], xb["label"])
            l.backward()
            opt.step()
            losses.append(float(l))
        out["traj_batches"] = batches
        out["traj_losses"] = np.array(losses)
        out["traj_params"] = {k: v.detach().clone() for k, v in model.state_dict().items()}
    dump(name, **out)


def module_cases(R):
    """Stand-alone encoder modules, forward + input/weight grads."""
    torch.manual_seed(7)
    g = torch.Generator().manual_seed(7)
    N, L, E, H = 6, 10, 16, 12
    man = fake_manager(bert_dim=E, hidden_dim=H, head_num=4)
    enc = R.CNN_Encoder(man)
    emb = torch.randn(2, 3, L, E, generator=g, requires_grad=True)
    mask = (torch.rand(2, 3, L, generator=g) > 0.3).long()
    mask[0, 0] = 0                                         # an all-masked title -> zero vector
    c, news = enc(emb, mask)
    wn = torch.randn(news.shape, generator=g)
    (news * wn).sum().backward()
    dump("module_cnn", emb=emb, mask=mask, c=c, news=news, wn=wn, d_emb=emb.grad,
         params={k: v for k, v in enc.state_dict().items()},
         grads={k: p.grad for k, p in enc.named_parameters()})
    # XSoftmax known answers incl. an all-masked row (Attention.py:56-80)
    from models.Modules.Attention import XSoftmax, scaled_dp_attention, get_attn_mask
    s = torch.randn(3, 5, generator=g, requires_grad=True)
    m = torch.tensor([[1, 1, 0, 1, 0], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1]])
    p = XSoftmax.apply(s, m, -1)
    w = torch.randn(3, 5, generator=g)
    (p * w).sum().backward()
    dump("xsoftmax", s=s, m=m, p=p, w=w, ds=s.grad)


def metric_cases():
    sys.path.insert(0, REF)
    from utils.Manager import cal_metric
    import scipy.stats as ss
    rng = np.random.default_rng(11)
    labels, preds, ranks = [], [], []
    for _ in range(40):
        n = int(rng.integers(2, 60))
        y = (rng.random(n) < 0.15).astype(np.float64)
        y[int(rng.integers(0, n))] = 1.0
        if y.sum() == n:
            y[0] = 0.0
        p = rng.permutation(n).astype(np.float64) / n + rng.random() * 1e-3     # tie-free
        labels.append(y)
        preds.append(1.0 / (1.0 + np.exp(-p)))
        ranks.append(ss.rankdata(1 - preds[-1], method="ordinal"))             # Manager.py:846
    res = cal_metric([l.tolist() for l in labels], [p.tolist() for p in preds],
                     ["auc", "mean_mrr", "ndcg@5", "ndcg@10"])
    offs = np.cumsum([0] + [len(l) for l in labels])
    dump("metrics", offsets=offs, labels=np.concatenate(labels), preds=np.concatenate(preds),
         ranks=np.concatenate(ranks).astype(np.int64),
         result=np.array([res["auc"], res["mean_mrr"], res["ndcg@5"], res["ndcg@10"]]))
    # the survey's known-answer vector (contains a tie; AUC is tie-order independent)
    ka = cal_metric([[1, 0, 0, 1, 0]], [[.9, .1, .9, .3, .2]], ["auc"])
    assert ka["auc"] == 0.75


def np_params(shapes, seed):
    """Deterministic parameters from numpy alone (independent of torch's RNG stream, so that the test can rebuild them
    from the (name, shape) list stored in the fixture): N(0, 0.3^2) for the token table, N(0, 0.1^2) for matrices,
    N(0, 0.05^2) for vectors, the LayerNorm weight around 1."""
    rng = np.random.RandomState(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(int(v) for v in shapes[name])
        scale = 0.3 if "bert_word_embedding" in name else (0.1 if len(shp) > 1 else 0.05)
        a = rng.normal(0.0, scale, size=shp).astype(np.float32)
        if name.endswith("layerNorm.weight"):
            a = a + 1.0
        if name.endswith("userEmbedding.weight"):
            a[0] = 0.0                                   # RNN.py:81-82: row 0 of the user table is zero
        out[name] = torch.from_numpy(a)
    return out


def grad_stats(t):
    """what the fixture keeps of a (large) gradient: sum, sum of |.|, L2 norm and the first 8 entries"""
    f = t.detach().double().reshape(-1)
    head = torch.zeros(8, dtype=torch.float64)
    head[: min(8, f.numel())] = f[:8]
    return torch.cat([torch.stack([f.sum(), f.abs().sum(), f.norm()]), head])


def config_case(R, name, encN, encU, B, C, S, L, seed, E=300, H=150, V=300, hn=10):
    """The BASELINE.json configurations at their real title / history / candidate / embedding / hidden sizes (small
    batch and vocabulary).  The parameters are rebuilt from numpy in the test, so the fixture holds the inputs, the
    outputs and per-parameter gradient statistics only (a full copy of the weights and gradients would be megabytes)."""
    g = torch.Generator().manual_seed(seed)
    man = fake_manager(encoderN=encN, encoderU=encU, cdd_size=C, his_size=S, signal_length=L, bert_dim=E,
                       hidden_dim=H, head_num=hn, dropout_p=0.0)
    model = build_reference_model(R, man, V)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(np_params(shapes, seed))
    x = synth_batch(g, B, C, S, L, V, man.n_users)
    extra = {}
    if encU == "lstur":
        keep = torch.randint(0, 2, (B,), generator=g)
        model.encoderU.keep_user = keep
        extra["keep_user"] = keep
    model.train()
    logp = model(x)[0]
    loss = nn.NLLLoss()(logp, x["label"])
    loss.backward()
    gstats = {k: grad_stats(p.grad) for k, p in model.named_parameters() if p.grad is not None}
    model.eval()
    with torch.no_grad():
        prob = model(x)[0]
        cdd_repr = model.encode_news(x)
        user_repr = model.encode_user(x)[0]
    names = sorted(shapes)
    dump(name, x=x, extra=extra, train_logp=logp.detach(), loss=loss.detach(), eval_prob=prob, cdd_repr=cdd_repr,
         user_repr=user_repr, grad_stats=gstats,
         param_names=np.array(names), param_shapes=np.array([list(shapes[n]) + [0] * (3 - len(shapes[n])) for n in names], dtype=np.int64),
         param_ndim=np.array([len(shapes[n]) for n in names], dtype=np.int64),
         meta=np.array([B, C, S, L, E, H, V, hn, seed], dtype=np.int64))


def main():
    R = _import_reference()
    if "--configs-only" in sys.argv:
        config_cases(R)
        return
    module_cases(R)
    metric_cases()
    small = dict(B=3, C=4, S=5, L=8, E=24, H=12, V=60, hn=3)
    model_case(R, "tt_cnn_lstm", "cnn", "lstm", seed=1, adam_steps=3, **small)
    model_case(R, "tt_cnn_gru", "cnn", "gru", seed=2, **small)
    model_case(R, "tt_cnn_attn", "cnn", "attn", seed=3, **small)
    model_case(R, "tt_cnn_avg", "cnn", "avg", seed=4, **small)
    model_case(R, "tt_cnn_mha", "cnn", "mha", seed=5, **small)
    model_case(R, "tt_mha_lstm", "mha", "lstm", seed=6, **small)
    model_case(R, "tt_mha_lstur", "mha", "lstur", seed=8, **small)
    model_case(R, "tt_cnn_lstur", "cnn", "lstur", seed=9, **small)
    # a MIND-small-shaped (but narrow) case: L=32, npratio 4
    model_case(R, "tt_cnn_lstm_L32", "cnn", "lstm", seed=10, B=4, C=5, S=6, L=32, E=40, H=20, V=120, hn=5)
    config_cases(R)


def config_cases(R):
    # BASELINE.json configs 1-2 (CNN + LSTM, title 32 / his 50 / npratio 4, 300d -> 150), 3 (CNN + MHA user), 5 (MHA news +
    # LSTUR, title 48 / his 100 / npratio 9)
    config_case(R, "cfg_cnn_lstm", "cnn", "lstm", B=2, C=5, S=50, L=32, seed=21)
    config_case(R, "cfg_cnn_mha", "cnn", "mha", B=2, C=5, S=50, L=32, seed=22)
    config_case(R, "cfg_mha_lstur", "mha", "lstur", B=2, C=10, S=100, L=48, seed=23)


if __name__ == "__main__":
    main()

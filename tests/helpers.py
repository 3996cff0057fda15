"""Shared helpers for the parity tests: build the CUDA model from a state dict / golden file."""
import types

import torch

import news_recommendation_mind_b200 as mr


def manager_for(encn, encu, C, S, L, E, H, hn, precision="fp32", device="cuda:0", n_users=40, dropout_p=0.2,
                descend_history=False):
    m = types.SimpleNamespace(scale="demo", mode="train", cdd_size=C, impr_size=2000, batch_size_news=500, his_size=S,
                              signal_length=L, device=device, bert_dim=E, hidden_dim=H, head_num=hn,
                              dropout_p=dropout_p, descend_history=descend_history, encoderN=encn, encoderU=encu,
                              precision=precision, n_users=n_users)
    m.get_user_num = lambda: m.n_users
    return m


def build_model(man, V, state=None):
    emb = mr.BERT_Embedding(man, vocab_size=V)
    encN = {"cnn": mr.CNN_Encoder, "mha": mr.MHA_Encoder}[man.encoderN](man)
    u = man.encoderU
    encU = {"lstm": mr.RNN_User_Encoder, "gru": mr.RNN_User_Encoder, "attn": mr.Attention_Pooling,
            "avg": mr.Average_Pooling, "mha": mr.MHA_User_Encoder, "lstur": mr.LSTUR_User_Encoder}[u](man)
    model = mr.TwoTower(man, emb, encN, encU)
    if state is not None:
        missing, unexpected = model.load_state_dict(state, strict=False)
        # the reference's MHA user wrapper in make_golden does not carry dropOut (no params) -- nothing may be missing
        assert not unexpected, unexpected
        assert all("layerNorm" in k for k in missing) or not missing, missing
    return model.to(man.device)


def model_from_golden(g, encn, encu, precision="fp32"):
    B, C, S, L, E, H, V, hn = [int(v) for v in g["meta"]]
    man = manager_for(encn, encu, C, S, L, E, H, hn, precision=precision)
    model = build_model(man, V, {k: v.clone() for k, v in g["params"].items()})
    ex = g.get("extra", {})
    if "keep_user" in ex:
        model.encoderU.keep_user = ex["keep_user"]
    return model


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_err(a: torch.Tensor, b: torch.Tensor):
    """element-wise bounds beside the norm ratio of rel_err: (max |a-b|, max |a-b| / max(|b|, floor)) with the floor at 1e-3 of
    the largest |b| -- a single bad row shows up here even when the Frobenius ratio hides it"""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    d = (a - b).abs()
    floor = 1e-3 * float(b.abs().max()) + 1e-30
    return float(d.max()), float((d / b.abs().clamp_min(floor)).max())


def random_batch(gen, B, C, S, L, V, n_users=40):
    def titles(n):
        ln = torch.randint(2, L + 1, (n,), generator=gen)
        ids = torch.randint(1, V, (n, L), generator=gen)
        mask = (torch.arange(L)[None, :] < ln[:, None]).long()
        return ids * mask, mask
    cid, cm = titles(B * C)
    hid, hm = titles(B * S)
    hl = torch.randint(0, S + 1, (B,), generator=gen)
    hl[0] = 0
    his_mask = (torch.arange(S)[None, :] < torch.clamp(hl, min=1)[:, None]).double().unsqueeze(-1)
    return {"cdd_encoded_index": cid.view(B, C, L), "cdd_attn_mask": cm.view(B, C, L),
            "his_encoded_index": hid.view(B, S, L), "his_attn_mask": hm.view(B, S, L), "his_mask": his_mask,
            "user_id": torch.randint(1, n_users + 1, (B,), generator=gen),
            "label": torch.randint(0, C, (B,), generator=gen)}

"""CPU oracle for the TwoTower hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 restatement of the algorithm the reference
(tyh666/News-Recommendation-MIND) runs on the TwoTower path.  It is the checker
the CUDA kernels are compared against; it is never imported by the product
package (`news_recommendation_mind_b200`).  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Parity pin: the reference ships no tests or golden vectors, so
the pin is made by us: `oracle/make_golden.py` imports the real reference from
`/root/reference` (with the harness shims of SURVEY.md section 8c), runs it on seeded
inputs and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every
function below against those files.

Every function is written functionally (weights are explicit arguments, no
nn.Module state) and cites the reference file:line whose arithmetic it follows.
Gradients are obtained with torch autograd on these functions.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# attention primitives                                     models/Modules/Attention.py
# --------------------------------------------------------------------------------------
def masked_softmax(score: Tensor, mask: Tensor) -> Tensor:
    """Softmax over the last axis restricted to positions where ``mask != 0``.

    Follows XSoftmax.forward (Attention.py:66-74): masked positions get -inf before
    the softmax and are forced to exactly 0 afterwards, so a row with no valid
    position yields all zeros (not NaN).  Its backward (Attention.py:77-80) is the
    ordinary softmax Jacobian applied to that output, p * (g - sum(p*g)), which is
    what autograd produces for the expression below (the where() blocks the NaNs).
    """
    keep = mask.to(torch.bool).expand_as(score)
    neg = torch.finfo(score.dtype).min
    filled = torch.where(keep, score, torch.full_like(score, neg))
    top = filled.max(dim=-1, keepdim=True).values
    ex = torch.where(keep, torch.exp(filled - top), torch.zeros_like(score))
    den = ex.sum(dim=-1, keepdim=True)
    safe = torch.where(den > 0, den, torch.ones_like(den))
    return ex / safe


def attend(query: Tensor, key: Tensor, value: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T / sqrt(d_q)) v with optional mask  (Attention.py:5-30).

    ``d_q`` is the last dimension of *query* (Attention.py:22)."""
    if query.shape[-1] != key.shape[-1]:
        raise AssertionError("query/key width mismatch")
    logit = torch.matmul(query, key.transpose(-2, -1)) / math.sqrt(query.shape[-1])
    prob = torch.softmax(logit, dim=-1) if mask is None else masked_softmax(logit, mask)
    return torch.matmul(prob, value)


def pair_mask(mask: Tensor) -> Tensor:
    """[B, n] 0/1 mask -> [B, 1, n, n] outer product  (Attention.py:33-53)."""
    if mask.dim() != 2:
        raise AssertionError("pair_mask expects a 2-D mask")
    return (mask[:, None, :, None] * mask[:, None, None, :])


def multihead_self_attention(x: Tensor, w_key: Tensor, b_key: Tensor, w_val: Tensor, b_val: Tensor,
                             head_num: int, mask: Optional[Tensor] = None) -> Tensor:
    """MultiheadAttention.forward (Attention.py:115-147).

    Query and key share one projection (Attention.py:125-126); scores are scaled
    by sqrt(key_dim) with key_dim = rows(w_key)/head_num; heads are concatenated,
    there is no output projection."""
    n, length, _ = x.shape
    qk = torch.nn.functional.linear(x, w_key, b_key)
    v = torch.nn.functional.linear(x, w_val, b_val)
    dk = w_key.shape[0] // head_num
    dv = w_val.shape[0] // head_num
    qk = qk.view(n, length, head_num, dk).permute(0, 2, 1, 3)
    v = v.view(n, length, head_num, dv).permute(0, 2, 1, 3)
    logit = torch.matmul(qk, qk.transpose(-1, -2)) / math.sqrt(dk)
    prob = torch.softmax(logit, dim=-1) if mask is None else masked_softmax(logit, mask)
    ctx = torch.matmul(prob, v)                       # [n, hn, len, dv]
    return ctx.permute(0, 2, 1, 3).reshape(n, length, head_num * dv)


# --------------------------------------------------------------------------------------
# embedding                                               models/Embeddings/BERT.py
# --------------------------------------------------------------------------------------
def embed_tokens(table: Tensor, ids: Tensor) -> Tensor:
    """``table[ids]`` (BERT.py:39).  The reference table is BERT's word-embedding
    matrix, an nn.Embedding with padding_idx=0: row 0 is looked up like any other
    row in forward but receives no gradient."""
    return torch.nn.functional.embedding(ids, table, padding_idx=0)


# --------------------------------------------------------------------------------------
# news encoders                                           models/Encoders/CNN.py, MHA.py
# --------------------------------------------------------------------------------------
def cnn_news_encoder(emb: Tensor, attn_mask: Optional[Tensor], conv_w: Tensor, conv_b: Tensor,
                     proj_w: Tensor, proj_b: Tensor, query: Tensor) -> Tuple[Tensor, Tensor]:
    """CNN_Encoder.forward (CNN.py:30-51).

    emb [..., L, E];  conv_w [H, E, 3] (Conv1d layout), zero padding of one
    position on each side of every title (CNN.py:12-17); ReLU; then additive
    attention pooling with key = tanh(c Wq^T + bq), the learned query [1, H] and
    the title mask (CNN.py:44-46).  Returns (c [..., L, H], news [..., H])."""
    lead = emb.shape[:-2]
    L, E = emb.shape[-2:]
    H = conv_w.shape[0]
    x = emb.reshape(-1, L, E)
    xpad = torch.nn.functional.pad(x, (0, 0, 1, 1))               # zero row before and after
    c = conv_b.view(1, 1, H).expand(x.shape[0], L, H).clone()
    for tap in range(3):                                           # out[l] += x[l+tap-1] W[:,:,tap]^T
        c = c + torch.matmul(xpad[:, tap:tap + L, :], conv_w[:, :, tap].t())
    c = torch.relu(c).view(*lead, L, H)
    key = torch.tanh(torch.nn.functional.linear(c, proj_w, proj_b))
    m = None if attn_mask is None else attn_mask.unsqueeze(-2)
    news = attend(query, key, c, m).squeeze(-2)
    return c, news


def mha_news_encoder(emb: Tensor, attn_mask: Tensor, w_key: Tensor, b_key: Tensor, w_val: Tensor,
                     b_val: Tensor, ln_w: Tensor, ln_b: Tensor, query: Tensor, head_num: int,
                     drop_keep: Optional[Tensor] = None, dropout_p: float = 0.0) -> Tuple[Tensor, Tensor]:
    """MHA_Encoder.forward (MHA.py:21-39): self-attention with the pair mask,
    LayerNorm (eps 1e-5), dropout, then query pooling over the encoded tokens
    (no tanh projection here).  ``drop_keep`` is an explicit 0/1 keep mask so the
    dropout draw can be injected for parity; None means eval / p = 0."""
    B = emb.shape[0]
    L, E = emb.shape[-2:]
    flat_mask = attn_mask.reshape(-1, L)
    h = multihead_self_attention(emb.reshape(-1, L, E), w_key, b_key, w_val, b_val, head_num,
                                 pair_mask(flat_mask))
    h = torch.nn.functional.layer_norm(h, (h.shape[-1],), ln_w, ln_b, 1e-5)
    if drop_keep is not None:
        h = h * drop_keep.reshape(h.shape).to(h.dtype) / (1.0 - dropout_p)
    h = h.view(B, -1, L, h.shape[-1])
    news = attend(query, h, h, attn_mask.view(B, -1, 1, L)).squeeze(-2)
    return h, news


# --------------------------------------------------------------------------------------
# user encoders                                  models/Encoders/RNN.py, Pooling.py, MHA.py
# --------------------------------------------------------------------------------------
def _history_lengths(his_mask: Tensor) -> Tensor:
    """his_mask [B, S, 1] (float64 in the reference, MIND.py:332) -> int64 [B]
    (RNN.py:65)."""
    return his_mask.squeeze(-1).sum(dim=-1).to(torch.int64)


def lstm_user_encoder(news: Tensor, his_mask: Optional[Tensor], w_ih: Tensor, w_hh: Tensor,
                      b_ih: Tensor, b_hh: Tensor, descend_history: bool = False,
                      h0: Optional[Tensor] = None) -> Tensor:
    """RNN_User_Encoder.forward with nn.LSTM (RNN.py:50-73).

    The reference packs the padded history by length and returns h_n, i.e. the
    hidden state after step len-1 of every sequence, in original batch order.
    Gate order i, f, g, o; both bias vectors are added.  Returns [B, 1, H]."""
    if descend_history:
        news = news.flip(dims=[1])
    B, S, H = news.shape
    lens = torch.full((B,), S, dtype=torch.int64) if his_mask is None else _history_lengths(his_mask)
    h = torch.zeros(B, H, dtype=news.dtype) if h0 is None else h0
    c = torch.zeros(B, H, dtype=news.dtype)
    for t in range(S):
        gates = news[:, t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, g, o = gates.split(H, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        live = (t < lens).view(B, 1)
        c = torch.where(live, c_new, c)
        h = torch.where(live, h_new, h)
    return h.unsqueeze(1)


def gru_user_encoder(news: Tensor, his_mask: Optional[Tensor], w_ih: Tensor, w_hh: Tensor,
                     b_ih: Tensor, b_hh: Tensor, descend_history: bool = False) -> Tensor:
    """RNN_User_Encoder.forward with nn.GRU (RNN.py:42-43,50-73).  Gate order
    r, z, n;  n = tanh(W_in x + b_in + r * (W_hn h + b_hn));  h' = (1-z) n + z h."""
    if descend_history:
        news = news.flip(dims=[1])
    B, S, H = news.shape
    lens = torch.full((B,), S, dtype=torch.int64) if his_mask is None else _history_lengths(his_mask)
    h = torch.zeros(B, H, dtype=news.dtype)
    for t in range(S):
        gi = news[:, t] @ w_ih.t() + b_ih
        gh = h @ w_hh.t() + b_hh
        ir, iz, inn = gi.split(H, dim=1)
        hr, hz, hn = gh.split(H, dim=1)
        r = torch.sigmoid(ir + hr)
        z = torch.sigmoid(iz + hz)
        n = torch.tanh(inn + r * hn)
        h_new = (1.0 - z) * n + z * h
        h = torch.where((t < lens).view(B, 1), h_new, h)
    return h.unsqueeze(1)


def lstur_user_encoder(news: Tensor, user_index: Tensor, keep_user: Tensor, user_table: Tensor,
                       w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """LSTUR_User_Encoder.forward (RNN.py:88-104), intended semantics.

    h0 = user_table[keep_user * user_index] where keep_user ~ Bernoulli(0.5) per
    sample (drawn in train *and* eval by the reference, RNN.py:100-101; here it
    is an explicit argument so parity tests can inject it); c0 = 0; the LSTM runs
    over the *flipped* history and ignores his_mask (no packing)."""
    h0 = user_table[(keep_user.to(torch.int64) * user_index)]
    return lstm_user_encoder(news.flip(dims=[1]), None, w_ih, w_hh, b_ih, b_hh, h0=h0)


def attention_pooling_user_encoder(news: Tensor, his_mask: Optional[Tensor], query: Tensor) -> Tensor:
    """Attention_Pooling.forward (Pooling.py:12-25): key = value = news vectors,
    mask = his_mask^T [B, 1, S]."""
    m = None if his_mask is None else his_mask.transpose(-1, -2)
    return attend(query, news, news, m)


def average_pooling_user_encoder(news: Tensor) -> Tensor:
    """Average_Pooling.forward (Pooling.py:32-43): plain mean over the history
    axis; the mask is ignored so padded slots (news 0) are averaged in."""
    return news.mean(dim=1, keepdim=True)


def mha_user_encoder(news: Tensor, his_mask: Optional[Tensor], w_key: Tensor, b_key: Tensor,
                     w_val: Tensor, b_val: Tensor, query: Tensor, head_num: int) -> Tensor:
    """MHA_User_Encoder.forward (MHA.py:58-75), intended semantics:
    self-attention over the history with the pair mask of his_mask, then query
    pooling with mask his_mask^T -- the shipped code passes the un-transposed
    [B,S,1] mask and yields a wrong shape.  LayerNorm/dropout exist in the module
    but are never applied."""
    if his_mask is None:
        h = multihead_self_attention(news, w_key, b_key, w_val, b_val, head_num)
        return attend(query, h, h)
    flat = his_mask.squeeze(-1)
    h = multihead_self_attention(news, w_key, b_key, w_val, b_val, head_num, pair_mask(flat))
    return attend(query, h, h, his_mask.transpose(-1, -2))


# --------------------------------------------------------------------------------------
# scoring / loss / optimiser                   models/TwoTowerBaseModel.py, utils/Manager.py
# --------------------------------------------------------------------------------------
def click_score(cdd: Tensor, user: Tensor) -> Tensor:
    """compute_score (TwoTowerBaseModel.py:51-62): <cdd[b,c], user[b]> / sqrt(H)."""
    return torch.matmul(cdd, user.transpose(-2, -1)).squeeze(-1) / math.sqrt(cdd.shape[-1])


def train_logits(score: Tensor) -> Tensor:
    """log_softmax over candidates (TwoTowerBaseModel.py:70-71)."""
    return torch.log_softmax(score, dim=1)


def eval_logits(score: Tensor) -> Tensor:
    """sigmoid (TwoTowerBaseModel.py:72-73, :83)."""
    return torch.sigmoid(score)


def nll_loss(logp: Tensor, label: Tensor) -> Tensor:
    """nn.NLLLoss() with mean reduction (Manager.py:381-382,641)."""
    return -logp.gather(1, label.view(-1, 1)).mean()


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam defaults as built by Manager._get_optim (Manager.py:404-413):
    no weight decay, no amsgrad; bias-corrected; in place.  ``step`` is 1-based."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# whole model                                      models/TwoTower.py, TwoTowerBaseModel.py
# --------------------------------------------------------------------------------------
def encode_news(params: Dict[str, Tensor], ids: Tensor, mask: Tensor, encoder_n: str = "cnn",
                head_num: int = 0, drop_keep: Optional[Tensor] = None, dropout_p: float = 0.0) -> Tensor:
    """TwoTower.encode_news body (TwoTower.py:21-33) for token ids [B, n, L]."""
    emb = embed_tokens(params["embedding.bert_word_embedding.weight"], ids)
    if encoder_n == "cnn":
        return cnn_news_encoder(emb, mask, params["encoderN.cnn.weight"], params["encoderN.cnn.bias"],
                                params["encoderN.wordQueryProject.weight"],
                                params["encoderN.wordQueryProject.bias"],
                                params["encoderN.query_words"])[1]
    if encoder_n == "mha":
        return mha_news_encoder(emb, mask, params["encoderN.mha.keyProject.weight"],
                                params["encoderN.mha.keyProject.bias"],
                                params["encoderN.mha.valueProject.weight"],
                                params["encoderN.mha.valueProject.bias"],
                                params["encoderN.layerNorm.weight"], params["encoderN.layerNorm.bias"],
                                params["encoderN.query_words"], head_num, drop_keep, dropout_p)[1]
    raise ValueError(encoder_n)


def encode_user(params: Dict[str, Tensor], x: Dict[str, Tensor], encoder_n: str = "cnn",
                encoder_u: str = "lstm", head_num: int = 0, descend_history: bool = False,
                keep_user: Optional[Tensor] = None, drop_keep: Optional[Tensor] = None,
                dropout_p: float = 0.0) -> Tensor:
    """TwoTower.encode_user body (TwoTower.py:36-49): the history titles go through
    the same embedding + news encoder, then the user encoder."""
    his = encode_news(params, x["his_encoded_index"], x["his_attn_mask"], encoder_n, head_num,
                      drop_keep, dropout_p)
    hm = x["his_mask"].to(his.dtype)
    if encoder_u == "lstm":
        return lstm_user_encoder(his, hm, params["encoderU.rnn.weight_ih_l0"], params["encoderU.rnn.weight_hh_l0"],
                                 params["encoderU.rnn.bias_ih_l0"], params["encoderU.rnn.bias_hh_l0"],
                                 descend_history)
    if encoder_u == "gru":
        return gru_user_encoder(his, hm, params["encoderU.rnn.weight_ih_l0"], params["encoderU.rnn.weight_hh_l0"],
                                params["encoderU.rnn.bias_ih_l0"], params["encoderU.rnn.bias_hh_l0"],
                                descend_history)
    if encoder_u == "attn":
        return attention_pooling_user_encoder(his, hm, params["encoderU.query_news"])
    if encoder_u == "avg":
        return average_pooling_user_encoder(his)
    if encoder_u == "mha":
        return mha_user_encoder(his, hm, params["encoderU.mha.keyProject.weight"],
                                params["encoderU.mha.keyProject.bias"],
                                params["encoderU.mha.valueProject.weight"],
                                params["encoderU.mha.valueProject.bias"],
                                params["encoderU.query_news"], head_num)
    if encoder_u == "lstur":
        return lstur_user_encoder(his, x["user_id"], keep_user, params["encoderU.userEmbedding.weight"],
                                  params["encoderU.rnn.weight_ih_l0"], params["encoderU.rnn.weight_hh_l0"],
                                  params["encoderU.rnn.bias_ih_l0"], params["encoderU.rnn.bias_hh_l0"])
    raise ValueError(encoder_u)


def forward(params: Dict[str, Tensor], x: Dict[str, Tensor], training: bool, encoder_n: str = "cnn",
            encoder_u: str = "lstm", head_num: int = 0, descend_history: bool = False,
            keep_user: Optional[Tensor] = None, drop_keep_cdd: Optional[Tensor] = None,
            drop_keep_his: Optional[Tensor] = None, dropout_p: float = 0.0) -> Tensor:
    """TwoTowerBaseModel.forward (TwoTowerBaseModel.py:65-75) -> [B, C]."""
    cdd = encode_news(params, x["cdd_encoded_index"], x["cdd_attn_mask"], encoder_n, head_num,
                      drop_keep_cdd, dropout_p)
    user = encode_user(params, x, encoder_n, encoder_u, head_num, descend_history, keep_user,
                       drop_keep_his, dropout_p)
    s = click_score(cdd, user)
    return train_logits(s) if training else eval_logits(s)


def predict_fast(params: Dict[str, Tensor], news_table: Tensor, x: Dict[str, Tensor], **kw) -> Tensor:
    """TwoTowerBaseModel.predict_fast (TwoTowerBaseModel.py:78-84): candidate
    vectors come from the pre-encoded table, the user is encoded from tokens."""
    cdd = news_table[x["cdd_id"]]
    user = encode_user(params, x, **kw)
    return eval_logits(click_score(cdd, user))


def train_step(params: Dict[str, Tensor], state: Dict[str, Tuple[Tensor, Tensor]], x: Dict[str, Tensor],
               step: int, lr: float, bert_lr: float, **kw) -> Tuple[float, Dict[str, Tensor]]:
    """One iteration of Manager._train (Manager.py:636-647): zero grad, forward,
    NLLLoss, backward, Adam with the two learning-rate groups of _get_optim
    (names matching 'bert' use bert_lr, Manager.py:396-413).  Updates ``params`` and
    ``state`` in place and returns (loss, grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    logp = forward(leaves, x, True, **kw)
    loss = nll_loss(logp, x["label"])
    names = [k for k in leaves]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    out = {}
    with torch.no_grad():
        for k, g in zip(names, grads):
            if g is None:
                continue
            out[k] = g
            if k not in state:
                state[k] = (torch.zeros_like(params[k]), torch.zeros_like(params[k]))
            m, v = state[k]
            adam_step(params[k], g, m, v, step, bert_lr if "bert" in k else lr)
    return float(loss.detach()), out


def init_params(V: int, E: int, H: int, encoder_u: str = "lstm", seed: int = 42) -> Dict[str, Tensor]:
    """Fresh CNN + LSTM/GRU TwoTower parameters with the reference's initialisers: xavier-normal conv /
    projection / query (CNN.py:12-24), orthogonal recurrent weights (RNN.py:46-48), PyTorch defaults
    for the biases, token table N(0, 0.02^2) standing in for the downloaded BERT table (BERT.py:16-21)."""
    g = torch.Generator().manual_seed(seed)
    G = 4 if encoder_u == "lstm" else 3

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def xavier(shape, fan_in, fan_out):
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / (fan_in + fan_out))

    def orthogonal(rows, cols):
        q, r = torch.linalg.qr(torch.randn(rows, cols, generator=g))
        return q * torch.sign(torch.diagonal(r))

    return {
        "embedding.bert_word_embedding.weight": torch.randn(V, E, generator=g) * 0.02,
        "encoderN.cnn.weight": xavier((H, E, 3), E * 3, H * 3),
        "encoderN.cnn.bias": uniform((H,), 1.0 / math.sqrt(3 * E)),
        "encoderN.query_words": xavier((1, H), H, 1),
        "encoderN.wordQueryProject.weight": xavier((H, H), H, H),
        "encoderN.wordQueryProject.bias": uniform((H,), 1.0 / math.sqrt(H)),
        "encoderU.rnn.weight_ih_l0": orthogonal(G * H, H),
        "encoderU.rnn.weight_hh_l0": orthogonal(G * H, H),
        "encoderU.rnn.bias_ih_l0": uniform((G * H,), 1.0 / math.sqrt(H)),
        "encoderU.rnn.bias_hh_l0": uniform((G * H,), 1.0 / math.sqrt(H)),
    }

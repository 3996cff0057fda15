"""Drop-in mirrors of the reference's embedding / encoder modules.

Same class names, constructor signature ``Cls(manager)``, parameter names and shapes (hence the
same state-dict keys, so utils/Manager.py save/load keep working) and the same ``forward``
contracts as

    models/Embeddings/BERT.py      BERT_Embedding
    models/Encoders/CNN.py         CNN_Encoder
    models/Encoders/RNN.py         RNN_User_Encoder, LSTUR_User_Encoder
    models/Encoders/Pooling.py     Attention_Pooling, Average_Pooling
    models/Encoders/MHA.py         MHA_Encoder, MHA_User_Encoder

but every forward/backward runs in libmindrec.so (hand-written sm_100a kernels).  torch.nn
containers (nn.Conv1d, nn.LSTM, nn.Linear, nn.Embedding, nn.LayerNorm) are used ONLY to hold
parameters under the reference's names and initialisers -- their own forward is never called.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import MR_BF16, MR_F32, MR_RNN_GRU, MR_RNN_LSTM

DEFAULT_PRECISION = os.environ.get("MINDREC_PRECISION", "bf16")


def precision_of(manager) -> int:
    name = getattr(manager, "precision", None) or DEFAULT_PRECISION
    if name not in ops.PRECISIONS:
        raise ValueError("precision must be one of %s, got %r" % (sorted(ops.PRECISIONS), name))
    return ops.PRECISIONS[name]


# ------------------------------------------------------------------------------------------------
class BERT_Embedding(nn.Module):
    """Token id -> word vector (models/Embeddings/BERT.py:11-41).

    The reference takes ``bert.embeddings.word_embeddings`` of a downloaded BERT (an
    nn.Embedding(30522, 768, padding_idx=0)).  Here the table is given (``weight=``), or loaded
    from a local HF checkpoint when ``manager.bert_checkpoint`` points at one, or -- for the
    synthetic "300d" configurations, which no real BERT has -- drawn N(0, 0.02^2)."""

    def __init__(self, manager, weight: torch.Tensor = None, vocab_size: int = 30522):
        super().__init__()
        self.hidden_dim = manager.bert_dim
        self.precision = precision_of(manager)
        ckpt = getattr(manager, "bert_checkpoint", None)
        if weight is None and ckpt:
            from transformers import AutoModel   # local files only; no network on the box
            weight = AutoModel.from_pretrained(ckpt, local_files_only=True).embeddings.word_embeddings.weight.detach()
        if weight is not None:
            vocab_size, dim = weight.shape
            if dim != self.hidden_dim:
                raise ValueError("embedding width %d != manager.bert_dim %d" % (dim, self.hidden_dim))
        self.bert_word_embedding = nn.Embedding(vocab_size, self.hidden_dim, padding_idx=0)
        with torch.no_grad():
            if weight is not None:
                self.bert_word_embedding.weight.copy_(weight)
            else:
                self.bert_word_embedding.weight.normal_(0.0, 0.02)
        self._shadow = None
        self._shadow_key = None
        # PAD / [CLS] / [SEP] of the BERT vocabulary (utils/Manager.py special-token table; MIND.py:103-127 puts
        # [CLS] first, [SEP] last and pads with 0): kept in shared memory by the fused gather
        hot = [t for t in getattr(manager, "hot_token_ids", (0, 101, 102)) if 0 <= t < vocab_size][:4]
        arr = (ctypes.c_int64 * max(len(hot), 1))(*hot)
        _lib.check(_lib.load().mr_news_cnn_set_hot_tokens(arr, len(hot)), "mr_news_cnn_set_hot_tokens")
        # optional: `hot_replicas` copies of every hot row appended to the bf16 shadow (TMA gather4 path)
        self.hot_ids = hot
        self.hot_reps = int(getattr(manager, "hot_replicas", os.environ.get("MINDREC_HOT_REPS", "0")))
        _lib.check(_lib.load().mr_news_cnn_set_hot_replicas(self.hot_reps), "mr_news_cnn_set_hot_replicas")

    @property
    def weight(self) -> torch.Tensor:
        return self.bert_word_embedding.weight

    def shadow_bf16(self) -> torch.Tensor:
        """bf16 copy of the table, rows padded to a multiple of 64 columns, refreshed whenever
        the fp32 master changed (optimizer steps bump ``_version``)."""
        w = self.weight
        key = (w.data_ptr(), w._version, tuple(w.shape))
        if self._shadow is None or self._shadow_key != key:
            # rows: V (+ optional hot-row replicas), then zero rows up to a multiple of 128 -- the token-grouped backward
            # reads the table in 128-row tiles (mr_news_cnn_bwd_table)
            extra = len(self.hot_ids) * self.hot_reps
            extra += ops.pad_to(w.shape[0] + extra, 128) - (w.shape[0] + extra)
            self._shadow = ops.cast_pad_bf16(w.detach(), ops.pad_to(w.shape[1], 64), extra_rows=extra)
            self._replicate_hot(self._shadow)
            self._shadow_key = key
        return self._shadow

    def _replicate_hot(self, shadow: torch.Tensor) -> None:
        """rows V + h*reps .. V + (h+1)*reps of the shadow <- hot row h (byte copies of bf16 rows)."""
        V = self.weight.shape[0]
        for h, t in enumerate(self.hot_ids):
            if self.hot_reps > 0:
                shadow[V + h * self.hot_reps: V + (h + 1) * self.hot_reps] = shadow[t]

    def invalidate_shadow(self) -> None:
        """Forget the bf16 shadow: the next forward rebuilds it from the fp32 master.  For writers that go around the
        tensor version counter (``weight.data.copy_``, collectives on ``weight.data``)."""
        self._shadow, self._shadow_key = None, None

    def refresh_shadow_inplace(self) -> None:
        """Re-cast the fp32 master into the EXISTING shadow buffer (same address: a captured CUDA graph gathers from it)."""
        if self._shadow is None:
            return
        w = self.weight
        ops.cast_pad_bf16(w.detach(), self._shadow.shape[1], out=self._shadow[: w.shape[0]])
        self._replicate_hot(self._shadow)
        self._shadow_key = (w.data_ptr(), w._version, tuple(w.shape))

    def mark_shadow_fresh(self, shadow: torch.Tensor) -> None:
        """Called by the fused optimiser, which rewrites the shadow inside the Adam kernel."""
        w = self.weight
        self._replicate_hot(shadow)
        self._shadow = shadow
        self._shadow_key = (w.data_ptr(), w._version, tuple(w.shape))

    def forward(self, news_batch: torch.Tensor) -> torch.Tensor:
        if news_batch.dim() == 4:
            # the reference's bag-of-words branch dereferences a non-existent freq_embedding (BERT.py:35-36)
            raise AttributeError("'BERT_Embedding' object has no attribute 'freq_embedding'")
        return ops.EmbeddingGather.apply(news_batch, self.weight, 0)


# ------------------------------------------------------------------------------------------------
class CNN_Encoder(nn.Module):
    """models/Encoders/CNN.py:5-51."""

    def __init__(self, manager):
        super().__init__()
        self.hidden_dim = manager.hidden_dim
        self.embedding_dim = manager.bert_dim
        self.precision = precision_of(manager)
        self.cnn = nn.Conv1d(in_channels=self.embedding_dim, out_channels=self.hidden_dim, kernel_size=3, padding=1)
        nn.init.xavier_normal_(self.cnn.weight)
        self.query_words = nn.Parameter(torch.randn((1, self.hidden_dim), requires_grad=True))
        nn.init.xavier_normal_(self.query_words)
        self.wordQueryProject = nn.Linear(self.hidden_dim, self.hidden_dim)
        nn.init.xavier_normal_(self.wordQueryProject.weight)

    def _run(self, ids, emb, mask, table, table_bf16, want_c):
        return ops.NewsCNN.apply(ids, emb, mask, table, table_bf16, self.cnn.weight, self.cnn.bias,
                                 self.wordQueryProject.weight, self.wordQueryProject.bias, self.query_words,
                                 self.precision, want_c, 0)

    def forward(self, news_embedding, attn_mask=None):
        """[B,*,L,E] (+ mask [B,*,L]) -> (token vectors [B,*,L,H], news vectors [B,*,H])."""
        if news_embedding.shape[-1] != self.embedding_dim:
            raise ValueError("expected embedding width %d, got %d" % (self.embedding_dim, news_embedding.shape[-1]))
        return self._run(None, news_embedding, attn_mask, None, None, True)

    def encode_ids(self, embedding: BERT_Embedding, ids, attn_mask=None):
        """Fused path: token ids straight to news vectors, the [B,*,L,E] tensor is never materialised."""
        shadow = embedding.shadow_bf16() if self.precision == MR_BF16 else None
        return self._run(ids, None, attn_mask, embedding.weight, shadow, False)[1]


# ------------------------------------------------------------------------------------------------
def _lengths_from_mask(his_mask: torch.Tensor) -> torch.Tensor:
    """his_mask [B,S,1] (CPU float64 in the reference batch, MIND.py:332) -> int32 lengths (RNN.py:65)."""
    return his_mask.squeeze(-1).sum(dim=-1).to(torch.int32)


class RNN_User_Encoder(nn.Module):
    """models/Encoders/RNN.py:36-73."""

    def __init__(self, manager):
        super().__init__()
        self.hidden_dim = manager.hidden_dim
        self.descend_history = manager.descend_history
        self.precision = precision_of(manager)
        if manager.encoderU == "gru":
            self.rnn = nn.GRU(self.hidden_dim, self.hidden_dim, batch_first=True)
            self.kind = MR_RNN_GRU
        elif manager.encoderU == "lstm":
            self.rnn = nn.LSTM(self.hidden_dim, self.hidden_dim, batch_first=True)
            self.kind = MR_RNN_LSTM
        else:
            raise ValueError("RNN_User_Encoder needs manager.encoderU in {'lstm','gru'}")
        for name, param in self.rnn.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(param)

    def forward(self, news_repr, **kwargs):
        lens = _lengths_from_mask(kwargs["his_mask"]) if "his_mask" in kwargs and kwargs["his_mask"] is not None else None
        r = self.rnn
        return ops.RNNUser.apply(news_repr, lens, None, r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0,
                                 self.kind, bool(self.descend_history), self.precision)


class LSTUR_User_Encoder(nn.Module):
    """models/Encoders/RNN.py:76-104 with the intended call signature: TwoTower
    passes ``user_id=``; the shipped forward names it ``user_index=`` -- both are accepted."""

    def __init__(self, manager):
        super().__init__()
        self.hidden_dim = manager.hidden_dim
        self.precision = precision_of(manager)
        self.rnn = nn.LSTM(self.hidden_dim, self.hidden_dim, batch_first=True)
        self.userEmbedding = nn.Embedding(manager.get_user_num() + 1, self.hidden_dim)
        with torch.no_grad():
            self.userEmbedding.weight[0].zero_()
        for name, param in self.rnn.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(param)
        self.keep_user = None          # parity tests inject the Bernoulli draw here

    def forward(self, news_repr, his_mask=None, user_index=None, user_id=None, **kwargs):
        idx = user_index if user_index is not None else user_id
        if idx is None:
            raise TypeError("LSTUR_User_Encoder.forward needs user_index / user_id")
        idx = idx.to(news_repr.device)
        B = news_repr.size(0)
        if self.keep_user is not None:
            keep = self.keep_user.to(device=idx.device, dtype=torch.long)
        else:   # RNN.py:100-101: zeros(B).bernoulli_() -> Bernoulli(0.5), train and eval alike
            keep = torch.zeros(B, dtype=torch.long, device=idx.device).bernoulli_()
        h0 = ops.EmbeddingGather.apply(keep * idx, self.userEmbedding.weight, None)
        r = self.rnn
        return ops.RNNUser.apply(news_repr, None, h0, r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0,
                                 MR_RNN_LSTM, True, self.precision)


LSTUR = LSTUR_User_Encoder      # the name twotower.py:44 imports


# ------------------------------------------------------------------------------------------------
class Attention_Pooling(nn.Module):
    """models/Encoders/Pooling.py:5-25."""

    def __init__(self, manager):
        super().__init__()
        self.query_news = nn.Parameter(torch.randn(1, manager.hidden_dim))
        nn.init.xavier_normal_(self.query_news)

    def forward(self, news_reprs, his_mask=None, *args, **kargs):
        return ops.AttnPool.apply(news_reprs, his_mask, self.query_news)


class Average_Pooling(nn.Module):
    """models/Encoders/Pooling.py:28-43 (the mask is ignored, as in the reference)."""

    def __init__(self, manager):
        super().__init__()

    def forward(self, news_reprs, *args, **kargs):
        return ops.AvgPool.apply(news_reprs)


# ------------------------------------------------------------------------------------------------
class MultiheadAttention(nn.Module):
    """models/Modules/Attention.py:83-147: q and k share keyProject, no output projection."""

    def __init__(self, hidden_dim, head_num, key_dim=None, value_dim=None, precision=MR_F32):
        super().__init__()
        self.head_num = head_num
        self.precision = precision
        if not (key_dim and value_dim):
            assert hidden_dim % head_num == 0, "hidden_dim {} must divide head_num {}".format(hidden_dim, head_num)
            head_dim = hidden_dim // head_num
        self.hidden_dim = hidden_dim
        self.key_dim = key_dim if key_dim else head_dim
        self.value_dim = value_dim if value_dim else head_dim
        self.keyProject = nn.Linear(hidden_dim, self.key_dim * head_num)
        self.valueProject = nn.Linear(hidden_dim, self.value_dim * head_num)
        nn.init.xavier_normal_(self.keyProject.weight)
        nn.init.xavier_normal_(self.valueProject.weight)

    def forward(self, hidden_states, token_mask=None):
        """hidden_states [n,len,in]; token_mask [n,len] 0/1 (the pair mask m_i*m_j of get_attn_mask is
        formed inside the kernel) -> [n,len,value_dim*head_num]."""
        if self.precision == MR_BF16:
            return self._forward_tc(hidden_states, None, None, token_mask)
        qk = ops.Linear.apply(hidden_states, self.keyProject.weight, self.keyProject.bias, 0)
        v = ops.Linear.apply(hidden_states, self.valueProject.weight, self.valueProject.bias, 0)
        return ops.MHACore.apply(qk, v, token_mask, self.head_num)

    def _forward_tc(self, hidden_states, ids, embedding, token_mask):
        """MR_BF16: both projections as ONE tensor-core GEMM over the concatenated [keyProject; valueProject] weights (bf16
        operands, fp32 output); with `ids` the token rows are gathered from the bf16 table inside the GEMM."""
        w = torch.cat([self.keyProject.weight, self.valueProject.weight], dim=0)
        b = torch.cat([self.keyProject.bias, self.valueProject.bias], dim=0)
        nk, nv = self.key_dim * self.head_num, self.value_dim * self.head_num
        if ids is None:
            return ops.MHABlock.apply(hidden_states, None, None, None, w, b, token_mask, self.head_num, nk, nv)
        return ops.MHABlock.apply(None, ids, embedding.weight, embedding.shadow_bf16(), w, b, token_mask, self.head_num, nk, nv)


class MHA_Encoder(nn.Module):
    """models/Encoders/MHA.py:5-39."""

    def __init__(self, manager):
        super().__init__()
        self.hidden_dim = manager.hidden_dim
        self.embedding_dim = manager.bert_dim
        self.head_num = manager.head_num
        value_dim, x = divmod(self.hidden_dim, self.head_num)
        assert x == 0, "hidden_dim {} must divide head_num {}".format(self.hidden_dim, self.head_num)
        self.precision = precision_of(manager)
        self.mha = MultiheadAttention(self.embedding_dim, self.head_num, value_dim=value_dim, precision=self.precision)
        self.query_words = nn.Parameter(torch.randn(1, self.hidden_dim))
        self.layerNorm = nn.LayerNorm(self.hidden_dim)
        self.dropOut = nn.Dropout(p=manager.dropout_p)
        self.keep_override = None      # parity tests inject the dropout keep mask here

    def forward(self, news_embedding, attn_mask=None):
        L = news_embedding.size(-2)
        flat = news_embedding.reshape(-1, L, self.embedding_dim)
        m = None if attn_mask is None else attn_mask.reshape(-1, L).to(device=flat.device, dtype=torch.float32)
        h = self.mha(flat, m)
        return self._finish(h, m, news_embedding.shape[:-2], L)

    def encode_ids(self, embedding: BERT_Embedding, ids, attn_mask=None):
        """MR_BF16 fused path: token ids straight into the projection GEMM (rows gathered from the bf16 table shadow), the
        [B,*,L,E] embedding tensor is never materialised.  Returns the news vectors only."""
        L = ids.shape[-1]
        flat_ids = ids.reshape(-1, L)
        m = None if attn_mask is None else attn_mask.reshape(-1, L).to(device=flat_ids.device, dtype=torch.float32)
        h = self.mha._forward_tc(None, flat_ids, embedding, m)
        return self._finish(h, m, ids.shape[:-1], L)[1]

    def _finish(self, h, m, lead, L):
        keep, scale = None, 1.0
        p = self.dropOut.p
        if self.training and p > 0:
            if self.keep_override is not None:
                keep = self.keep_override.to(device=h.device, dtype=torch.uint8).reshape(h.shape)
            else:
                keep = (torch.rand(h.shape, device=h.device) >= p).to(torch.uint8)
            scale = 1.0 / (1.0 - p)
        h = ops.LayerNorm.apply(h, self.layerNorm.weight, self.layerNorm.bias, keep, scale)
        news = ops.AttnPool.apply(h, m, self.query_words).squeeze(1)
        return h.view(*lead, L, self.hidden_dim), news.view(*lead, self.hidden_dim)


class MHA_User_Encoder(nn.Module):
    """models/Encoders/MHA.py:42-75 with the intended mask handling: the pooling
    step uses his_mask^T as Pooling.py:23 does.  layerNorm / dropOut are registered (state-dict
    parity) but, as in the reference forward, never applied."""

    def __init__(self, manager):
        super().__init__()
        self.name = "mha-u"
        self.hidden_dim = manager.hidden_dim
        head_num = manager.head_num
        value_dim, x = divmod(self.hidden_dim, head_num)
        assert x == 0, "hidden_dim {} must divide head_num {}".format(self.hidden_dim, head_num)
        self.precision = precision_of(manager)
        self.mha = MultiheadAttention(self.hidden_dim, manager.head_num, value_dim=value_dim, precision=self.precision)
        self.query_news = nn.Parameter(torch.randn(1, self.hidden_dim))
        self.layerNorm = nn.LayerNorm(self.hidden_dim)
        self.dropOut = nn.Dropout(p=manager.dropout_p)

    def forward(self, news_repr, his_mask=None, **kargs):
        m = None
        if his_mask is not None:
            m = his_mask.squeeze(-1).to(device=news_repr.device, dtype=torch.float32)
        h = self.mha(news_repr, m)
        return ops.AttnPool.apply(h, m, self.query_news)

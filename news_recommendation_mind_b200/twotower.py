"""TwoTower model: drop-in for models/TwoTower.py + models/TwoTowerBaseModel.py.

Same constructor ``TwoTower(manager, embedding, encoderN, encoderU)``, same methods
(forward / encode_news / encode_user / compute_score / predict_fast / init_encoding /
destroy_encoding / init_embedding / destroy_embedding), same attributes (.device, .hidden_dim,
.name), same state-dict keys.  ``x`` is the batch dict produced by utils/MIND.py (CPU tensors);
the model moves what it needs to its device exactly as the reference does.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .modules import BERT_Embedding, CNN_Encoder, MHA_Encoder


class TwoTowerBaseModel(nn.Module):
    """models/TwoTowerBaseModel.py:6-84."""

    def __init__(self, manager):
        super().__init__()
        self.scale = manager.scale
        self.cdd_size = manager.cdd_size
        self.mode = "test" if manager.mode == "test" else "dev"
        self.impr_size = manager.impr_size
        self.batch_size_news = manager.batch_size_news
        self.encoding = False
        self.his_size = manager.his_size
        self.signal_length = manager.signal_length
        self.device = manager.device
        self.hidden_dim = manager.bert_dim

    def init_encoding(self):
        self.encoding = True

    def destroy_encoding(self):
        self.encoding = False

    def init_embedding(self, news_reprs: torch.Tensor = None):
        """Prepare fast inference.  The reference reloads ``news.pt`` from disk on every rank
        (TwoTowerBaseModel.py:34-39); passing the table directly (e.g. the all-gathered shards
        of evaluate.encode_all_news) skips the disk round trip."""
        if news_reprs is None:
            cache_directory = "data/cache/tensors/{}/{}/{}/".format(self.name, self.scale, self.mode)
            news_reprs = torch.load(cache_directory + "news.pt", map_location=torch.device(self.device))
        self.news_reprs = nn.Embedding.from_pretrained(news_reprs.to(self.device).float())

    def destroy_embedding(self):
        self.news_reprs = None
        del self.news_reprs

    # ---- device-resident token table ------------------------------------------
    def attach_news_tokens(self, encoded_news: torch.Tensor, attn_mask: torch.Tensor):
        """Keep the tokenised news set ([N+1, L] ids and attention masks, row 0 = the empty article; utils/MIND.py:103-127)
        in HBM as int32.  Batches may then carry ``cdd_id`` / ``his_id`` only: the ``*_encoded_index`` / ``*_attn_mask``
        tensors the reference's dataset assembles per sample on the host (MIND.py:347-355) are gathered on the device."""
        L = self.signal_length
        self.news_tok_ids = encoded_news[:, :L].to(device=self.device, dtype=torch.int32).contiguous()
        self.news_tok_mask = attn_mask[:, :L].to(device=self.device, dtype=torch.int32).contiguous()

    def detach_news_tokens(self):
        self.news_tok_ids = self.news_tok_mask = None

    def _titles_of(self, x, which):
        """(ids, mask) of the `which` ("cdd" | "his") titles of a batch: from the batch when it carries tokens (the
        reference's contract), else from the resident token table by news id."""
        key = which + "_encoded_index"
        if key in x:
            return x[key].to(self.device, non_blocking=True), x[which + "_attn_mask"].to(self.device, non_blocking=True)
        if getattr(self, "news_tok_ids", None) is None:
            raise KeyError("batch has no %r and no token table is attached (TwoTower.attach_news_tokens)" % key)
        nid = x[which + "_id"]
        ids, mask = ops.gather_titles(self.news_tok_ids, self.news_tok_mask, nid)
        return ids.view(*nid.shape, -1), mask.view(*nid.shape, -1)

    def compute_score(self, cdd_news_repr, user_repr):
        """[B,C,H] x [B,1,H] -> raw scores [B,C] = <cdd, user>/sqrt(H) (TwoTowerBaseModel.py:51-62).
        Kept for API parity (no autograd); forward() uses the fused score+log-softmax / sigmoid kernels."""
        return ops.score_sigmoid(cdd_news_repr, user_repr, apply_sigmoid=False)

    def _logits(self, cdd_repr, user_repr):
        if self.training:
            return ops.ScoreLogSoftmax.apply(cdd_repr, user_repr)
        return ops.score_sigmoid(cdd_repr, user_repr)

    def forward(self, x):
        cdd_repr = self.encode_news(x)
        user_repr, kid = self.encode_user(x)
        return self._logits(cdd_repr, user_repr), kid

    def predict_fast(self, x):
        cdd_id = x["cdd_id"].to(self.device)
        if getattr(self, "history_from_table", False) and "his_id" in x:
            user_repr = self.encode_user_from_table(self.news_reprs.weight, x["his_id"], x)
        else:
            user_repr, _ = self.encode_user(x)
        B, n = cdd_id.shape
        offsets = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=cdd_id.device)
        prob = ops.score_sigmoid_gather(self.news_reprs.weight, cdd_id, offsets, user_repr)
        return prob.view(B, n)


class TwoTower(TwoTowerBaseModel):
    """models/TwoTower.py:3-49."""

    def __init__(self, manager, embedding, encoderN, encoderU):
        super().__init__(manager)
        self.embedding = embedding
        self.encoderN = encoderN
        self.encoderU = encoderU
        self.hidden_dim = manager.hidden_dim
        manager.name = "__".join(["twotower", manager.encoderN, manager.encoderU])
        self.name = manager.name
        self._fused = isinstance(embedding, BERT_Embedding) and isinstance(encoderN, CNN_Encoder)
        # MHA news encoder in bf16 mode: the token gather is fused into the projection GEMM (candidates and history stay two
        # calls, as in the reference, so that the dropout draws keep their order)
        self._fused_mha = isinstance(embedding, BERT_Embedding) and isinstance(encoderN, MHA_Encoder) and \
            getattr(encoderN, "precision", None) == ops.PRECISIONS["bf16"]
        # opt-in (manager.dedup_titles): encode every distinct news of a batch once and gather the vectors back to
        # the (candidate | history) slots -- identical outputs, the backward sums the slot gradients per news with
        # the deterministic segmented reduction.  Needs cdd_id / his_id in the batch (utils/MIND.py:354-355).
        self.dedup_titles = bool(getattr(manager, "dedup_titles", False))
        # opt-in (manager.history_from_table): predict_fast looks the clicked-news vectors up in the news table built by
        # init_embedding instead of re-encoding them from tokens (what models/PLM.py:112-113 does).
        # Same bits: the table rows come from the same batch-invariant encoder, row 0 must be the encoded empty article
        # (evaluate.encode_all_news encodes it, as the reference's own table build does: MIND_news starts at index 0, MIND.py:462-487).
        self.history_from_table = bool(getattr(manager, "history_from_table", False))
        self.news_tok_ids = self.news_tok_mask = None

    # ---- news side ---------------------------------------------------------------------------
    def _encode_titles(self, ids, mask):
        if self._fused or self._fused_mha:
            return self.encoderN.encode_ids(self.embedding, ids, mask)
        return self.encoderN(self.embedding(ids), mask)[1]

    def encode_news(self, x):
        cdd_news, cdd_attn_mask = self._titles_of(x, "cdd")
        return self._encode_titles(cdd_news, cdd_attn_mask)

    # ---- user side ---------------------------------------------------------------------------
    def _encode_user_from(self, his_news_repr, x):
        return self.encoderU(his_news_repr, his_mask=x["his_mask"], user_id=x["user_id"].to(self.device, non_blocking=True))

    def encode_user(self, x):
        his_news, his_attn_mask = self._titles_of(x, "his")
        his_news_repr = self._encode_titles(his_news, his_attn_mask)
        return self._encode_user_from(his_news_repr, x), None

    def encode_user_from_table(self, table, his_id, x):
        """User vectors from already encoded news: ``his_news_repr = table[his_id]`` (models/PLM.py:112-113), then the
        user encoder.  `table` [N+1, H] fp32 with row 0 = the encoded empty article; eval only (no gradient to the table)."""
        his_news_repr = ops.EmbeddingGather.apply(his_id.to(self.device, non_blocking=True), table, None)
        return self._encode_user_from(his_news_repr, x)

    # ---- whole model ---------------------------------------------------------------------------
    def forward(self, x):
        """Same result as encode_news + encode_user + score (TwoTowerBaseModel.py:65-75), but the
        candidate and history titles go through the news encoder as ONE batch so that the big
        kernels see B*(C+S) titles per launch instead of two launches."""
        if not self._fused:
            return super().forward(x)
        if "uniq_id" in x and self.news_tok_ids is not None:
            # id-only batch with a host-made dedup plan (data.dedup_plan): every distinct news of the batch is encoded ONCE (fixed
            # capacity, padded with news 0) and its vector gathered to the (candidate | history) slots -- identical outputs,
            # the backward sums the slot gradients per news with the deterministic segmented reduction
            (B, C), S = x["cdd_id"].shape, x["his_id"].shape[1]
            u_ids, u_mask = ops.gather_titles(self.news_tok_ids, self.news_tok_mask, x["uniq_id"])
            news_u = self._encode_titles(u_ids, u_mask)
            news = ops.EmbeddingGather.apply(x["uniq_inverse"].to(self.device, non_blocking=True), news_u, None)
            self.last_unique_titles = int(x["uniq_id"].numel())
            return self._finish(news, x, B, C, S)
        dedup = self.dedup_titles and "cdd_id" in x and "his_id" in x
        if "cdd_encoded_index" in x:
            cdd = x["cdd_encoded_index"].to(self.device, non_blocking=True)
            his = x["his_encoded_index"].to(self.device, non_blocking=True)
            cm = x["cdd_attn_mask"].to(self.device, non_blocking=True)
            hm = x["his_attn_mask"].to(self.device, non_blocking=True)
            B, C, L = cdd.shape
            S = his.shape[1]
            ids = torch.cat([cdd.reshape(B * C, L), his.reshape(B * S, L)], dim=0)
            mask = torch.cat([cm.reshape(B * C, L), hm.reshape(B * S, L)], dim=0)
        else:
            # id-only batch: the token rows come from the resident table, candidates and history in ONE gather launch
            if self.news_tok_ids is None:
                raise KeyError("batch has no 'cdd_encoded_index' and no token table is attached (TwoTower.attach_news_tokens)")
            (B, C), S = x["cdd_id"].shape, x["his_id"].shape[1]
            ids = mask = None
            if not dedup:
                ids, mask = ops.gather_titles(self.news_tok_ids, self.news_tok_mask, x["cdd_id"], x["his_id"])
        if dedup:
            nid = torch.cat([x["cdd_id"].reshape(-1), x["his_id"].reshape(-1)])           # integer bookkeeping only
            uniq, inverse = torch.unique(nid, return_inverse=True)
            if ids is None:
                u_ids, u_mask = ops.gather_titles(self.news_tok_ids, self.news_tok_mask, uniq)
            else:
                first = torch.full((uniq.numel(),), nid.numel(), dtype=torch.int64, device=nid.device)
                first.scatter_reduce_(0, inverse, torch.arange(nid.numel(), device=nid.device), reduce="amin")
                first = first.to(self.device, non_blocking=True)
                u_ids, u_mask = ids.index_select(0, first), mask.index_select(0, first)
            inverse = inverse.to(self.device, non_blocking=True)
            news_u = self._encode_titles(u_ids, u_mask)
            news = ops.EmbeddingGather.apply(inverse, news_u, None)
            self.last_unique_titles = int(uniq.numel())
        else:
            news = self._encode_titles(ids, mask)
        return self._finish(news, x, B, C, S)

    def _finish(self, news, x, B, C, S):
        cdd_repr = news[: B * C].view(B, C, -1)
        his_repr = news[B * C:].view(B, S, -1)
        user_repr = self._encode_user_from(his_repr, x)
        return self._logits(cdd_repr, user_repr), None

"""Synthetic MIND-shaped data with the exact batch-dict schema utils/MIND.py::__getitem__ emits
(MIND.py:352-363 train, :394-405 dev), for benchmarks and tests (there is no dataset on the box).

Distributions follow SURVEY.md section 8(d): Zipfian word ids in [1000, 30522) between [CLS]=101
and [SEP]=102, title lengths ~ N(0.8L, 0.2L) clipped to [4, L], history lengths ~ LogNormal(3.0,
0.9) clipped to [0, S] (empty history -> his_mask[0] = 1, MIND.py:332-337), Zipfian news ids,
positive always first (label 0, MIND.py:316-327), news 0 = the empty article [101, 102, 0, ...].
"""
from __future__ import annotations

import numpy as np
import torch

VOCAB = 30522
CLS, SEP = 101, 102
NEWS_NUM = {"small_train": 51282, "small_dev": 42416, "large_train": 101527, "large_dev": 72023, "large_test": 120961}
USER_NUM = {"small": 94057, "large": 876956}


def make_news_table(n_news: int, L: int, seed: int = 42, vocab: int = VOCAB):
    """-> (encoded_news [n_news+1, L] int64, attn_mask [n_news+1, L] int64)."""
    rng = np.random.RandomState(seed)
    n = n_news + 1
    length = np.clip(np.rint(rng.normal(0.8 * L, 0.2 * L, size=n)), 4, L).astype(np.int64)
    length[0] = 2
    words = (rng.zipf(1.1, size=(n, L)) - 1) % (vocab - 1000) + 1000
    pos = np.arange(L)[None, :]
    ids = np.where(pos < length[:, None], words, 0)
    ids[:, 0] = CLS
    ids[np.arange(n), length - 1] = SEP
    mask = (pos < length[:, None]).astype(np.int64)
    return torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(mask)


def _zipf_ids(rng, a, size, n_news):
    return (rng.zipf(a, size=size) - 1) % n_news + 1


def dedup_plan(cdd_id, his_id, capacity: int = 0):
    """Host-side bookkeeping of the in-batch unique-news dedup (part of batch assembly, like the reference's collate):
    -> (uniq_id [capacity] int32, inverse [B*(C+S)] int32, n_unique).  ``uniq_id`` lists the distinct news of the batch in
    ascending id order, padded with news 0 up to a FIXED `capacity` (fixed shapes: the step can be replayed as one CUDA
    graph); slot k of the flat (candidates | history) list is title ``uniq_id[inverse[k]]``.  capacity = 0: exact size.
    Returns None when the batch has more distinct news than `capacity` (the caller then sends the plain batch)."""
    nid = np.concatenate([np.asarray(cdd_id).reshape(-1), np.asarray(his_id).reshape(-1)]).astype(np.int64)
    uniq, inverse = np.unique(nid, return_inverse=True)
    n = int(uniq.size)
    cap = capacity or n
    if n > cap:
        return None
    out = np.zeros(cap, dtype=np.int32)
    out[:n] = uniq
    return torch.from_numpy(out), torch.from_numpy(inverse.astype(np.int32)), n


def make_train_batch(news_ids, news_mask, B: int, C: int, S: int, seed: int, n_users: int = USER_NUM["small"],
                     pin: bool = False, id_only: bool = False, dedup_capacity: int = None):
    """One training batch (dict of CPU tensors) in the reference schema.  id_only: the batch of the device-resident
    input pipeline (TwoTower.attach_news_tokens) -- the same sample, but only news ids (int32), the history mask (uint8),
    user ids and labels cross PCIe; the token tensors are gathered on the device."""
    rng = np.random.RandomState(seed)
    n_news = news_ids.shape[0] - 1
    cdd_id = _zipf_ids(rng, 1.05, (B, C), n_news)
    his_len = np.clip(np.rint(rng.lognormal(3.0, 0.9, size=B)), 0, S).astype(np.int64)
    his_id = _zipf_ids(rng, 1.05, (B, S), n_news)
    pos = np.arange(S)[None, :]
    his_id = np.where(pos < his_len[:, None], his_id, 0)
    his_mask = (pos < np.maximum(his_len, 1)[:, None]).astype(np.float64)[:, :, None]
    cdd_t, his_t = torch.from_numpy(cdd_id), torch.from_numpy(his_id)
    x = {
        "user_id": torch.from_numpy(rng.randint(1, n_users + 1, size=B).astype(np.int64)),
        "cdd_id": cdd_t, "his_id": his_t,
        "cdd_encoded_index": news_ids[cdd_t], "his_encoded_index": news_ids[his_t],
        "cdd_attn_mask": news_mask[cdd_t], "his_attn_mask": news_mask[his_t],
        "cdd_mask": torch.ones(B, C, 1, dtype=torch.float64),
        "his_mask": torch.from_numpy(his_mask),
        "label": torch.zeros(B, dtype=torch.int64),
    }
    if id_only:
        x = {"user_id": x["user_id"], "cdd_id": cdd_t.to(torch.int32), "his_id": his_t.to(torch.int32),
             "his_mask": x["his_mask"].to(torch.uint8), "label": x["label"]}
        if dedup_capacity is not None:
            plan = dedup_plan(cdd_id, his_id, dedup_capacity)
            if plan is not None:
                x["uniq_id"], x["uniq_inverse"] = plan[0], plan[1]
    if pin:
        x = {k: v.pin_memory() for k, v in x.items()}
    return x


def make_eval_impressions(news_ids, news_mask, n_impr: int, S: int, seed: int, n_users: int = USER_NUM["small"],
                          with_tokens: bool = True, impr_size: int = 0):
    """Dev impressions in CSR form: candidate counts ~ LogNormal(3.3, 0.8) clipped to [2, 300], labels
    Bernoulli(0.04) with at least one positive and one negative, distinct candidates per impression
    (no score ties).  -> dict with his_* [n_rows, S, ...], cdd_id [n_cand], offsets [n_rows+1], label [n_cand],
    impr_index [n_rows].  Vectorised (the full-scale dev set is 376 k impressions).
    with_tokens=False leaves out his_encoded_index / his_attn_mask (history-from-table evaluation needs his_id only).
    impr_size > 0 cuts impressions with more candidates into chunks of at most impr_size rows that share one
    impr_index and one history, as utils/MIND.py:225-226 does."""
    rng = np.random.RandomState(seed)
    n_news = news_ids.shape[0] - 1
    n_c = np.clip(np.rint(rng.lognormal(3.3, 0.8, size=n_impr)), 2, min(300, n_news)).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(n_c)])
    total = int(offsets[-1])
    # distinct candidates: an arithmetic progression modulo n_news with a stride coprime to n_news, random start
    stride = next(p for p in (7919, 7907, 7901, 104729, 3, 5, 7, 11, 13, 1) if np.gcd(p, n_news) == 1)
    start = rng.randint(0, n_news, size=n_impr).astype(np.int64)
    k = np.arange(total, dtype=np.int64) - np.repeat(offsets[:-1], n_c)
    cdd = (np.repeat(start, n_c) + k * stride) % n_news + 1
    lab = (rng.random_sample(total) < 0.04)
    lab[offsets[:-1] + (rng.random_sample(n_impr) * n_c).astype(np.int64)] = True
    all_pos = np.add.reduceat(lab.astype(np.int64), offsets[:-1]) == n_c
    lab[offsets[:-1][all_pos]] = False
    lab = lab.astype(np.float32)
    his_len = np.clip(np.rint(rng.lognormal(3.0, 0.9, size=n_impr)), 0, S).astype(np.int64)
    his_id = _zipf_ids(rng, 1.05, (n_impr, S), n_news)
    pos = np.arange(S)[None, :]
    his_id = np.where(pos < his_len[:, None], his_id, 0)
    his_mask = (pos < np.maximum(his_len, 1)[:, None]).astype(np.float64)[:, :, None]
    user_id = rng.randint(1, n_users + 1, size=n_impr).astype(np.int64)
    impr_index = np.arange(n_impr, dtype=np.int64)
    if impr_size and impr_size > 0:
        n_chunks = (n_c + impr_size - 1) // impr_size
        row_of = np.repeat(np.arange(n_impr), n_chunks)                      # source impression of every row
        chunk_no = np.arange(row_of.size) - np.repeat(np.concatenate([[0], np.cumsum(n_chunks)])[:-1], n_chunks)
        row_cnt = np.minimum(impr_size, n_c[row_of] - chunk_no * impr_size)
        offsets = np.concatenate([[0], np.cumsum(row_cnt)])                  # candidate order is unchanged
        his_id, his_mask, user_id, impr_index = his_id[row_of], his_mask[row_of], user_id[row_of], impr_index[row_of]
    his_t = torch.from_numpy(his_id)
    out = {
        "impr_index": torch.from_numpy(impr_index), "user_id": torch.from_numpy(user_id),
        "cdd_id": torch.from_numpy(cdd), "offsets": torch.from_numpy(offsets.astype(np.int64)), "label": torch.from_numpy(lab),
        "his_id": his_t, "his_mask": torch.from_numpy(his_mask),
    }
    if with_tokens:
        out["his_encoded_index"] = news_ids[his_t]
        out["his_attn_mask"] = news_mask[his_t]
    return out

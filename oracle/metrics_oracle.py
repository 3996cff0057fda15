"""CPU oracle for the ranking metrics and eval bookkeeping -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the functions the reference evaluates with
(utils/Manager.py:1205-1344, utils/utils.py:60-80,267-283, Manager.py:842-850).
Pinned against the reference's own `cal_metric` by `oracle/make_golden.py`
(-> tests/golden/metrics.npz) and `tests/test_oracle_golden.py`.

Tie rules.  prediction.txt ranks are ``scipy.stats.rankdata(1 - p, "ordinal")``
(Manager.py:846): descending score, ties by ASCENDING candidate position -- exact.
MRR / nDCG order candidates with ``np.argsort(score)[::-1]`` (Manager.py:1216,1269):
a reversed ascending sort, i.e. ties by DESCENDING position wherever numpy's sort
is stable (numpy 1.x, the reference's era: insertion sort for n <= 16).  For longer
arrays, and in numpy 2.x with AVX-512 for n >= 4 (measured in this container), the
default sort is not stable and the reference's tie order is an implementation
detail no rule reproduces.  This oracle and the CUDA ranking kernel define the
MRR / nDCG order as *stable ascending sort, reversed*; the golden vectors from the
reference's cal_metric are tie-free, so all three agree to the last bit there, and
AUC is tie-order independent.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np


def descending_order(score: np.ndarray) -> np.ndarray:
    """Candidate positions sorted by descending score, ties by position."""
    score = np.asarray(score)
    return np.argsort(-score, kind="stable")


def ordinal_rank(score: np.ndarray) -> np.ndarray:
    """1-based rank of every candidate, best score = 1, ties by position:
    ``scipy.stats.rankdata(1 - score, method="ordinal")`` (Manager.py:846)."""
    order = descending_order(score)
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(1, len(order) + 1)
    return rank


def auc(label: np.ndarray, score: np.ndarray) -> float:
    """Area under the ROC curve of one impression = P(score_pos > score_neg) +
    0.5 P(equal), the Mann-Whitney form of sklearn.roc_auc_score
    (Manager.py:1280-1287)."""
    label = np.asarray(label)
    score = np.asarray(score, dtype=np.float64)
    pos = score[label == 1]
    neg = score[label != 1]
    if len(pos) == 0 or len(neg) == 0:
        raise ValueError("AUC needs at least one positive and one negative")
    gt = (pos[:, None] > neg[None, :]).sum()
    eq = (pos[:, None] == neg[None, :]).sum()
    return float((gt + 0.5 * eq) / (len(pos) * len(neg)))


def metric_order(score: np.ndarray) -> np.ndarray:
    """``np.argsort(score)[::-1]`` with a stable sort: descending score, ties by DESCENDING position
    (Manager.py:1216,1269)."""
    return np.argsort(np.asarray(score), kind="stable")[::-1]


def mrr(label: np.ndarray, score: np.ndarray) -> float:
    """mrr_score (Manager.py:1205-1221): sum of label/rank over sum of labels."""
    y = np.asarray(label, dtype=np.float64)[metric_order(score)]
    return float(np.sum(y / (np.arange(len(y)) + 1.0)) / np.sum(y))


def dcg(label: np.ndarray, score: np.ndarray, k: int) -> float:
    """dcg_score (Manager.py:1257-1273): gains 2^y - 1, discounts log2(pos + 2),
    first min(k, n) positions."""
    k = min(len(label), k)
    y = np.asarray(label, dtype=np.float64)[metric_order(score)[:k]]
    return float(np.sum((2.0 ** y - 1.0) / np.log2(np.arange(len(y)) + 2.0)))


def ndcg(label: np.ndarray, score: np.ndarray, k: int) -> float:
    """ndcg_score (Manager.py:1224-1237): DCG over the ideal DCG (labels ranked by
    themselves)."""
    return dcg(label, score, k) / dcg(label, label, k)


def ranking_metrics(labels: Sequence[Sequence[float]], preds: Sequence[Sequence[float]],
                    ndcg_at: Tuple[int, ...] = (5, 10)) -> Dict[str, float]:
    """cal_metric for the default list ``auc,mean_mrr,ndcg@5,ndcg@10``
    (Manager.py:106,1276-1344): per-impression values, mean, rounded to 4 dp."""
    res = {
        "auc": round(float(np.mean([auc(l, p) for l, p in zip(labels, preds)])), 4),
        "mean_mrr": round(float(np.mean([mrr(l, p) for l, p in zip(labels, preds)])), 4),
    }
    for k in ndcg_at:
        res["ndcg@%d" % k] = round(float(np.mean([ndcg(l, p, k) for l, p in zip(labels, preds)])), 4)
    return res


def group_by_impression(impr_indexes: Iterable[int], *columns: Iterable[Sequence]) -> List[List[list]]:
    """_group_lists (utils.py:60-80): chunks that share an impression index are
    concatenated in arrival order; groups come out in first-appearance order."""
    buckets = [OrderedDict() for _ in columns]
    for row in zip(impr_indexes, *columns):
        key = row[0]
        for b, chunk in zip(buckets, row[1:]):
            b.setdefault(key, []).extend(chunk)
    return [list(b.values()) for b in buckets]


def partition_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Partition_Sampler (utils.py:267-283): contiguous shards of n // ws items,
    the remainder goes to the last rank."""
    per, extra = divmod(n_items, world_size)
    start = per * rank
    end = start + per + (extra if rank + 1 == world_size else 0)
    return start, end

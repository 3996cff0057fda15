"""Fast evaluation of the hot path: what utils/Manager.py::_eval_fast does (Manager.py:473-541),
re-shaped for 8 GPUs over NVLink.

reference                                   here
---------                                   ----
rank 0 encodes every news, torch.save to    every rank encodes a contiguous shard of the news set
disk, barrier, all ranks torch.load         and the [ceil((N+1)/ws), H] shards are ALL-GATHERED
(Manager.py:490-510)                        over NCCL into the replicated [N+1, H] table
one impression per Python iteration,        impressions of this rank's contiguous partition
.tolist() each (Manager.py:514-517)         (Partition_Sampler, utils.py:267-283) are scored in
                                            batches: encode_user -> CSR gather+dot+sigmoid
all_gather_object of Python lists to        per-impression AUC/MRR/nDCG@5/10 on the GPU, then an
rank 0 + cal_metric (Manager.py:525,577)    all-reduce of (sum, count) in fp64
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_rows: int, world: int, rank: int):
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per), per


def partition_bounds(n_items: int, world: int, rank: int):
    """Partition_Sampler (utils.py:267-283): n // ws each, remainder to the last rank."""
    per, extra = divmod(n_items, world)
    start = per * rank
    return start, start + per + (extra if rank + 1 == world else 0)


@torch.no_grad()
def encode_all_news(model, news_ids: torch.Tensor, news_mask: torch.Tensor, batch: int = 32768) -> torch.Tensor:
    """[N+1, L] token table -> replicated [N+1, H] news-vector table (fp32).  Shards over ranks.

    The reference walks the news set in DataLoader batches of `batch_size_news` = 500 titles (Manager.py:498-499); the
    encoder here is batch invariant (same vector whatever the batch a title sits in), so this rank's shard goes to the
    device in ONE copy (asynchronous when the host table is pinned) and through the encoder in chunks of `batch`
    titles -- three kernel launches per chunk instead of per 500 titles."""
    rank, world = _world()
    core = model.module if hasattr(model, "module") else model
    dev = core.device
    n_rows = news_ids.shape[0]
    lo, hi, per = shard_bounds(n_rows, world, rank)
    shard = torch.zeros(per, core.hidden_dim, dtype=torch.float32, device=dev)
    was_training = core.training
    core.eval()
    core.init_encoding()
    ids_d = news_ids[lo:hi].to(dev, non_blocking=True)
    mask_d = news_mask[lo:hi].to(dev, non_blocking=True)
    for a in range(0, hi - lo, batch):
        b = min(hi - lo, a + batch)
        x = {"cdd_encoded_index": ids_d[a:b].unsqueeze(1), "cdd_attn_mask": mask_d[a:b].unsqueeze(1)}
        shard[a:b] = core.encode_news(x).squeeze(-2)
    core.destroy_encoding()
    core.train(was_training)
    return gather_news_shards(shard, n_rows)


def gather_news_shards(shard: torch.Tensor, n_rows: int) -> torch.Tensor:
    """All-gather of the per-rank [ceil(n_rows/ws), H] shards into the replicated [n_rows, H] table -- replaces
    the reference's torch.save / barrier / torch.load round trip through the file system (Manager.py:503-510)."""
    _, world = _world()
    if world == 1:
        return shard[:n_rows]
    full = torch.empty(world * shard.shape[0], shard.shape[1], dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(full, shard.contiguous())
    return full[:n_rows]


def reduce_metric_sums(per_impression: torch.Tensor) -> torch.Tensor:
    """[n_local, 4] fp64 per-impression (auc, mrr, ndcg@5, ndcg@10) -> global means [4] (all-reduce of sums and
    the impression count; replaces all_gather_object + rank-0 cal_metric, Manager.py:525,577)."""
    _, world = _world()
    acc = torch.cat([per_impression.sum(0), torch.tensor([float(per_impression.shape[0])], dtype=torch.float64,
                                                          device=per_impression.device)])
    if world > 1:
        dist.all_reduce(acc)
    return acc[:4] / acc[4]


def group_rows(impr_index):
    """_group_lists (utils/utils.py:60-80) as index bookkeeping.  The reference cuts an impression with more than `impr_size`
    candidates into chunks that share one ``impr_index`` (utils/MIND.py:225-226), scores the chunks as separate rows and
    concatenates rows with equal index -- in arrival order, groups in first-appearance order -- before the metrics
    (Manager.py:525-536).  Returns (row_order, group_row_offsets): ``row_order`` is None when every group is already
    contiguous (the usual case: chunks are adjacent), else the stable permutation that makes them so;
    ``group_row_offsets`` [n_groups + 1] indexes rows (after the permutation)."""
    idx = torch.as_tensor(impr_index).reshape(-1).cpu()
    n = idx.numel()
    if n == 0:
        return None, torch.zeros(1, dtype=torch.int64)
    uniq, inverse = torch.unique(idx, return_inverse=True)
    first = torch.full((uniq.numel(),), n, dtype=torch.int64)
    first.scatter_reduce_(0, inverse, torch.arange(n), reduce="amin")
    gid = torch.argsort(torch.argsort(first))[inverse]          # group number in first-appearance order, per row
    order = None
    if bool((gid[1:] < gid[:-1]).any()):
        order = torch.sort(gid, stable=True).indices
        gid = gid[order]
    starts = torch.nonzero(torch.cat([torch.ones(1, dtype=torch.bool), gid[1:] != gid[:-1]])).reshape(-1)
    return order, torch.cat([starts, torch.tensor([n])])


def reorder_rows(impr: dict, order: torch.Tensor) -> dict:
    """Apply a row permutation to a CSR impression dict (per-row tensors + the candidate arrays)."""
    offs = impr["offsets"]
    n_rows = offs.numel() - 1
    cnt = (offs[1:] - offs[:-1])[order]
    new_off = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(cnt, 0)])
    cand = torch.repeat_interleave(offs[:-1][order] - new_off[:-1], cnt) + torch.arange(int(new_off[-1]))
    out = {}
    for k, v in impr.items():
        if not torch.is_tensor(v):
            out[k] = v
        elif k == "offsets":
            out[k] = new_off
        elif k in ("cdd_id", "label"):
            out[k] = v[cand]
        elif v.shape[0] == n_rows:
            out[k] = v[order]
        else:
            out[k] = v
    return out


def group_partition_bounds(group_row_offsets: torch.Tensor, world: int, rank: int):
    """Rows [r0, r1) and groups [g0, g1) of this rank: Partition_Sampler's row split (utils.py:267-283) with every boundary moved
    up to the next group start, so that no impression is ranked from a partial candidate list (the reference gets the same effect
    by gathering every rank's rows to rank 0 before _group_lists, Manager.py:525-536)."""
    n_rows = int(group_row_offsets[-1])
    n_groups = group_row_offsets.numel() - 1

    def cut(r):
        if r >= world:
            return n_groups
        return int(torch.searchsorted(group_row_offsets, torch.tensor(partition_bounds(n_rows, world, r)[0])))
    g0, g1 = cut(rank), cut(rank + 1)
    return int(group_row_offsets[g0]), int(group_row_offsets[g1]), g0, g1


@torch.no_grad()
def score_impressions(model, table: torch.Tensor, impr: dict, batch: int = 4096, rows=None, history: str = "auto"):
    """Scores CSR impression rows [rows[0], rows[1]) (default: this rank's Partition_Sampler share).
    Returns (prob [n_cand_local], label, offsets_local).

    history = "table": the clicked-news vectors are looked up in `table` by ``his_id`` (what models/PLM.py:112-113 does; for
    TwoTower the reference re-encodes 50 titles per impression from tokens, TwoTowerBaseModel.py:78-84 -> TwoTower.py:36-49).
    Identical bits: the table rows were produced by the same batch-invariant encoder, row 0 = the empty article that pads
    histories.  history = "tokens": the reference's flow.  "auto": table when the impressions carry ``his_id``."""
    rank, world = _world()
    core = model.module if hasattr(model, "module") else model
    dev = core.device
    n_impr = impr["offsets"].numel() - 1
    i0, i1 = rows if rows is not None else partition_bounds(n_impr, world, rank)
    if history == "auto":
        history = "table" if "his_id" in impr else "tokens"
    offs = impr["offsets"]
    c0, c1 = int(offs[i0]), int(offs[i1])
    cdd = impr["cdd_id"][c0:c1].to(dev, non_blocking=True)
    local_off = (offs[i0:i1 + 1] - c0).to(dev)
    prob = torch.empty(c1 - c0, dtype=torch.float32, device=dev)
    was_training = core.training
    core.eval()
    if history == "table":
        his_id = impr["his_id"][i0:i1].to(dev, non_blocking=True)
        his_mask = impr["his_mask"][i0:i1].to(dev, non_blocking=True)
        user_id = impr["user_id"][i0:i1].to(dev, non_blocking=True)
    offs_l = offs[i0:i1 + 1].tolist()
    for a in range(i0, i1, batch):
        b = min(i1, a + batch)
        if history == "table":
            x = {"his_mask": his_mask[a - i0:b - i0], "user_id": user_id[a - i0:b - i0]}
            user = core.encode_user_from_table(table, his_id[a - i0:b - i0], x)
        else:
            x = {"his_encoded_index": impr["his_encoded_index"][a:b], "his_attn_mask": impr["his_attn_mask"][a:b],
                 "his_mask": impr["his_mask"][a:b], "user_id": impr["user_id"][a:b]}
            user = core.encode_user(x)[0]
        ca, cb = offs_l[a - i0] - c0, offs_l[b - i0] - c0
        sub_off = local_off[a - i0:b - i0 + 1] - ca
        prob[ca:cb] = ops.score_sigmoid_gather(table, cdd[ca:cb], sub_off, user)
    core.train(was_training)
    return prob, impr["label"][c0:c1].to(dev, non_blocking=True), local_off


@torch.no_grad()
def evaluate(model, news_ids, news_mask, impr, metrics=("auc", "mean_mrr", "ndcg@5", "ndcg@10"), history: str = "auto",
             table: torch.Tensor = None, ndigits: int = 4):
    """-> dict of the reference's default metrics rounded to 4 dp (Manager.py:106,1276-1344; ndigits=None: unrounded).
    Rows that share an ``impr_index`` are merged before ranking (group_rows)."""
    rank, world = _world()
    if table is None:
        table = encode_all_news(model, news_ids, news_mask)
    n_rows = impr["offsets"].numel() - 1
    if "impr_index" in impr:
        order, g_off = group_rows(impr["impr_index"])
        if order is not None:
            impr = reorder_rows(impr, order)
    else:
        g_off = torch.arange(n_rows + 1)
    r0, r1, g0, g1 = group_partition_bounds(g_off, world, rank)
    prob, label, off = score_impressions(model, table, impr, rows=(r0, r1), history=history)
    merged = off[(g_off[g0:g1 + 1] - r0).to(off.device)]        # candidate offsets at the group starts
    if g1 > g0:
        m, _ = ops.rank_metrics(prob, label, merged)
    else:
        m = torch.zeros(0, 4, dtype=torch.float64, device=prob.device)
    mean = reduce_metric_sums(m).tolist()
    names = ["auc", "mean_mrr", "ndcg@5", "ndcg@10"]
    return {k: (round(v, ndigits) if ndigits is not None else v) for k, v in zip(names, mean) if k in metrics}


def write_predictions(path: str, ranks, offsets, first_index: int = 1) -> int:
    """`prediction.txt` of Manager.test (utils/Manager.py:842-850): one line per impression,
    ``<index> [r1,r2,...]`` with the ordinal ranks of the candidates (1 = highest probability, ties in candidate
    order = scipy.stats.rankdata(1 - p, method="ordinal")), impressions numbered from 1.  `ranks` is the int32 output
    of ops.rank_metrics(..., want_rank=True) (CSR over `offsets`); both may live on the device.  Returns the number
    of lines written."""
    r = ranks.detach().cpu().tolist() if torch.is_tensor(ranks) else list(ranks)
    off = offsets.detach().cpu().tolist() if torch.is_tensor(offsets) else list(offsets)
    with open(path, "w") as f:
        for i in range(len(off) - 1):
            f.write(str(first_index + i) + " [" + ",".join(str(int(v)) for v in r[off[i]:off[i + 1]]) + "]" + "\n")
    return len(off) - 1

"""torch.autograd.Function wrappers over the C ABI (libmindrec.so).

PyTorch is used for memory, streams and autograd bookkeeping only: every arithmetic step on
the hot path is a hand-written sm_100a kernel behind include/mindrec.h.  Nothing here has a
CPU or eager-PyTorch fallback; a missing library or a non-B200 device raises.
"""
from __future__ import annotations

from ctypes import byref, c_float, c_int, c_void_p

import os

import torch

from . import _lib
from ._lib import CnnShape, RnnShape, MR_BF16, MR_F32, check, index_flag, ptr, stream_ptr, workspace

PRECISIONS = {"fp32": MR_F32, "f32": MR_F32, "bf16": MR_BF16}


# MINDREC_GROUPED=0 switches the bf16 ids path back to the per-token d_emb + segmented reduction (A/B, tests)
GROUPED_TABLE_GRAD = os.environ.get("MINDREC_GROUPED", "1") != "0"


def ctypes_stream(stream: torch.cuda.Stream):
    return c_void_p(stream.cuda_stream)


# set by trainer.GradSync: called as hook(d_table, event) inside the encoder backward, `event` (a torch.cuda.Event) having been
# recorded the moment d_table was complete -- the data-parallel all-reduce of the table gradient starts there.  A hook that
# returns True owns the gradient: backward then returns None for the table (autograd's AccumulateGrad would otherwise CLONE a
# tensor somebody else still references, i.e. copy it before the in-place all-reduce has run)
TABLE_GRAD_HOOK = None

_SIDE_STREAMS = {}


def side_stream(device) -> torch.cuda.Stream:
    """One auxiliary stream per device for work that only depends on the integer inputs (token-grouping plans)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return _SIDE_STREAMS[key]


def pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _idx(t: torch.Tensor) -> torch.Tensor:
    if t.dtype not in (torch.int32, torch.int64):
        t = t.long()
    return t.contiguous()


# --------------------------------------------------------------------------------------------
# embedding gather                                                  models/Embeddings/BERT.py:39
# --------------------------------------------------------------------------------------------
def embed_grad(ids: torch.Tensor, d_emb: torch.Tensor, V: int, E: int, padding_idx: int = 0) -> torch.Tensor:
    """Dense [V,E] fp32 gradient of the table from per-token gradients (atomic-free segmented
    reduction, mr_embed_grad_segreduce)."""
    lib = _lib.load()
    T = ids.numel()
    d_table = torch.empty(V, E, dtype=torch.float32, device=d_emb.device)
    ws = workspace(lib.mr_embed_grad_workspace_bytes(T, E, V), d_emb.device)
    dt = MR_BF16 if d_emb.dtype == torch.bfloat16 else MR_F32
    ld = d_emb.shape[-1]
    check(lib.mr_embed_grad_segreduce(ptr(ids), index_flag(ids), ptr(d_emb), dt, ld, ptr(d_table), T, E, V, padding_idx,
                                      ptr(ws), ws.numel(), stream_ptr(d_emb.device)), "mr_embed_grad_segreduce")
    return d_table


def gather_titles(tok_ids: torch.Tensor, tok_mask: torch.Tensor, nid_a: torch.Tensor, nid_b: torch.Tensor = None):
    """Token rows of the news `nid_a` (then `nid_b`) from the device-resident int32 token table -> (ids, mask) int32
    [n_a + n_b, L].  Integer work, bit-exact (utils/MIND.py:347-355 done on the device)."""
    lib = _lib.load()
    if tok_ids.dtype != torch.int32 or tok_mask.dtype != torch.int32 or tok_ids.shape != tok_mask.shape:
        raise ValueError("gather_titles: token table must be two int32 [n_rows, L] tensors")
    dev = tok_ids.device
    a = _idx(nid_a.to(dev, non_blocking=True)).view(-1)
    b = None if nid_b is None else _idx(nid_b.to(dev, non_blocking=True)).view(-1)
    if b is not None and b.dtype != a.dtype:
        b = b.to(a.dtype)
    n_rows, L = tok_ids.shape
    n = a.numel() + (0 if b is None else b.numel())
    out_ids = torch.empty(n, L, dtype=torch.int32, device=dev)
    out_mask = torch.empty(n, L, dtype=torch.int32, device=dev)
    check(lib.mr_gather_titles(ptr(tok_ids), ptr(tok_mask), n_rows, L, ptr(a), a.numel(), ptr(b), 0 if b is None else b.numel(),
                               index_flag(a), ptr(out_ids), ptr(out_mask), stream_ptr(dev)), "mr_gather_titles")
    return out_ids, out_mask


class EmbeddingGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table, padding_idx):
        lib = _lib.load()
        ids_c = _idx(ids)
        table_c = _f32c(table)
        V, E = table_c.shape
        out = torch.empty(*ids.shape, E, dtype=torch.float32, device=table.device)
        check(lib.mr_embed_gather_f32(ptr(ids_c), index_flag(ids_c), ptr(table_c), ptr(out), ids_c.numel(), E, V,
                                      stream_ptr(table.device)), "mr_embed_gather_f32")
        ctx.save_for_backward(ids_c)
        ctx.shape = (V, E)
        ctx.padding_idx = -1 if padding_idx is None else int(padding_idx)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (ids_c,) = ctx.saved_tensors
        V, E = ctx.shape
        d_table = embed_grad(ids_c, _f32c(d_out), V, E, ctx.padding_idx)
        return None, d_table, None


# --------------------------------------------------------------------------------------------
# CNN news encoder                                                  models/Encoders/CNN.py:30-51
# --------------------------------------------------------------------------------------------
class NewsCNN(torch.autograd.Function):
    """(ids | emb, mask, table, conv_w, conv_b, proj_w, proj_b, query) -> (c, news).

    ids path: `table` is the fp32 master table [V,E]; in bf16 mode `table_bf16` is its padded
    bf16 shadow [V,Epad] that the kernel gathers from.  The gradient of the table is produced
    by the segmented reduction inside backward (padding row skipped)."""

    @staticmethod
    def forward(ctx, ids, emb, mask, table, table_bf16, conv_w, conv_b, proj_w, proj_b, query, precision, want_c,
                padding_idx):
        lib = _lib.load()
        dev = conv_w.device
        H, E, _ = conv_w.shape
        if ids is not None:
            ids_c = _idx(ids)
            L = ids_c.shape[-1]
            N = ids_c.numel() // L
            emb_c = None
            V = table.shape[0]
            lead = ids.shape[:-1]
        else:
            ids_c = None
            emb_c = _f32c(emb)
            L = emb_c.shape[-2]
            N = emb_c.numel() // (L * E)
            V = 0
            lead = emb.shape[:-2]
        mask_c = None if mask is None else _idx(mask)
        shape = CnnShape(N, L, E, H, V, precision)
        if precision == MR_BF16:
            Hp = pad_to(H, 16)
            # rows [0, N*L): c; behind them 32 bytes per token: the sign mask of c the conv epilogue writes for the backward
            c_save = torch.empty(N * L + (N * L * 32 + 2 * Hp - 1) // (2 * Hp), Hp, dtype=torch.bfloat16, device=dev)
            key_save = torch.empty(N * L, Hp, dtype=torch.bfloat16, device=dev)
            tab = table_bf16 if ids is not None else None
        else:
            c_save = torch.empty(N * L, H, dtype=torch.float32, device=dev)
            key_save = torch.empty(N * L, H, dtype=torch.float32, device=dev)
            tab = _f32c(table) if ids is not None else None
        prob = torch.empty(N, L, dtype=torch.float32, device=dev)
        news = torch.empty(N, H, dtype=torch.float32, device=dev)
        cw, cb, pw, pb, q = _f32c(conv_w), _f32c(conv_b), _f32c(proj_w), _f32c(proj_b), _f32c(query)
        ws = workspace(lib.mr_news_cnn_workspace_bytes(byref(shape), 0), dev)
        check(lib.mr_news_cnn_fwd(byref(shape), ptr(ids_c), index_flag(ids_c) if ids_c is not None else 0, ptr(emb_c),
                                  ptr(mask_c), index_flag(mask_c) if mask_c is not None else 0, ptr(tab), ptr(cw),
                                  ptr(cb), ptr(pw), ptr(pb), ptr(q), ptr(c_save), ptr(key_save), ptr(prob), ptr(news),
                                  ptr(ws), ws.numel(), stream_ptr(dev)), "mr_news_cnn_fwd")
        # token-grouping plan of the backward (a function of the ids only): queued on a side stream BEHIND this encoder's
        # forward kernels (they own every SM's shared memory, nothing can run beside them), so that the sort and the
        # segment bookkeeping overlap the latency-bound user encoder (128 of 148 SMs) instead of sitting on the
        # backward's critical path
        ctx.group_plan, ctx.group_event = None, None
        if (precision == MR_BF16 and ids is not None and table is not None and ctx.needs_input_grad[3]
                and GROUPED_TABLE_GRAD and E % 4 == 0 and tab.shape[0] >= pad_to(V, 32)):
            nb = lib.mr_token_group_plan_bytes(N * L, V)
            plan = torch.empty(nb, dtype=torch.uint8, device=dev)
            cur, side = torch.cuda.current_stream(dev), side_stream(dev)
            side.wait_stream(cur)
            check(lib.mr_token_group_plan(ptr(ids_c), index_flag(ids_c), N * L, V, ptr(plan), nb, ctypes_stream(side)), "mr_token_group_plan")
            ctx.group_event = torch.cuda.Event()
            ctx.group_event.record(side)
            ids_c.record_stream(side)
            plan.record_stream(side)         # allocated on the compute stream, written on the side stream: if the backward never
            ctx.group_plan = plan            # runs (and so never waits for group_event) the block must not be recycled early
        ctx.save_for_backward(ids_c, emb_c, tab, cw, pw, q, c_save, key_save, prob)
        ctx.shape = shape
        ctx.table_shape = None if table is None else tuple(table.shape)
        ctx.padding_idx = -1 if padding_idx is None else int(padding_idx)
        ctx.need_table_grad = table is not None and table.requires_grad
        ctx.need_emb_grad = emb is not None and emb.requires_grad
        ctx.set_materialize_grads(False)
        news_out = news.view(*lead, H)
        if want_c:
            c_out = c_save if precision == MR_F32 else c_save[:N * L, :H].float()
            c_out = c_out.view(*lead, L, H)
        else:
            c_out = None
        return c_out, news_out

    @staticmethod
    def backward(ctx, d_c, d_news):
        lib = _lib.load()
        ids_c, emb_c, tab, cw, pw, q, c_save, key_save, prob = ctx.saved_tensors
        s = ctx.shape
        dev = cw.device
        N, L, E, H = s.N, s.L, s.E, s.H
        d_news_c = torch.zeros(N, H, dtype=torch.float32, device=dev) if d_news is None else _f32c(d_news).view(N, H)
        d_c_c = None if d_c is None else _f32c(d_c).view(N * L, H)
        d_cw = torch.empty(H, E, 3, dtype=torch.float32, device=dev)
        d_cb = torch.empty(H, dtype=torch.float32, device=dev)
        d_pw = torch.empty(H, H, dtype=torch.float32, device=dev)
        d_pb = torch.empty(H, dtype=torch.float32, device=dev)
        d_q = torch.empty(H, dtype=torch.float32, device=dev)
        need_x = ctx.need_table_grad or ctx.need_emb_grad
        d_table = None
        d_emb_out = None
        if (ctx.need_table_grad and s.precision == MR_BF16 and d_c_c is None and E % 4 == 0 and GROUPED_TABLE_GRAD
                and tab.shape[0] >= pad_to(s.V, 32)):
            # token-grouped backward: the table gradient comes out of the encoder backward directly
            V, _ = ctx.table_shape
            d_table = torch.empty(V, E, dtype=torch.float32, device=dev)
            ws = workspace(lib.mr_news_cnn_bwd_table_workspace_bytes(byref(s)), dev)
            plan = ctx.group_plan
            if plan is not None:
                torch.cuda.current_stream(dev).wait_event(ctx.group_event)
            hook = TABLE_GRAD_HOOK
            ev = hook.event(dev) if hook is not None else None
            check(lib.mr_news_cnn_bwd_table(byref(s), ptr(ids_c), index_flag(ids_c), ptr(tab), tab.shape[0], ptr(cw), ptr(pw),
                                            ptr(q), ptr(c_save), ptr(key_save), ptr(prob), ptr(d_news_c), ptr(d_cw), ptr(d_cb),
                                            ptr(d_pw), ptr(d_pb), ptr(d_q), ptr(d_table), ctx.padding_idx, ptr(plan),
                                            plan.numel() if plan is not None else 0,
                                            c_void_p(ev.cuda_event) if ev is not None else None, ptr(ws), ws.numel(),
                                            stream_ptr(dev)), "mr_news_cnn_bwd_table")
            ctx.group_plan = None
            if hook is not None and hook(d_table, ev):
                d_table = None               # the hook took the gradient (it assigns table.grad itself, see trainer.GradSync)
            return (None, None, None, d_table, None, d_cw, d_cb, d_pw, d_pb, d_q.view(1, H), None, None, None)
        d_emb = None
        if need_x:
            if s.precision == MR_BF16:      # row pitch = E rounded up to 16 (16-byte aligned tensor-core epilogue stores)
                d_emb = torch.empty(N * L, pad_to(E, 16), dtype=torch.bfloat16, device=dev)
            else:
                d_emb = torch.empty(N * L, E, dtype=torch.float32, device=dev)
        ws = workspace(lib.mr_news_cnn_workspace_bytes(byref(s), 1), dev)
        check(lib.mr_news_cnn_bwd(byref(s), ptr(ids_c), index_flag(ids_c) if ids_c is not None else 0, ptr(emb_c),
                                  ptr(tab), ptr(cw), ptr(pw), ptr(q), ptr(c_save), ptr(key_save), ptr(prob),
                                  ptr(d_news_c), ptr(d_c_c), ptr(d_cw), ptr(d_cb), ptr(d_pw), ptr(d_pb), ptr(d_q),
                                  ptr(d_emb), ptr(ws), ws.numel(), stream_ptr(dev)), "mr_news_cnn_bwd")
        if ctx.need_table_grad:
            V, _ = ctx.table_shape
            d_table = embed_grad(ids_c.view(-1), d_emb, V, E, ctx.padding_idx)
        elif ctx.need_emb_grad:
            d_emb_out = d_emb[:, :E].float().view(*emb_c.shape) if d_emb.dtype != torch.float32 else d_emb.view(*emb_c.shape)
        return (None, d_emb_out, None, d_table, None, d_cw, d_cb, d_pw, d_pb, d_q.view(1, H), None, None, None)


# --------------------------------------------------------------------------------------------
# recurrent user encoders                                           models/Encoders/RNN.py:36-104
# --------------------------------------------------------------------------------------------
class RNNUser(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lens, h0, w_ih, w_hh, b_ih, b_hh, kind, reverse, precision):
        lib = _lib.load()
        dev = x.device
        xc = _f32c(x)
        B, S, H = xc.shape
        G = 4 if kind == _lib.MR_RNN_LSTM else 3
        shape = RnnShape(B, S, H, kind, 1 if reverse else 0, precision)
        lens_c = None if lens is None else lens.to(device=dev, dtype=torch.int32).contiguous()
        h0_c = None if h0 is None else _f32c(h0)
        wi, wh, bi, bh = _f32c(w_ih), _f32c(w_hh), _f32c(b_ih), _f32c(b_hh)
        gates = torch.empty(B, S, G * H, dtype=torch.float32, device=dev)
        hs = torch.empty(B, S, H, dtype=torch.float32, device=dev)
        cs = torch.empty(B, S, H, dtype=torch.float32, device=dev)
        user = torch.empty(B, H, dtype=torch.float32, device=dev)
        ws = workspace(lib.mr_rnn_workspace_bytes(byref(shape), 0), dev)
        check(lib.mr_rnn_user_fwd(byref(shape), ptr(xc), ptr(lens_c), ptr(h0_c), ptr(wi), ptr(wh), ptr(bi), ptr(bh),
                                  ptr(gates), ptr(hs), ptr(cs), ptr(user), ptr(ws), ws.numel(), stream_ptr(dev)),
              "mr_rnn_user_fwd")
        ctx.save_for_backward(xc, lens_c, h0_c, wi, wh, gates, hs, cs)
        ctx.shape = shape
        ctx.need_h0 = h0 is not None and h0.requires_grad
        return user.view(B, 1, H)

    @staticmethod
    def backward(ctx, d_user):
        lib = _lib.load()
        xc, lens_c, h0_c, wi, wh, gates, hs, cs = ctx.saved_tensors
        s = ctx.shape
        dev = xc.device
        B, S, H = s.B, s.S, s.H
        G = 4 if s.kind == _lib.MR_RNN_LSTM else 3
        du = _f32c(d_user).view(B, H)
        d_x = torch.empty(B, S, H, dtype=torch.float32, device=dev)
        d_h0 = torch.empty(B, H, dtype=torch.float32, device=dev) if ctx.need_h0 else None
        d_wi = torch.empty(G * H, H, dtype=torch.float32, device=dev)
        d_wh = torch.empty(G * H, H, dtype=torch.float32, device=dev)
        d_bi = torch.empty(G * H, dtype=torch.float32, device=dev)
        d_bh = torch.empty(G * H, dtype=torch.float32, device=dev)
        ws = workspace(lib.mr_rnn_workspace_bytes(byref(s), 1), dev)
        check(lib.mr_rnn_user_bwd(byref(s), ptr(xc), ptr(lens_c), ptr(h0_c), ptr(wi), ptr(wh), ptr(gates), ptr(hs),
                                  ptr(cs), ptr(du), ptr(d_x), ptr(d_h0), ptr(d_wi), ptr(d_wh), ptr(d_bi), ptr(d_bh),
                                  ptr(ws), ws.numel(), stream_ptr(dev)), "mr_rnn_user_bwd")
        return d_x, None, d_h0, d_wi, d_wh, d_bi, d_bh, None, None, None


# --------------------------------------------------------------------------------------------
# pooling user encoders                                             models/Encoders/Pooling.py
# --------------------------------------------------------------------------------------------
class AttnPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r, mask, query):
        lib = _lib.load()
        rc = _f32c(r)
        B, S, H = rc.shape
        dev = rc.device
        mc = None if mask is None else mask.to(device=dev, dtype=torch.float32).reshape(B, S).contiguous()
        q = _f32c(query).view(-1)
        prob = torch.empty(B, S, dtype=torch.float32, device=dev)
        out = torch.empty(B, H, dtype=torch.float32, device=dev)
        check(lib.mr_attnpool_fwd(ptr(rc), ptr(mc), ptr(q), ptr(prob), ptr(out), B, S, H, stream_ptr(dev)), "mr_attnpool_fwd")
        ctx.save_for_backward(rc, q, prob)
        return out.view(B, 1, H)

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        rc, q, prob = ctx.saved_tensors
        B, S, H = rc.shape
        dev = rc.device
        go = _f32c(d_out).view(B, H)
        d_r = torch.empty_like(rc)
        dq_part = torch.empty(B, H, dtype=torch.float32, device=dev)
        check(lib.mr_attnpool_bwd(ptr(rc), ptr(q), ptr(prob), ptr(go), ptr(d_r), ptr(dq_part), B, S, H, stream_ptr(dev)),
              "mr_attnpool_bwd")
        return d_r, None, column_sum(dq_part).view(1, H)


class AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r):
        lib = _lib.load()
        rc = _f32c(r)
        B, S, H = rc.shape
        out = torch.empty(B, H, dtype=torch.float32, device=rc.device)
        check(lib.mr_avgpool_fwd(ptr(rc), ptr(out), B, S, H, stream_ptr(rc.device)), "mr_avgpool_fwd")
        ctx.shape = (B, S, H)
        return out.view(B, 1, H)

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        B, S, H = ctx.shape
        go = _f32c(d_out).view(B, H)
        d_r = torch.empty(B, S, H, dtype=torch.float32, device=go.device)
        check(lib.mr_avgpool_bwd(ptr(go), ptr(d_r), B, S, H, stream_ptr(go.device)), "mr_avgpool_bwd")
        return d_r


def column_sum(x: torch.Tensor) -> torch.Tensor:
    """[R,C] -> [C] through the library's linear-layer bias-gradient path (fixed order)."""
    lib = _lib.load()
    R, C = x.shape
    out = torch.empty(C, dtype=torch.float32, device=x.device)
    ws = workspace(lib.mr_linear_workspace_bytes(R, C, 1), x.device)
    check(lib.mr_linear_bwd(None, None, ptr(x), None, None, ptr(out), R, C, 1, MR_F32, ptr(ws), ws.numel(),
                            stream_ptr(x.device)), "mr_linear_bwd(colsum)")
    return out


# --------------------------------------------------------------------------------------------
# scoring                                                  models/TwoTowerBaseModel.py:51-75
# --------------------------------------------------------------------------------------------
class ScoreLogSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cdd, user):
        lib = _lib.load()
        cc, uc = _f32c(cdd), _f32c(user)
        B, C, H = cc.shape
        logp = torch.empty(B, C, dtype=torch.float32, device=cc.device)
        check(lib.mr_score_logsoftmax_fwd(ptr(cc), ptr(uc), ptr(logp), None, 0, None, B, C, H, stream_ptr(cc.device)),
              "mr_score_logsoftmax_fwd")
        ctx.save_for_backward(cc, uc, logp)
        return logp

    @staticmethod
    def backward(ctx, d_logp):
        lib = _lib.load()
        cc, uc, logp = ctx.saved_tensors
        B, C, H = cc.shape
        g = _f32c(d_logp)
        d_cdd = torch.empty_like(cc)
        d_user = torch.empty(B, 1, H, dtype=torch.float32, device=cc.device)
        check(lib.mr_score_logsoftmax_bwd(ptr(cc), ptr(uc), ptr(logp), ptr(g), ptr(d_cdd), ptr(d_user), B, C, H,
                                          stream_ptr(cc.device)), "mr_score_logsoftmax_bwd")
        return d_cdd, d_user


class NLLMean(torch.autograd.Function):
    """nn.NLLLoss() (mean) on log-probabilities [B,C] (utils/Manager.py:381-382,641)."""

    @staticmethod
    def forward(ctx, logp, label):
        lib = _lib.load()
        lc = _f32c(logp)
        lab = _idx(label.to(lc.device))
        B, C = lc.shape
        loss = torch.empty(1, dtype=torch.float32, device=lc.device)
        check(lib.mr_nll_loss_fwd(ptr(lc), ptr(lab), index_flag(lab), ptr(loss), B, C, stream_ptr(lc.device)), "mr_nll_loss_fwd")
        ctx.save_for_backward(lab)
        ctx.dims = (B, C)
        return loss.view(())

    @staticmethod
    def backward(ctx, d_loss):
        lib = _lib.load()
        (lab,) = ctx.saved_tensors
        B, C = ctx.dims
        g = _f32c(d_loss).view(1)
        d_logp = torch.empty(B, C, dtype=torch.float32, device=g.device)
        check(lib.mr_nll_loss_bwd(ptr(lab), index_flag(lab), ptr(g), ptr(d_logp), B, C, stream_ptr(g.device)), "mr_nll_loss_bwd")
        return d_logp, None


def score_sigmoid(cdd: torch.Tensor, user: torch.Tensor, apply_sigmoid: bool = True) -> torch.Tensor:
    """sigmoid(<cdd, user>/sqrt(H)), eval only (TwoTowerBaseModel.py:72-73,83); raw scores when
    apply_sigmoid is False (compute_score, TwoTowerBaseModel.py:61)."""
    lib = _lib.load()
    cc, uc = _f32c(cdd), _f32c(user)
    B, C, H = cc.shape
    prob = torch.empty(B, C, dtype=torch.float32, device=cc.device)
    check(lib.mr_score_sigmoid_fwd(ptr(cc), ptr(uc), ptr(prob), B, C, H, 1 if apply_sigmoid else 0,
                                   stream_ptr(cc.device)), "mr_score_sigmoid_fwd")
    return prob


def score_sigmoid_gather(table: torch.Tensor, cdd_id: torch.Tensor, offsets: torch.Tensor, user: torch.Tensor) -> torch.Tensor:
    """Batched fast-eval scoring over CSR impressions (TwoTowerBaseModel.py:78-84)."""
    lib = _lib.load()
    tc, uc = _f32c(table), _f32c(user).view(-1, table.shape[1])
    ids = _idx(cdd_id).view(-1)
    off = offsets.to(device=tc.device, dtype=torch.int64).contiguous()
    n_impr = off.numel() - 1
    prob = torch.empty(ids.numel(), dtype=torch.float32, device=tc.device)
    check(lib.mr_score_sigmoid_gather_fwd(ptr(tc), ptr(ids), index_flag(ids), ptr(off), ptr(uc), ptr(prob), n_impr,
                                          ids.numel(), tc.shape[0], tc.shape[1], stream_ptr(tc.device)),
          "mr_score_sigmoid_gather_fwd")
    return prob


def rank_metrics(prob: torch.Tensor, label: torch.Tensor, offsets: torch.Tensor, want_rank: bool = False):
    """Per-impression (auc, mrr, ndcg@5, ndcg@10) in fp64 and optional ordinal ranks
    (utils/Manager.py:1205-1344, 842-850)."""
    lib = _lib.load()
    pc = _f32c(prob).view(-1)
    lc = label.to(device=pc.device, dtype=torch.float32).contiguous().view(-1)
    off = offsets.to(device=pc.device, dtype=torch.int64).contiguous()
    n_impr = off.numel() - 1
    metrics = torch.empty(n_impr, 4, dtype=torch.float64, device=pc.device)
    rank = torch.empty(pc.numel(), dtype=torch.int32, device=pc.device) if want_rank else None
    check(lib.mr_rank_metrics(ptr(pc), ptr(lc), ptr(off), ptr(metrics), ptr(rank), n_impr, pc.numel(),
                              stream_ptr(pc.device)), "mr_rank_metrics")
    return metrics, rank


# --------------------------------------------------------------------------------------------
# optimiser                                                         utils/Manager.py:404-413
# --------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, shadow=None):
    lib = _lib.load()
    n = p.numel()
    row_len, ld = (0, 0)
    if shadow is not None:
        row_len, ld = p.shape[-1], shadow.shape[-1]
    check(lib.mr_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), n, step, float(lr), float(beta1), float(beta2),
                           float(eps), float(grad_scale), ptr(shadow), row_len, ld, stream_ptr(p.device)),
          "mr_adam_step")


def cast_pad_bf16(src: torch.Tensor, ld: int, extra_rows: int = 0, out: torch.Tensor = None) -> torch.Tensor:
    lib = _lib.load()
    sc = _f32c(src)
    rows, cols = sc.shape
    if out is not None:
        if out.dtype != torch.bfloat16 or tuple(out.shape) != (rows, ld) or not out.is_contiguous():
            raise ValueError("cast_pad_bf16: out must be a contiguous bf16 [%d, %d] tensor" % (rows, ld))
        dst = out
    else:
        dst = torch.empty(rows + extra_rows, ld, dtype=torch.bfloat16, device=sc.device)
        if extra_rows:
            dst[rows:].zero_()
    check(lib.mr_cast_pad_bf16(ptr(sc), ptr(dst), rows, cols, ld, stream_ptr(sc.device)), "mr_cast_pad_bf16")
    return dst


# --------------------------------------------------------------------------------------------
# MHA building blocks                          models/Modules/Attention.py:83-147, Encoders/MHA.py
# --------------------------------------------------------------------------------------------
class Linear(torch.autograd.Function):
    """y = act(x W^T + b)  (act: 0 none)."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        lib = _lib.load()
        xc, wc = _f32c(x), _f32c(w)
        bc = None if b is None else _f32c(b)
        N, K = wc.shape
        M = xc.numel() // K
        y = torch.empty(*x.shape[:-1], N, dtype=torch.float32, device=xc.device)
        check(lib.mr_linear_fwd(ptr(xc), ptr(wc), ptr(bc), ptr(y), M, N, K, act, MR_F32, stream_ptr(xc.device)),
              "mr_linear_fwd")
        if act != 0:
            raise RuntimeError("ops.Linear only differentiates act=0")
        ctx.save_for_backward(xc, wc)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, d_y):
        lib = _lib.load()
        xc, wc = ctx.saved_tensors
        N, K = wc.shape
        M = xc.numel() // K
        g = _f32c(d_y).view(M, N)
        d_x = torch.empty_like(xc)
        d_w = torch.empty_like(wc)
        d_b = torch.empty(N, dtype=torch.float32, device=xc.device) if ctx.has_bias else None
        ws = workspace(lib.mr_linear_workspace_bytes(M, N, K), xc.device)
        check(lib.mr_linear_bwd(ptr(xc), ptr(wc), ptr(g), ptr(d_x), ptr(d_w), ptr(d_b), M, N, K, MR_F32, ptr(ws),
                                ws.numel(), stream_ptr(xc.device)), "mr_linear_bwd")
        return d_x, d_w, d_b, None


def _linear_tc_fwd(x, ids, table, table_bf16, w, b):
    """-> (y [M, ldy] fp32 padded, saved state) -- mr_linear_tc_fwd"""
    lib = _lib.load()
    wc = _f32c(w)
    bc = None if b is None else _f32c(b)
    N, K = wc.shape
    dev = wc.device
    if x is not None:
        xc = _f32c(x).reshape(-1, K)
        M, V, ids_c, tab = xc.shape[0], 0, None, None
    else:
        ids_c = _idx(ids).reshape(-1)
        xc, tab = None, table_bf16
        M, V = ids_c.numel(), table.shape[0]
        if table.shape[1] != K or tab.dtype != torch.bfloat16:
            raise ValueError("LinearTC: table width %d != K %d or shadow not bf16" % (table.shape[1], K))
    ldy = pad_to(N, 4)
    y = torch.empty(M, ldy, dtype=torch.float32, device=dev)
    ws = workspace(lib.mr_linear_tc_workspace_bytes(M, N, K, V, 0), dev)
    check(lib.mr_linear_tc_fwd(ptr(xc), ptr(ids_c), index_flag(ids_c) if ids_c is not None else 0, ptr(tab),
                               tab.shape[1] if tab is not None else 0, V, ptr(wc), ptr(bc), ptr(y), ldy, M, N, K, ptr(ws), ws.numel(),
                               stream_ptr(dev)), "mr_linear_tc_fwd")
    return y, (xc, ids_c, tab, wc), (M, N, K, V, ldy)


def _linear_tc_bwd(saved, dims, d_y, has_bias, need_x):
    """-> (d_x | None, d_table | None, d_w, d_b | None) -- mr_linear_tc_bwd; d_y [M, ldy] fp32 contiguous"""
    lib = _lib.load()
    xc, ids_c, tab, wc = saved
    M, N, K, V, ldy = dims
    dev = wc.device
    d_w = torch.empty(N, K, dtype=torch.float32, device=dev)
    d_b = torch.empty(N, dtype=torch.float32, device=dev) if has_bias else None
    d_x = torch.empty(M, K, dtype=torch.float32, device=dev) if (need_x and xc is not None) else None
    d_table = torch.empty(V, K, dtype=torch.float32, device=dev) if xc is None else None
    ws = workspace(lib.mr_linear_tc_workspace_bytes(M, N, K, V, 1), dev)
    check(lib.mr_linear_tc_bwd(ptr(xc), ptr(ids_c), index_flag(ids_c) if ids_c is not None else 0, ptr(tab),
                               tab.shape[1] if tab is not None else 0, tab.shape[0] if tab is not None else 0, V, 0, ptr(wc), ptr(d_y), ldy,
                               ptr(d_x), ptr(d_table), ptr(d_w), ptr(d_b), M, N, K, ptr(ws), ws.numel(), stream_ptr(dev)),
          "mr_linear_tc_bwd")
    return d_x, d_table, d_w, d_b


class LinearTC(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 tensor cores (bf16 operands, fp32 accumulate / output) -- the MR_BF16 path of the attention
    projections.  Input: dense ``x`` [M, K] fp32, or (``x`` None) token ``ids`` [M] whose rows of the bf16 table shadow are gathered
    inside the GEMM (then ``table`` is the fp32 master [V, K] that receives the gradient, ``table_bf16`` its padded shadow).
    Returns the PADDED output [M, ldy], ldy = N rounded up to 4 (16-byte rows for the vector epilogue); slice what you need."""

    @staticmethod
    def forward(ctx, x, ids, table, table_bf16, w, b):
        y, saved, dims = _linear_tc_fwd(x, ids, table, table_bf16, w, b)
        ctx.save_for_backward(*saved)
        ctx.dims = dims
        ctx.has_bias = b is not None
        ctx.x_shape = None if x is None else tuple(x.shape)
        ctx.need_x = x is not None and x.requires_grad
        ctx.need_table = table is not None and table.requires_grad
        return y

    @staticmethod
    def backward(ctx, d_y):
        d_x, d_table, d_w, d_b = _linear_tc_bwd(ctx.saved_tensors, ctx.dims, _f32c(d_y), ctx.has_bias, ctx.need_x)
        if d_x is not None:
            d_x = d_x.view(ctx.x_shape)
        return d_x, None, (d_table if ctx.need_table else None), None, d_w, d_b


class MHABlock(torch.autograd.Function):
    """MR_BF16 multi-head self-attention block (Attention.py:115-147): [keyProject; valueProject] as ONE tcgen05 GEMM
    (mr_linear_tc_*) followed by the register-tiled attention core (mr_mha_attn_*).  The q|k and v slices of the projection
    output are read in place (row pitch), and the backward writes d_qk / d_v straight into the projection gradient: no slice
    copies, no concatenation.  Input: dense x [n, len, K], or token ids [n, len] (rows gathered from the bf16 table)."""

    @staticmethod
    def forward(ctx, x, ids, table, table_bf16, w, b, mask, head_num, nk, nv):
        lib = _lib.load()
        n, length = (x.shape[0], x.shape[1]) if x is not None else tuple(ids.shape)
        y, saved, dims = _linear_tc_fwd(x, ids, table, table_bf16, w, b)
        dev, ldy = y.device, dims[4]
        dk, dv = nk // head_num, nv // head_num
        mc = None if mask is None else mask.to(device=dev, dtype=torch.float32).reshape(n, length).contiguous()
        prob = torch.empty(n, head_num, length, length, dtype=torch.float32, device=dev)
        out = torch.empty(n, length, nv, dtype=torch.float32, device=dev)
        check(lib.mr_mha_attn_fwd(ptr(y), ldy, c_void_p(y.data_ptr() + 4 * nk), ldy, ptr(mc), ptr(prob), ptr(out), n, length, head_num,
                                  dk, dv, stream_ptr(dev)), "mr_mha_attn_fwd")
        ctx.save_for_backward(*saved, y, prob)
        ctx.dims, ctx.geom = dims, (n, length, head_num, dk, dv, nk, nv)
        ctx.has_bias = b is not None
        ctx.x_shape = None if x is None else tuple(x.shape)
        ctx.need_x = x is not None and x.requires_grad
        ctx.need_table = table is not None and table.requires_grad
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        *saved, y, prob = ctx.saved_tensors
        n, length, hn, dk, dv, nk, nv = ctx.geom
        ldy = ctx.dims[4]
        g = _f32c(d_out)
        d_y = torch.empty_like(y)                       # columns >= nk + nv are padding: never read by mr_linear_tc_bwd
        check(lib.mr_mha_attn_bwd(ptr(y), ldy, c_void_p(y.data_ptr() + 4 * nk), ldy, ptr(prob), ptr(g), ptr(d_y), ldy,
                                  c_void_p(d_y.data_ptr() + 4 * nk), ldy, n, length, hn, dk, dv, stream_ptr(y.device)), "mr_mha_attn_bwd")
        d_x, d_table, d_w, d_b = _linear_tc_bwd(tuple(saved), ctx.dims, d_y, ctx.has_bias, ctx.need_x)
        if d_x is not None:
            d_x = d_x.view(ctx.x_shape)
        return d_x, None, (d_table if ctx.need_table else None), None, d_w, d_b, None, None, None, None


class MHACore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qk, v, mask, head_num):
        lib = _lib.load()
        qc, vc = _f32c(qk), _f32c(v)
        n, length, dkh = qc.shape
        dk, dv = dkh // head_num, vc.shape[-1] // head_num
        mc = None if mask is None else mask.to(device=qc.device, dtype=torch.float32).reshape(n, length).contiguous()
        prob = torch.empty(n, head_num, length, length, dtype=torch.float32, device=qc.device)
        out = torch.empty(n, length, head_num * dv, dtype=torch.float32, device=qc.device)
        check(lib.mr_mha_core_fwd(ptr(qc), ptr(vc), ptr(mc), ptr(prob), ptr(out), n, length, head_num, dk, dv,
                                  stream_ptr(qc.device)), "mr_mha_core_fwd")
        ctx.save_for_backward(qc, vc, prob)
        ctx.dims = (n, length, head_num, dk, dv)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        qc, vc, prob = ctx.saved_tensors
        n, length, hn, dk, dv = ctx.dims
        g = _f32c(d_out)
        d_qk, d_v = torch.empty_like(qc), torch.empty_like(vc)
        check(lib.mr_mha_core_bwd(ptr(qc), ptr(vc), ptr(prob), ptr(g), ptr(d_qk), ptr(d_v), n, length, hn, dk, dv,
                                  stream_ptr(qc.device)), "mr_mha_core_bwd")
        return d_qk, d_v, None, None


class LayerNorm(torch.autograd.Function):
    N_PARTIAL = 296

    @staticmethod
    def forward(ctx, x, gamma, beta, keep, keep_scale):
        lib = _lib.load()
        xc, gc, bc = _f32c(x), _f32c(gamma), _f32c(beta)
        H = xc.shape[-1]
        M = xc.numel() // H
        kc = None if keep is None else keep.contiguous()
        y = torch.empty_like(xc)
        mean = torch.empty(M, dtype=torch.float32, device=xc.device)
        rstd = torch.empty(M, dtype=torch.float32, device=xc.device)
        check(lib.mr_layernorm_fwd(ptr(xc), ptr(gc), ptr(bc), ptr(kc), c_float(keep_scale), ptr(y), ptr(mean),
                                   ptr(rstd), M, H, stream_ptr(xc.device)), "mr_layernorm_fwd")
        ctx.save_for_backward(xc, gc, kc, mean, rstd)
        ctx.keep_scale = keep_scale
        return y

    @staticmethod
    def backward(ctx, d_y):
        lib = _lib.load()
        xc, gc, kc, mean, rstd = ctx.saved_tensors
        H = xc.shape[-1]
        M = xc.numel() // H
        g = _f32c(d_y)
        d_x = torch.empty_like(xc)
        npart = LayerNorm.N_PARTIAL
        dg = torch.empty(npart, H, dtype=torch.float32, device=xc.device)
        db = torch.empty(npart, H, dtype=torch.float32, device=xc.device)
        check(lib.mr_layernorm_bwd(ptr(xc), ptr(gc), ptr(kc), c_float(ctx.keep_scale), ptr(mean), ptr(rstd), ptr(g),
                                   ptr(d_x), ptr(dg), ptr(db), npart, M, H, stream_ptr(xc.device)), "mr_layernorm_bwd")
        return d_x, column_sum(dg), column_sum(db), None, None

"""Diagnosis of an order-dependent failure of the bf16 LSTM user encoder at B=700 (8 sequences per CTA): calls mr_rnn_user_fwd
directly with a POISONED workspace and saved-tensor buffers and reports where the hidden states differ from the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from ctypes import byref
from news_recommendation_mind_b200 import _lib
from news_recommendation_mind_b200._lib import RnnShape, ptr, stream_ptr, check
from oracle import twotower_oracle as O

B, S, H = int(os.environ.get("B", 700)), int(os.environ.get("S", 9)), 150
lib = _lib.load()
gen = torch.Generator().manual_seed(3)
x = torch.randn(B, S, H, generator=gen) * 0.5
ln = torch.randint(1, S + 1, (B,), generator=gen); ln[0] = S
w_ih, w_hh = torch.randn(4 * H, H, generator=gen) * 0.08, torch.randn(4 * H, H, generator=gen) * 0.08
b_ih, b_hh = torch.randn(4 * H, generator=gen) * 0.1, torch.randn(4 * H, generator=gen) * 0.1
his_mask = (torch.arange(S)[None, :] < ln[:, None]).double().unsqueeze(-1)
# oracle hidden states per step
xo, wi, wh = x.bfloat16().double(), w_ih.bfloat16().double(), w_hh.bfloat16().double()
h = torch.zeros(B, H, dtype=torch.float64); c = torch.zeros(B, H, dtype=torch.float64)
hs_ref = torch.zeros(B, S, H, dtype=torch.float64)
for t in range(S):
    g = xo[:, t] @ wi.t() + b_ih.double() + h @ wh.t() + b_hh.double()
    i, f, gg, o = g.split(H, dim=1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h = torch.sigmoid(o) * torch.tanh(c)
    hs_ref[:, t] = h
shape = RnnShape(B, S, H, 0, 0, 1)
nws = lib.mr_rnn_workspace_bytes(byref(shape), 0)
xd, lens = x.cuda(), ln.to(torch.int32).cuda()
wd = [t.cuda() for t in (w_ih, w_hh, b_ih, b_hh)]
for poison in (float("nan"), 3.0, float("nan"), 3.0, 0.0):
    for trial in range(3):
        ws = torch.full((nws // 4 + 64,), poison, dtype=torch.float32, device="cuda")
        gates = torch.full((B, S, 4 * H), float("nan"), device="cuda"); hs = torch.full((B, S, H), float("nan"), device="cuda")
        cs = torch.full((B, S, H), float("nan"), device="cuda"); user = torch.full((B, H), float("nan"), device="cuda")
        torch.cuda.synchronize()
        check(lib.mr_rnn_user_fwd(byref(shape), ptr(xd), ptr(lens), None, ptr(wd[0]), ptr(wd[1]), ptr(wd[2]), ptr(wd[3]), ptr(gates), ptr(hs), ptr(cs),
                                  ptr(user), ptr(ws), ws.numel() * 4, stream_ptr("cuda:0")), "fwd")
        torch.cuda.synchronize()
        valid = (torch.arange(S)[None, :] < ln[:, None])
        d = (hs.cpu().double() - hs_ref).abs().amax(-1)                       # [B, S]
        d = torch.where(valid, torch.nan_to_num(d, nan=1e9), torch.zeros_like(d))
        bad = d > 1e-3
        msg = "poison %-4s trial %d: bad (seq, step) pairs %d of %d" % (poison, trial, int(bad.sum()), int(valid.sum()))
        if bad.any():
            bs, ss = torch.nonzero(bad, as_tuple=True)
            msg += " | by step %s | by slot b%%8 %s | first %s" % (torch.bincount(ss, minlength=S).tolist(), torch.bincount(bs % 8, minlength=8).tolist(),
                                                              list(zip(bs[:6].tolist(), ss[:6].tolist())))
            b0, s0 = int(bs[0]), int(ss[0])
            row = (hs[b0, s0].cpu().double() - hs_ref[b0, s0]).abs()
            wrong = torch.nonzero(torch.nan_to_num(row, nan=1e9) > 1e-3).reshape(-1)
            msg += " | wrong units of (%d,%d): %d e.g. %s val %s" % (b0, s0, wrong.numel(), wrong[:10].tolist(), hs[b0, s0, wrong[:4]].tolist())
            gbad = torch.nonzero(~torch.isfinite(gates[b0, s0].cpu())).reshape(-1)
            msg += " | non-finite gates %d e.g. %s" % (gbad.numel(), gbad[:8].tolist())
        print(msg, flush=True)

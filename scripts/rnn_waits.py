"""Per-phase cycle counters of rnn_tc_bwd_kernel (csrc/rnn_tc.cu) at the bench shape (256 sequences x 50 steps, H = 150)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import manager_for
import news_recommendation_mind_b200 as mr
from news_recommendation_mind_b200 import _lib

lib = _lib.load()
lib.mr_debug_tapgemm_counters.argtypes = [ctypes.c_void_p]
B, S, H = 256, 50, 150
man = manager_for("cnn", "lstm", 5, S, 32, 300, H, 10, precision="bf16")
enc = mr.RNN_User_Encoder(man).cuda()
g = torch.Generator().manual_seed(0)
x = torch.randn(B, S, H, generator=g).cuda().requires_grad_(True)
lens = torch.randint(1, S + 1, (B,), generator=g)
mask = (torch.arange(S)[None, :] < lens[:, None]).double().unsqueeze(-1)
for _ in range(2):
    enc(x, his_mask=mask).sum().backward()
torch.cuda.synchronize()
buf = torch.zeros(16, 148, 4, 5, dtype=torch.int64, device="cuda")
out = enc(x, his_mask=mask).sum()
torch.cuda.synchronize()
lib.mr_debug_tapgemm_counters(ctypes.c_void_p(buf.data_ptr()))
out.backward()
torch.cuda.synchronize()
lib.mr_debug_tapgemm_counters(None)
b = buf.double().cpu()
for i in range(16):
    if b[i].abs().sum() == 0:
        continue
    r = b[i, :128, 0]
    print("launch %d: cycles avg %.0f; per step (50): gate math %.0f, B-tile stores %.0f, fence.proxy.async %.0f, __syncthreads %.0f" % (
        i, r[:, 4].mean(), r[:, 0].mean() / 50, r[:, 1].mean() / 50, r[:, 2].mean() / 50, r[:, 3].mean() / 50))

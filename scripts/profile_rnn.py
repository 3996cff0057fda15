"""Runs the bf16 LSTM user encoder forward+backward at the bench shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import manager_for
import news_recommendation_mind_b200 as mr
B, S, H = 256, 50, 150
man = manager_for("cnn", "lstm", 5, S, 32, 300, H, 10, precision="bf16")
enc = mr.RNN_User_Encoder(man).cuda()
rng = np.random.RandomState(0)
ln = np.clip(np.rint(rng.lognormal(3.0, 0.9, size=B)), 1, S).astype(np.int64)
his_mask = (torch.arange(S)[None, :] < torch.from_numpy(ln)[:, None]).double().unsqueeze(-1)
x = torch.randn(B, S, H, device="cuda", requires_grad=True)
for _ in range(3):
    enc(x, his_mask=his_mask).sum().backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    enc(x, his_mask=his_mask).sum().backward()
e1.record(); torch.cuda.synchronize()
print("ms per fwd+bwd: %.3f" % (e0.elapsed_time(e1) / 10))

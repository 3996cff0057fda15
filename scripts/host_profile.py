"""cProfile of the host side of the training step (what the CPU spends its ~1.2 ms per step on)."""
import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import news_recommendation_mind_b200 as mr
from news_recommendation_mind_b200 import data, trainer
CFG = bench.CFG
dev = "cuda:0"
torch.manual_seed(42)
man = bench.manager_ns(dev, "bf16")
model = mr.TwoTower(man, mr.BERT_Embedding(man, vocab_size=CFG["V"]), mr.CNN_Encoder(man), mr.RNN_User_Encoder(man)).to(dev)
opt = trainer.FusedAdam(model, lr=1e-4, bert_lr=6e-6)
ids, mask = data.make_news_table(CFG["n_news"], CFG["L"])
devb = [{k: v.to(dev) for k, v in data.make_train_batch(ids, mask, CFG["B"], CFG["C"], CFG["S"], seed=i).items()} for i in range(4)]
for s in range(5):
    trainer.train_step(model, devb[s % 4], opt)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for s in range(50):
    trainer.train_step(model, devb[s % 4], opt)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)

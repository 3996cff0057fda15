"""Eager vs CUDA-graph replay of the training step (bench config): ms/step and host ms/step."""
import os, sys, time
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
def run(fn, steps=40):
    for s in range(5):
        fn(devb[s % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for s in range(steps):
        fn(devb[s % 4])
    th = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, 1e3 * th / steps
print("eager: %.3f ms/step (host %.3f)" % run(lambda x: trainer.train_step(model, x, opt)))
gs = trainer.GraphStep(model, opt, devb[0])
print("graph: %.3f ms/step (host %.3f)" % run(gs))
print("graph: %.3f ms/step (host %.3f)" % run(gs))

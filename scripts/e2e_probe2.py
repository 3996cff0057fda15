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
host = [data.make_train_batch(ids, mask, CFG["B"], CFG["C"], CFG["S"], seed=i, pin=True) for i in range(4)]
devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
for s in range(5):
    trainer.train_step(model, devb[s % 4], opt)
torch.cuda.synchronize()
steps = 30
def run(stage, read):
    pf = trainer.BatchPrefetcher(dev)
    slots = [torch.empty(1).pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    staged = pf.stage(host[0]) if stage else None
    for s in range(steps):
        if stage:
            x = pf.take(staged)
            staged = pf.stage(host[(s + 1) % 4])
        else:
            x = devb[s % 4]
        loss = trainer.train_step(model, x, opt)
        if read:
            slots[s % 2].copy_(loss.detach().reshape(1), non_blocking=True)
            evs[s % 2].record()
            if s >= 1:
                evs[(s - 1) % 2].synchronize()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / steps
for stage in (False, True):
    for read in (False, True):
        print("stage=%s read=%s: %.3f ms/step" % (stage, read, run(stage, read)))
# staging only the big tensors synchronously before the step (no side stream)
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(steps):
    x = {k: v.to(dev, non_blocking=True) for k, v in host[s % 4].items()}
    trainer.train_step(model, x, opt)
torch.cuda.synchronize()
print("same-stream async copies: %.3f ms/step" % (1e3 * (time.perf_counter() - t0) / steps))

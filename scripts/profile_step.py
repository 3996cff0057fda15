"""One bf16 training step at the bench shape between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
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
devb = [{k: v.to(dev) for k, v in data.make_train_batch(ids, mask, CFG["B"], CFG["C"], CFG["S"], seed=i).items()} for i in range(2)]
for s in range(3):
    trainer.train_step(model, devb[s % 2], opt)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for s in range(int(os.environ.get("PSTEPS", "1"))):
    trainer.train_step(model, devb[s % 2], opt)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

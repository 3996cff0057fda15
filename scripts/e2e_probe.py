"""Where does the end-to-end (host batches) step spend host time?  Times stage / take / train_step / loss read."""
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
for s in range(5):
    trainer.train_step(model, {k: v.to(dev) for k, v in host[s % 4].items()}, opt)
torch.cuda.synchronize()
for skip in (("his_mask",), ()):
    pf = trainer.BatchPrefetcher(dev)
    def stage(x):
        with torch.cuda.stream(pf.stream):
            out = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) and k not in skip else v) for k, v in x.items()}
            ev = torch.cuda.Event(); ev.record(pf.stream)
        return out, ev
    T = dict(stage=0.0, take=0.0, step=0.0, read=0.0)
    steps = 20
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    staged = stage(host[0]); pending = None
    for s in range(steps):
        t0 = time.perf_counter(); x = pf.take(staged); t1 = time.perf_counter()
        staged = stage(host[(s + 1) % 4]); t2 = time.perf_counter()
        loss = trainer.train_step(model, x, opt); t3 = time.perf_counter()
        if pending is not None:
            float(pending.detach())
        t4 = time.perf_counter()
        pending = loss
        T["take"] += t1 - t0; T["stage"] += t2 - t1; T["step"] += t3 - t2; T["read"] += t4 - t3
    float(pending.detach())
    torch.cuda.synchronize()
    tot = time.perf_counter() - t_all
    print("skip=%s: %.3f ms/step; host ms/step: %s" % (skip, 1e3 * tot / steps, {k: round(1e3 * v / steps, 3) for k, v in T.items()}))

# kernel timeline of the staged loop: gaps between consecutive kernels on the device
from torch.profiler import profile, ProfilerActivity
pf = trainer.BatchPrefetcher(dev)
staged = pf.stage(host[0]); pending = None
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for s in range(6):
        x = pf.take(staged)
        staged = pf.stage(host[(s + 1) % 4])
        loss = trainer.train_step(model, x, opt)
        if pending is not None:
            float(pending.detach())
        pending = loss
    torch.cuda.synchronize()
evs = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA])
t0 = evs[0][0]
busy = sum(b - a for a, b, _ in evs)
print("span %.1f us, busy %.1f us, events %d" % (evs[-1][1] - t0, busy, len(evs)))
gaps = sorted([(evs[i + 1][0] - max(e[1] for e in evs[:i + 1][-3:]), evs[i][2][:50], evs[i + 1][2][:50]) for i in range(len(evs) - 1)], reverse=True)
for g in gaps[:14]:
    print("gap %8.1f us after %-50s before %s" % g)

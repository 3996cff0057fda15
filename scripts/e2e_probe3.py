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
pf = trainer.BatchPrefetcher(dev)
T = dict(take=0.0, stage=0.0, step=0.0, rel=0.0)
torch.cuda.synchronize()
t_all = time.perf_counter()
staged = pf.stage(host[0])
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(steps):
    t0 = time.perf_counter(); cur = staged; x = pf.take(cur); t1 = time.perf_counter()
    staged = pf.stage(host[(s + 1) % 4]); t2 = time.perf_counter()
    loss = trainer.train_step(model, x, opt); t3 = time.perf_counter()
    pf.release(cur); t4 = time.perf_counter()
    T["take"] += t1 - t0; T["stage"] += t2 - t1; T["step"] += t3 - t2; T["rel"] += t4 - t3
e1.record()
t_enq = time.perf_counter() - t_all
torch.cuda.synchronize()
tot = time.perf_counter() - t_all
print("staged loop: wall %.3f ms/step, host enqueue %.3f ms/step, device %.3f ms/step; host parts %s" % (
    1e3 * tot / steps, 1e3 * t_enq / steps, e0.elapsed_time(e1) / steps, {k: round(1e3 * v / steps, 3) for k, v in T.items()}))
# copies alone
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(steps):
    st = pf.stage(host[s % 4]); pf.take(st); pf.release(st)
torch.cuda.synchronize()
print("copies alone: %.3f ms/batch" % (1e3 * (time.perf_counter() - t0) / steps))
from torch.profiler import profile, ProfilerActivity
staged = pf.stage(host[0])
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for s in range(6):
        cur = staged; x = pf.take(cur); staged = pf.stage(host[(s + 1) % 4])
        trainer.train_step(model, x, opt); pf.release(cur)
    torch.cuda.synchronize()
evs = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA])
cp = [(a, b, n) for a, b, n in evs if n.startswith("Memcpy HtoD")]
ks = [(a, b, n) for a, b, n in evs if not n.startswith("Memcpy") and not n.startswith("Memset")]
print("HtoD copies: %d, total %.1f us, longest %.1f us" % (len(cp), sum(b - a for a, b, _ in cp), max(b - a for a, b, _ in cp)))
t0 = ks[0][0]
# print timeline of copies relative to kernels: for each copy, which kernel overlaps
for a, b, n in cp[12:36]:
    ov = [kn for ka, kb, kn in ks if ka < b and kb > a]
    print("copy %8.1f..%8.1f (%.1f us) overlaps %s" % (a - t0, b - t0, b - a, (ov[0][:40] if ov else "NOTHING")))

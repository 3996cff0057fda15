"""Per-kernel device time of the bf16 training step (bench config), measured in situ with CUPTI activity records
(torch.profiler): warm caches, real overlap -- complements the serialised cold-cache ncu launch list.

    python scripts/step_kernels.py [steps] > gpurun_out/step_kernels.txt
    CONFIG=2|3|5 (BASELINE configuration, default 2)   BATCH=tokens|ids|dedup (batch contract, default tokens)
"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import news_recommendation_mind_b200 as mr
from news_recommendation_mind_b200 import data, trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
CFG = bench.CONFIGS[int(os.environ.get("CONFIG", "2"))]
dev = "cuda:0"
torch.manual_seed(42)
model = bench.build_model(CFG, dev, os.environ.get("MINDREC_PRECISION", "bf16"))
opt = trainer.FusedAdam(model, lr=1e-4, bert_lr=6e-6)
ids, mask = data.make_news_table(CFG["n_news"], CFG["L"])
mode = os.environ.get("BATCH", "tokens")
kw = {} if mode == "tokens" else ({"id_only": True} if mode == "ids" else {"id_only": True, "dedup_capacity": (CFG["B"] * (CFG["C"] + CFG["S"]) * 7 // 16 + 255) // 256 * 256})
if mode != "tokens":
    model.attach_news_tokens(ids, mask)
devb = [{k: v.to(dev) for k, v in data.make_train_batch(ids, mask, CFG["B"], CFG["C"], CFG["S"], seed=i, n_users=CFG["n_users"], **kw).items()} for i in range(4)]
print("config", os.environ.get("CONFIG", "2"), "batch", mode, {k: tuple(v.shape) for k, v in devb[0].items()})
for s in range(5):
    trainer.train_step(model, devb[s % 4], opt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record()
t0 = time.perf_counter()
for s in range(steps):
    trainer.train_step(model, devb[s % 4], opt)
t_enq = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print("unprofiled: %.3f ms/step (host enqueue %.3f ms/step)" % (e0.elapsed_time(e1) / steps, 1e3 * t_enq / steps))
# host cost alone: the same loop with the GPU idle at the start of every step
t_host = 0.0
for s in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    trainer.train_step(model, devb[s % 4], opt)
    t_host += time.perf_counter() - t0
torch.cuda.synchronize()
print("host time to enqueue one step with an idle GPU: %.3f ms" % (1e3 * t_host / steps))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for s in range(steps):
        trainer.train_step(model, devb[s % 4], opt)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
seq = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        seq.append((ev.time_range.start, name, ev.device_time if hasattr(ev, "device_time") else ev.cuda_time))
tot = sum(v[1] for v in agg.values())
print("%-64s %5s %10s %7s %9s" % ("kernel", "n/step", "us/step", "share", "us/launch"))
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-64s %5.1f %10.1f %6.1f%% %9.1f" % (name[:64], n / steps, us / steps, 100 * us / tot, us / n))
print("sum of kernel time per step: %.1f us, launches per step: %.1f" % (tot / steps, sum(v[0] for v in agg.values()) / steps))
if os.environ.get("SEQ"):
    seq.sort()
    per = len(seq) // steps
    print("--- launch sequence of the last step ---")
    for t, name, us in seq[-per:]:
        print("%8.1f  %s" % (us, name[:90]))

"""torchrun --nproc-per-node 2 scripts/ddp_probe.py : device timeline of one DDP training step (NCCL kernels, gaps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
import news_recommendation_mind_b200 as mr
from news_recommendation_mind_b200 import data, trainer
CFG = bench.CFG
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = "cuda:%d" % local
dist.init_process_group("nccl", device_id=torch.device(dev))
torch.manual_seed(42)
man = bench.manager_ns(dev, "bf16")
core = mr.TwoTower(man, mr.BERT_Embedding(man, vocab_size=CFG["V"]), mr.CNN_Encoder(man), mr.RNN_User_Encoder(man)).to(dev)
kw = {}
if os.environ.get("BUCKET_VIEW") == "1":
    kw["gradient_as_bucket_view"] = True
model = torch.nn.parallel.DistributedDataParallel(core, device_ids=[local], output_device=local, find_unused_parameters=False, **kw)
opt = trainer.FusedAdam(model, lr=1e-4, bert_lr=6e-6)
ids, mask = data.make_news_table(CFG["n_news"], CFG["L"])
devb = [{k: v.to(dev) for k, v in data.make_train_batch(ids, mask, CFG["B"], CFG["C"], CFG["S"], seed=10 * rank + i).items()} for i in range(4)]
for s in range(6):
    trainer.train_step(model, devb[s % 4], opt)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(30):
    trainer.train_step(model, devb[s % 4], opt)
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print("ddp: %.3f ms/step" % (e0.elapsed_time(e1) / 30))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for s in range(4):
        trainer.train_step(model, devb[s % 4], opt)
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA])
    nccl = [(a, b, n) for a, b, n in evs if "nccl" in n.lower()]
    print("nccl kernels per step: %.1f, us per step: %.1f" % (len(nccl) / 4, sum(b - a for a, b, _ in nccl) / 4))
    for a, b, n in nccl[-4:]:
        print("  %8.1f us  %s" % (b - a, n[:80]))
    # what runs after the last backward kernel of a step: list the last 14 events of the last step
    for a, b, n in evs[-16:]:
        print("%10.1f %8.1f %s" % (a - evs[-16][0], b - a, n[:70]))
dist.destroy_process_group()

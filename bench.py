#!/usr/bin/env python
"""Benchmark of the TwoTower hot path on B200 (BASELINE.json metric: training impressions/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--impl reference]

One "step" = one Manager._train iteration (utils/Manager.py:636-647: zero_grad, forward, NLLLoss,
backward, Adam step) over one synthetic MIND-small-shaped batch of 256 impressions per GPU
(TwoTower CNN news encoder + LSTM user encoder, title 32, history 50, npratio 4, 300d -> 150).
Prints ONE JSON line (rank 0).  See the module-level contract in the task statement for the keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(B=256, C=5, S=50, L=32, E=300, H=150, V=30522, n_news=51282, encoderN="cnn", encoderU="lstm")
FLOP_PER_TOKEN_FWD = 2 * 3 * CFG["E"] * CFG["H"] + 2 * CFG["H"] * CFG["H"] + 4 * CFG["H"]        # 315,600
FLOP_PER_TOKEN_CONV = 2 * 3 * CFG["E"] * CFG["H"]                                                 # 270,000: the conv GEMM alone


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        mhz = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = next((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), None)
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


def manager_ns(device, precision):
    import types
    m = types.SimpleNamespace(scale="small", mode="train", cdd_size=CFG["C"], impr_size=2000, batch_size_news=500,
                              his_size=CFG["S"], signal_length=CFG["L"], device=device, bert_dim=CFG["E"],
                              hidden_dim=CFG["H"], head_num=10, dropout_p=0.2, descend_history=False,
                              encoderN=CFG["encoderN"], encoderU=CFG["encoderU"], precision=precision)
    m.get_user_num = lambda: 94057
    return m


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path, timed on the box's host cores.  The reference
    is pure Python and cannot travel to the GPU box, so this arm runs the oracle port of it
    (oracle/twotower_oracle.py, pinned to the reference by tests/golden) with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import twotower_oracle as O
    from news_recommendation_mind_b200 import data
    torch.manual_seed(42)
    torch.set_num_threads(os.cpu_count() or 1)
    Bs = 64                                           # bounded sample of the 256-impression step
    params = O.init_params(CFG["V"], CFG["E"], CFG["H"], "lstm", seed=42)
    ids, mask = data.make_news_table(CFG["n_news"], CFG["L"])
    state = {}
    t_steps = []
    for s in range(args.warmup + args.steps):
        x = data.make_train_batch(ids, mask, Bs, CFG["C"], CFG["S"], seed=1000 + s)
        t0 = time.perf_counter()
        O.train_step(params, state, x, s + 1, lr=1e-4, bert_lr=6e-6, encoder_n="cnn", encoder_u="lstm")
        if s >= args.warmup:
            t_steps.append(time.perf_counter() - t0)
    total = sum(t_steps)
    val = Bs * len(t_steps) / total
    line = {"impl": "reference", "metric": "train_impressions_per_sec", "value": val, "unit": "impressions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(t_steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, "fp32"),
            "cpu_baseline": {"value": val, "unit": "impressions/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d-impression slice of the 256-impression step, %d steps" % (Bs, len(t_steps))},
            "e2e": {"value": val, "unit": "impressions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(args, precision):
    return {"workload": "TwoTower CNN+LSTM training, batch 256/GPU, title 32, his 50, npratio 4, 300d->150 "
                        "(BASELINE configs[1])", "global_batch": CFG["B"] * args.gpus, "per_gpu_batch": CFG["B"],
            "precision": precision, "parallelism": "dp%d" % args.gpus, "grad_sync": ("torch DDP" if getattr(args, "ddp", False) else "trainer.GradSync (in-place NCCL all-reduce)") if args.gpus > 1 else "none", "optimizer": "Adam lr 1e-4 / bert_lr 6e-6",
            "l2": "8 distinct batches cycled; per-step working set (saved activations ~0.5-1 GB) exceeds the 126 MB L2"}


def cpu_baseline_sample():
    from oracle import twotower_oracle as O
    from news_recommendation_mind_b200 import data
    torch.set_num_threads(os.cpu_count() or 1)
    Bs = 32
    params = O.init_params(CFG["V"], CFG["E"], CFG["H"], "lstm", seed=42)
    ids, mask = data.make_news_table(CFG["n_news"], CFG["L"])
    state, ts = {}, []
    n = 0
    t_all = time.perf_counter()
    for s in range(40):
        x = data.make_train_batch(ids, mask, Bs, CFG["C"], CFG["S"], seed=2000 + s)
        t0 = time.perf_counter()
        O.train_step(params, state, x, s + 1, lr=1e-4, bert_lr=6e-6, encoder_n="cnn", encoder_u="lstm")
        if s >= 1:
            ts.append(time.perf_counter() - t0)
            n += Bs
        if time.perf_counter() - t_all > 15 and len(ts) >= 3:
            break
    return {"value": n / sum(ts), "unit": "impressions/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle fp32 train step (fwd+NLL+bwd+Adam) on %d-impression slices, %d steps after 1 warm-up" % (Bs, len(ts))}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import news_recommendation_mind_b200 as mr
    from news_recommendation_mind_b200 import _lib, build, data, ops, trainer
    build.build()
    lib = _lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    if lib.mr_device_check(local) != 0:
        raise RuntimeError(lib.mr_last_error().decode())
    torch.manual_seed(42)
    man = manager_ns(dev, args.precision)
    model = mr.TwoTower(man, mr.BERT_Embedding(man, vocab_size=CFG["V"]), mr.CNN_Encoder(man), mr.RNN_User_Encoder(man)).to(dev)
    core = model
    sync = None
    if world > 1 and args.ddp:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], output_device=local,
                                                          find_unused_parameters=False)
    opt = trainer.FusedAdam(model, lr=1e-4, bert_lr=6e-6)
    if world > 1 and not args.ddp:
        sync = trainer.GradSync(model, opt)           # in-place NCCL all-reduce, table gradient overlapped with the backward
    ids, mask = data.make_news_table(CFG["n_news"], CFG["L"])
    NB = 8
    host = [data.make_train_batch(ids, mask, CFG["B"], CFG["C"], CFG["S"], seed=100 * rank + i, pin=True) for i in range(NB)]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values() if torch.is_tensor(v))      # every field of the batch dict

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, steps, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if read_loss:
            # end to end through the public training loop: pinned host batches, every step copies its inputs to the device
            # (next batch staged on a side stream) and its loss is read back on the host (one step of lag)
            losses = loop.run(batches, steps)
            assert len(losses) == steps
        else:
            for s in range(steps):
                trainer.train_step(model, batches[s % NB], opt, sync)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        if read_loss:
            ms = max(ms, wall * 1e3)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    warm = max(args.warmup, 3) if world == 1 else max(args.warmup, 12)      # multi-GPU: NCCL channels / algorithm tuning settle later
    for s in range(warm):
        trainer.train_step(model, devb[s % NB], opt, sync)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    import ctypes
    lib.mr_debug_conv_timing.argtypes = [ctypes.c_int]
    lib.mr_debug_conv_timing_read.argtypes = [ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_float)]
    timing_on = args.precision == "bf16" and lib.mr_debug_conv_timing(1) == 0
    l0 = lib.mr_launch_count()
    ms = timed(devb, args.steps, False)
    launches = lib.mr_launch_count() - l0
    conv_launches, conv_ms = 0, 0.0
    if timing_on:
        n_, m_ = ctypes.c_int64(0), ctypes.c_float(0)
        if lib.mr_debug_conv_timing_read(ctypes.byref(n_), ctypes.byref(m_)) == 0:
            conv_launches, conv_ms = n_.value, m_.value
        lib.mr_debug_conv_timing(0)
    # single GPU, extra (not the headline): the same step replayed as ONE CUDA graph (trainer.GraphStep: identical device work,
    # bit-identical results, tests/test_gpu_tc.py::test_graph_step_matches_eager_steps; ~0.07 ms of host work per step instead
    # of 1.2-1.5 ms).  `value` / `e2e` stay on the eager path, which is also what the multi-GPU runs execute (GraphStep cannot
    # capture the NCCL calls yet), so that the per-N numbers compare like with like.
    graph_info = None
    if world == 1 and not args.no_graph:
        try:
            gstep = trainer.GraphStep(model, opt, devb[0])
            for s in range(3):
                gstep(devb[s % NB])
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for s in range(args.steps):
                gstep(devb[s % NB])
            g1.record()
            torch.cuda.synchronize()
            ms_graph = float(g0.elapsed_time(g1))
            graph_info = {"ms_per_step": ms_graph / args.steps, "impressions_per_sec": CFG["B"] * args.steps / (ms_graph * 1e-3),
                          "ms_per_step_eager": ms / args.steps,
                          "note": "trainer.GraphStep: the whole step as one CUDA graph replay, device-resident batches"}
            del gstep
        except Exception as exc:                                   # noqa: BLE001 -- an extra; the eager numbers stand
            graph_info = {"error": repr(exc)[:300]}
        opt.dyn = None                                             # back to host-side Adam bias corrections for the eager steps below
    clocks = sampler.stop() if sampler else None
    loop = trainer.TrainLoop(model, opt, sync)    # the public training loop (staging buffers / pinned loss slots made once)
    loop.run(host, 3)
    ms_e2e = timed(host, args.steps, True)

    # dominant kernel = the conv-forward tap GEMM (gather + 3-tap implicit GEMM + bias + ReLU on tcgen05): its
    # launches inside the timed region above were bracketed by CUDA events on the launching stream
    roof = None
    if rank == 0 and conv_launches > 0:
        pk = peaks()
        n_titles = CFG["B"] * (CFG["C"] + CFG["S"])
        flops = FLOP_PER_TOKEN_CONV * n_titles * CFG["L"]
        t_k = conv_ms * 1e-3
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("conv_fwd_tapgemm_dram_bytes_per_launch")
        roof = {"kernel": "tapgemm_kernel conv forward (token gather + 3-tap implicit GEMM + bias + ReLU), %d titles x %d tokens"
                          % (n_titles, CFG["L"]),
                "bound": "tensor", "achieved": flops / t_k / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": flops / t_k / 1e12 / pk["tf_sustained"], "traffic": traffic,
                "peak_source": pk["source"] + " sustained (kernel timed inside the training step)",
                "us_per_launch": t_k * 1e6, "launches_timed": int(conv_launches),
                "algorithmic_flop_per_token": FLOP_PER_TOKEN_CONV}
    # extra (not the headline): the same training step with in-batch unique-news dedup switched on
    dedup_info = None
    if args.precision == "bf16":
        core.dedup_titles = True
        for s_ in range(3):
            trainer.train_step(model, devb[s_ % NB], opt, sync)
        ms_d = timed(devb, args.steps, False)
        core.dedup_titles = False
        dedup_info = {"value": world * CFG["B"] * args.steps / (ms_d * 1e-3), "unit": "impressions/s", "ms_per_step": ms_d / args.steps,
                      "unique_titles_last_step": getattr(core, "last_unique_titles", None), "titles_per_step": CFG["B"] * (CFG["C"] + CFG["S"]),
                      "note": "identical outputs; every distinct news of the batch is encoded once (history padding = news 0)"}
    # second half of the BASELINE metric: evaluation news-encoded/s (Manager._eval_fast hot loop 1: the whole news set
    # through encode_news, sharded over ranks + all-gather), timed with CUDA events around the whole table build
    from news_recommendation_mind_b200 import evaluate as ev
    ids, mask = ids.pin_memory(), mask.pin_memory()          # the host copy of the token table, pinned once
    ids_d, mask_d = ids.to(dev), mask.to(dev)

    def time_eval(i_, m_):
        with torch.no_grad():
            ev.encode_all_news(core, i_, m_)                 # warm-up (allocator, first-use initialisation)
            best = None
            for _ in range(3):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ev.encode_all_news(core, i_, m_)
                e1.record()
                torch.cuda.synchronize()
                t_ = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                best = float(t_) if best is None else min(best, float(t_))
        return best
    t_eval = time_eval(ids_d, mask_d)                        # token table resident in HBM (inputs pre-staged)
    t_eval_h = time_eval(ids, mask)                          # from the pinned host table, H2D inside the timed region
    eval_info = {"metric": "eval_news_encoded_per_sec", "value": ids.shape[0] / (t_eval * 1e-3), "unit": "news/s",
                 "news": int(ids.shape[0]), "ms": t_eval,
                 "e2e": {"value": ids.shape[0] / (t_eval_h * 1e-3), "unit": "news/s", "ms": t_eval_h,
                         "h2d_bytes": int(ids.numel() * 8 * 2 // world)},
                 "note": "full small-train news set (51,283 titles): sharded over ranks, encoded, NCCL all-gather of the [N+1,H] "
                         "table; value = token table resident in HBM, e2e = from the pinned host table; best of 3, max over ranks"}
    if world > 1:
        dist.barrier()
    if rank == 0:
        per_step = ms / args.steps
        line = {"metric": "train_impressions_per_sec", "value": world * CFG["B"] * args.steps / (ms * 1e-3),
                "unit": "impressions/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": config_dict(args, args.precision), "clocks": clocks,
                "e2e": {"value": world * CFG["B"] * args.steps / (ms_e2e * 1e-3), "unit": "impressions/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches), "roofline": roof, "eval": eval_info, "dedup": dedup_info, "cuda_graph": graph_info}
        line["config"]["step_execution"] = "eager launches (ctypes -> libmindrec.so) with programmatic dependent launch; see cuda_graph for the graph replay"
        if world == 1:
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="N = 1: time the eager step only (no trainer.GraphStep)")
    ap.add_argument("--ddp", action="store_true", help="N > 1: wrap the model in torch DDP (the reference's scheme) instead of trainer.GradSync")
    ap.add_argument("--precision", default=os.environ.get("MINDREC_PRECISION", "bf16"), choices=["bf16", "fp32"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

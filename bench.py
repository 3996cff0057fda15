#!/usr/bin/env python
"""Benchmark of the TwoTower hot path on B200 (BASELINE.json metric: training impressions/s, eval news-encoded/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--precision bf16|fp32] [--impl reference]

--config selects the BASELINE.json configuration (the default, 2, is the one the headline metric is quoted on):
  2  TwoTower CNN + LSTM training, batch 256 per GPU, bf16, MIND-small shape (title 32, his 50, npratio 4, 300d -> 150)
  3  TwoTower CNN + MHA user encoder, MIND-large tables (101,527 news, 876,956 users), data-parallel training
  4  full-news-set evaluation: sharded news encoding + NCCL all-gather + impression scoring + metrics (large dev / test sets)
  5  TwoTower MHA news encoder + LSTUR user encoder, title 48, his 100, npratio 9, MIND-large tables
One training "step" = one Manager._train iteration (utils/Manager.py:636-647: zero_grad, forward, NLLLoss, backward,
gradient all-reduce, Adam step) over one synthetic MIND-shaped batch of 256 impressions per GPU.
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

SMALL = dict(n_news=51282, n_users=94057, scale="small")
LARGE = dict(n_news=101527, n_users=876956, scale="large")
CONFIGS = {
    2: dict(B=256, C=5, S=50, L=32, E=300, H=150, V=30522, hn=10, encoderN="cnn", encoderU="lstm", **SMALL,
            workload="TwoTower CNN+LSTM training, batch 256/GPU, title 32, his 50, npratio 4, 300d->150 (BASELINE configs[1])"),
    3: dict(B=256, C=5, S=50, L=32, E=300, H=150, V=30522, hn=10, encoderN="cnn", encoderU="mha", **LARGE,
            workload="TwoTower CNN + MHA user encoder (10 heads) training, batch 256/GPU, MIND-large tables (101,527 news, "
                     "876,956 users), title 32, his 50, npratio 4 (BASELINE configs[2])"),
    4: dict(B=256, C=5, S=50, L=32, E=300, H=150, V=30522, hn=10, encoderN="cnn", encoderU="lstm", **LARGE,
            news_sets={"large_dev": 72023, "large_test": 120961}, n_impr=376000,
            workload="full-news-set evaluation: sharded news encoding + NCCL all-gather + scoring of 376,000 impressions + "
                     "AUC/MRR/nDCG (BASELINE configs[3])"),
    5: dict(B=256, C=10, S=100, L=48, E=300, H=150, V=30522, hn=10, encoderN="mha", encoderU="lstur", **LARGE,
            workload="TwoTower MHA news encoder (10 heads) + LSTUR user encoder training, batch 256/GPU, title 48, his 100, "
                     "npratio 9, MIND-large tables (BASELINE configs[4])"),
}
CFG = CONFIGS[2]            # the headline configuration (tests / scripts import this)


def flop_per_impression(cfg):
    """algorithmic training FLOP per impression (forward x 3, padding and recompute not counted)"""
    E, H, L, C, S, hn = cfg["E"], cfg["H"], cfg["L"], cfg["C"], cfg["S"], cfg["hn"]
    if cfg["encoderN"] == "cnn":
        per_news = L * (2 * 3 * E * H + 2 * H * H + 4 * H)
    else:
        per_news = 2 * L * E * E + 2 * L * E * H + hn * 2 * L * L * (E // hn + H // hn) + 4 * L * H
    if cfg["encoderU"] in ("lstm", "lstur"):
        user = S * 2 * (2 * H * 4 * H)
    elif cfg["encoderU"] == "gru":
        user = S * 2 * (2 * H * 3 * H)
    elif cfg["encoderU"] == "mha":
        user = 2 * S * H * H * 2 + hn * 2 * S * S * (2 * H // hn) + 4 * S * H
    else:
        user = 4 * S * H
    return 3 * ((C + S) * per_news + user + 2 * C * H)


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


def manager_ns(device, precision, cfg=None):
    import types
    cfg = cfg or CFG
    m = types.SimpleNamespace(scale=cfg["scale"], mode="train", cdd_size=cfg["C"], impr_size=2000, batch_size_news=500,
                              his_size=cfg["S"], signal_length=cfg["L"], device=device, bert_dim=cfg["E"],
                              hidden_dim=cfg["H"], head_num=cfg["hn"], dropout_p=0.2, descend_history=False,
                              encoderN=cfg["encoderN"], encoderU=cfg["encoderU"], precision=precision)
    m.get_user_num = lambda: cfg["n_users"]
    return m


def build_model(cfg, dev, precision):
    import news_recommendation_mind_b200 as mr
    man = manager_ns(dev, precision, cfg)
    encN = {"cnn": mr.CNN_Encoder, "mha": mr.MHA_Encoder}[cfg["encoderN"]](man)
    encU = {"lstm": mr.RNN_User_Encoder, "gru": mr.RNN_User_Encoder, "attn": mr.Attention_Pooling, "avg": mr.Average_Pooling,
            "mha": mr.MHA_User_Encoder, "lstur": mr.LSTUR_User_Encoder}[cfg["encoderU"]](man)
    return mr.TwoTower(man, mr.BERT_Embedding(man, vocab_size=cfg["V"]), encN, encU).to(dev)


def config_dict(args, cfg, extra=None):
    """the WORKLOAD of the line -- identical in both arms (`--impl reference` times the reference on this arm's config); what is
    specific to an implementation (precision, gradient exchange, how the step is launched) goes under `implementation`"""
    d = {"workload": cfg["workload"], "baseline_config": args.config, "global_batch": cfg["B"] * args.gpus, "per_gpu_batch": cfg["B"],
         "parallelism": "dp%d" % args.gpus, "optimizer": "Adam lr 1e-4 / bert_lr 6e-6",
         "l2": "8 distinct batches cycled; per-step working set (saved activations ~0.5-1 GB) exceeds the 126 MB L2"}
    if extra:
        d.update(extra)
    return d


def eval_config(args, cfg):
    return {"workload": cfg["workload"], "baseline_config": args.config, "parallelism": "dp%d" % args.gpus, "news_sets": cfg["news_sets"],
            "impressions": cfg["n_impr"],
            "l2": "every pass re-encodes the whole news set in 32k-title chunks; a chunk's activations (~0.6 GB) exceed the 126 MB L2"}


def implementation_dict(args, precision):
    return {"precision": precision,
            "grad_sync": ("torch DDP" if getattr(args, "ddp", False) else "trainer.GradSync (in-place NCCL all-reduce, one NCCL stream)") if args.gpus > 1 else "none"}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path, on the box's host cores
# ------------------------------------------------------------------------------------------------
def _reference_available():
    try:
        from oracle import ref_harness as RH
        return RH.reference_root() is not None
    except Exception:                                 # noqa: BLE001
        return False


def cpu_train_baseline(cfg, warmup, steps, Bs=None, budget_s=None):
    """Times the reference (oracle/_ref: the unmodified models.TwoTower + torch.optim.Adam through the Manager._train loop) --
    or, where it was not staged, the oracle port -- on every host core.  -> (impressions/s, ms/step, cores, kind, sample)"""
    from news_recommendation_mind_b200 import data
    torch.manual_seed(42)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    Bs = Bs or cfg["B"]
    ids, mask = data.make_news_table(cfg["n_news"], cfg["L"])
    batches = [data.make_train_batch(ids, mask, Bs, cfg["C"], cfg["S"], seed=1000 + s, n_users=cfg["n_users"]) for s in range(4)]
    kind = "reference" if _reference_available() else "port"
    ts = []
    t_all = time.perf_counter()
    if kind == "reference":
        from oracle import ref_harness as RH
        model = RH.build_model(cfg["encoderN"], cfg["encoderU"], V=cfg["V"], E=cfg["E"], H=cfg["H"], C=cfg["C"], S=cfg["S"],
                               L=cfg["L"], hn=cfg["hn"], n_users=cfg["n_users"])
        model.train()
        if cfg["encoderU"] == "lstur":
            model.encoderU.keep_user = torch.ones(Bs, dtype=torch.long)
        opt = RH.make_optimizer(model)
        step_fn = lambda x: RH.train_step(model, opt, x)                                      # noqa: E731
    else:
        from oracle import twotower_oracle as O
        params = O.init_params(cfg["V"], cfg["E"], cfg["H"], cfg["encoderU"], seed=42)
        state, cnt = {}, [0]

        def step_fn(x):
            cnt[0] += 1
            O.train_step(params, state, x, cnt[0], lr=1e-4, bert_lr=6e-6, encoder_n=cfg["encoderN"], encoder_u=cfg["encoderU"])
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        step_fn(batches[s % len(batches)])
        if s >= warmup:
            ts.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_all > budget_s and len(ts) >= 2:
            break
    total = sum(ts)
    sample = "%s %d-impression step (zero_grad, forward, NLLLoss, backward, Adam), %d steps after %d warm-up" % (
        "full" if Bs == cfg["B"] else "%d-impression slice of the" % Bs, cfg["B"], len(ts), warmup)
    return Bs * len(ts) / total, 1e3 * total / len(ts), torch.get_num_threads(), kind, sample


def cpu_eval_baseline(cfg, batches=12):
    """Manager._eval_fast hot loop 1 on the host: encode_news over DataLoader batches of 500 titles (Manager.py:498-499)."""
    from news_recommendation_mind_b200 import data
    torch.set_num_threads(os.cpu_count() or 1)
    ids, mask = data.make_news_table(batches * 500, cfg["L"])
    kind = "reference" if _reference_available() else "port"
    if kind == "reference":
        from oracle import ref_harness as RH
        model = RH.build_model(cfg["encoderN"], cfg["encoderU"], V=cfg["V"], E=cfg["E"], H=cfg["H"], C=cfg["C"], S=cfg["S"], L=cfg["L"],
                               hn=cfg["hn"], n_users=40)
        model.eval()
        enc = lambda i, m: model.encode_news({"cdd_encoded_index": i.unsqueeze(1), "cdd_attn_mask": m.unsqueeze(1)})   # noqa: E731
    else:
        from oracle import twotower_oracle as O
        params = O.init_params(cfg["V"], cfg["E"], cfg["H"], cfg["encoderU"], seed=42)
        enc = lambda i, m: O.encode_news(params, i.unsqueeze(1), m.unsqueeze(1), cfg["encoderN"])                       # noqa: E731
    ts = []
    with torch.no_grad():
        for b in range(batches):
            t0 = time.perf_counter()
            enc(ids[1 + b * 500: 1 + (b + 1) * 500], mask[1 + b * 500: 1 + (b + 1) * 500])
            if b >= 2:
                ts.append(time.perf_counter() - t0)
    return 500 * len(ts) / sum(ts), torch.get_num_threads(), kind, "encode_news over %d batches of 500 titles after 2 warm-up" % len(ts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    if not _reference_available() and (cfg["encoderN"] != "cnn" or cfg["encoderU"] not in ("lstm", "gru")):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not staged (run oracle/build_ref.py where /root/reference "
                          "exists) and the oracle port's trainer covers CNN + LSTM/GRU only"}))
        return
    if args.config == 4:
        val, cores, kind, sample = cpu_eval_baseline(cfg, batches=2 + max(args.steps, 4))
        line = {"impl": "reference", "metric": "eval_news_encoded_per_sec", "value": val, "unit": "news/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * 500 / val, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": eval_config(args, cfg),
                "implementation": {"precision": "fp32", "what": "the reference's encode_news on the host cores (oracle/_ref)"},
                "cpu_baseline": {"value": val, "unit": "news/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": val, "unit": "news/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    Bs = cfg["B"] if args.config in (2, 3) else 64            # config 5 is ~3x the work per impression: bounded slice
    # the steps asked for, but never more than ~3 minutes of host time (a slow host must still end with a line: the loop stops
    # early once the budget is spent and `sample` says how many steps were timed)
    val, ms, cores, kind, sample = cpu_train_baseline(cfg, args.warmup, args.steps, Bs=Bs, budget_s=170)
    line = {"impl": "reference", "metric": "train_impressions_per_sec", "value": val, "unit": "impressions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, cfg),
            "implementation": {"precision": "fp32", "what": "the reference's models.TwoTower + torch.optim.Adam through the Manager._train loop "
                                                            "on the host cores (oracle/_ref)" if kind == "reference" else "oracle port on the host cores"},
            "cpu_baseline": {"value": val, "unit": "impressions/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "impressions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
class Env:
    """process-group / device set-up shared by the training and evaluation benchmarks"""

    def __init__(self):
        import torch.distributed as dist
        from news_recommendation_mind_b200 import _lib, build
        build.build()
        self.lib = _lib.load()
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = "cuda:%d" % self.local
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
        if self.lib.mr_device_check(self.local) != 0:
            raise RuntimeError(self.lib.mr_last_error().decode())
        torch.manual_seed(42)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def close(self, graphs=()):
        """orderly teardown: captured graphs first (a live graph with captured NCCL kernels makes the communicator teardown
        hang), then the process group; a watchdog ends the process if the teardown itself stalls -- the JSON line is out"""
        sys.stdout.flush()
        for g in graphs:
            if g is not None:
                g.close()
        torch.cuda.synchronize()
        if self.world > 1:
            t = threading.Timer(30.0, lambda: os._exit(0))
            t.daemon = True
            t.start()
            self.dist.barrier()
            self.dist.destroy_process_group()
            t.cancel()


def run_train(args):
    from news_recommendation_mind_b200 import data, trainer
    import ctypes
    cfg = CONFIGS[args.config]
    env = Env()
    dist, world, rank, dev, lib = env.dist, env.world, env.rank, env.dev, env.lib
    # ---------------- set-up (not timed, not warm-up): model, optimiser, NCCL channels, token table, batches, graph capture
    model = build_model(cfg, dev, args.precision)
    core = model
    sync = None
    if world > 1 and args.ddp:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[env.local], output_device=env.local,
                                                          find_unused_parameters=False)
    opt = trainer.FusedAdam(model, lr=1e-4, bert_lr=6e-6)
    if world > 1 and not args.ddp:
        sync = trainer.GradSync(model, opt, prewarm=16)        # broadcasts the weights and runs the step's collectives on scratch buffers
    ids, mask = data.make_news_table(cfg["n_news"], cfg["L"])
    fused = cfg["encoderN"] == "cnn"                           # the id-only input pipeline feeds the fused CNN path
    if fused:
        core.attach_news_tokens(ids, mask)
    NB = 8
    mk = lambda i, **kw: data.make_train_batch(ids, mask, cfg["B"], cfg["C"], cfg["S"], seed=100 * rank + i, pin=True,       # noqa: E731
                                                 n_users=cfg["n_users"], **kw)
    host_tok = [mk(i) for i in range(NB)]                      # the reference's batch contract: int64 token tensors (MIND.py:352-363)
    host = [mk(i, id_only=True) for i in range(NB)] if fused else host_tok
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    devb_tok = [{k: v.to(dev) for k, v in b.items()} for b in host_tok] if fused else devb
    nbytes = lambda b: sum(v.numel() * v.element_size() for v in b.values() if torch.is_tensor(v))                             # noqa: E731
    if cfg["encoderU"] == "lstur":
        core.encoderU.keep_user = None                         # Bernoulli(0.5) drawn per step, as the reference does
    gstep, graph_error = None, None
    if not args.no_graph and not args.ddp:
        try:
            gstep = trainer.GraphStep(model, opt, devb[0], sync)
        except Exception as exc:                               # noqa: BLE001 -- the eager path below is the same device work
            graph_error = repr(exc)[:300]
            opt.dyn = None
            if world > 1:
                raise                                          # a rank-local fallback would desynchronise the collectives

    def step_fn(x):
        return gstep(x) if gstep is not None else trainer.train_step(model, x, opt, sync)

    def timed(fn, steps):
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn(steps)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        env.barrier()
        return env.max_over_ranks(e0.elapsed_time(e1)), env.max_over_ranks(wall)

    def loop_dev(batches, f=None):
        f = f or step_fn
        return lambda steps: [f(batches[s % NB]) for s in range(steps)]

    # ---------------- warm-up: exactly the W steps asked for
    for s in range(args.warmup):
        step_fn(devb[s % NB])
    sampler = ClockSampler(env.local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.mr_launch_count()
    ms, _ = timed(loop_dev(devb), args.steps)                                   # ---- the headline timed region
    launches_counted = lib.mr_launch_count() - l0
    windows = [timed(loop_dev(devb), args.steps)[0] / args.steps for _ in range(4)]
    clocks = sampler.stop() if sampler else None

    # ---------------- the same steps launched eagerly (kernel by kernel through ctypes): launch count, conv-kernel event timing
    lib.mr_debug_conv_timing.argtypes = [ctypes.c_int]
    lib.mr_debug_conv_timing_read.argtypes = [ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_float)]
    opt_dyn = opt.dyn
    opt.dyn = None                                            # eager steps: host-side Adam scalars
    eager = lambda x: trainer.train_step(model, x, opt, sync)                                                                  # noqa: E731
    for s in range(3):
        eager(devb[s % NB])
    timing_on = args.precision == "bf16" and fused and lib.mr_debug_conv_timing(1) == 0
    l0 = lib.mr_launch_count()
    ms_eager, _ = timed(loop_dev(devb, eager), args.steps)
    launches_eager = lib.mr_launch_count() - l0
    conv_launches, conv_ms = 0, 0.0
    ktimes = {}                                                # which -> (launches, mean ms): 0 conv forward, 1 tail forward, 2 tail backward
    if timing_on:
        lib.mr_debug_kernel_timing_read.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_float)]
        for which in range(3):
            n_, m_ = ctypes.c_int64(0), ctypes.c_float(0)
            if lib.mr_debug_kernel_timing_read(which, ctypes.byref(n_), ctypes.byref(m_)) == 0 and n_.value > 0:
                ktimes[which] = (n_.value, m_.value)
        conv_launches, conv_ms = ktimes.get(0, (0, 0.0))
        lib.mr_debug_conv_timing(0)
    tok_info = None
    if fused and not args.quick:
        for s in range(2):
            eager(devb_tok[s % NB])
        ms_tok, _ = timed(loop_dev(devb_tok, eager), args.steps)
        tok_info = {"ms_per_step": ms_tok / args.steps, "value": world * cfg["B"] * args.steps / (ms_tok * 1e-3),
                    "note": "eager steps on device-resident batches in the reference's contract (int64 token tensors, MIND.py:352-363)"}
    opt.dyn = opt_dyn

    # ---------------- end to end through the public training loop: pinned host batches, every step copies its inputs to the device
    # (next batch staged on a side stream) and its loss is read back on the host (one step of lag)
    loop = trainer.TrainLoop(model, opt, sync, graph_step=gstep)
    loop.run(host, 3)

    def loop_e2e(batches):
        def f(steps):
            losses = loop.run(batches, steps)
            assert len(losses) == steps
        return f
    d_ms, w_ms = timed(loop_e2e(host), args.steps)
    ms_e2e = max(d_ms, w_ms)
    e2e_tok = None
    if fused and gstep is None and not args.quick:
        loop.run(host_tok, 2)
        d_ms, w_ms = timed(loop_e2e(host_tok), args.steps)
        e2e_tok = {"value": world * cfg["B"] * args.steps / (max(d_ms, w_ms) * 1e-3), "unit": "impressions/s",
                   "h2d_bytes_per_step": nbytes(host_tok[0]), "note": "batches in the reference's contract (int64 token tensors)"}

    # ---------------- roofline of the dominant kernel (conv forward) from the event-bracketed launches of the eager region
    pk = peaks()
    roof = None
    if rank == 0 and conv_launches > 0:
        n_titles = cfg["B"] * (cfg["C"] + cfg["S"])
        flops = FLOP_PER_TOKEN_CONV * n_titles * cfg["L"]
        t_k = conv_ms * 1e-3
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("conv_fwd_tapgemm_dram_bytes_per_launch")
        roof = {"kernel": "conv forward (token gather + 3-tap implicit GEMM + bias + ReLU on tcgen05), %d titles x %d tokens"
                          % (n_titles, cfg["L"]),
                "bound": "tensor", "achieved": flops / t_k / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": flops / t_k / 1e12 / pk["tf_sustained"], "traffic": traffic,
                "peak_source": pk["source"] + " sustained (kernel timed inside the training step)",
                "us_per_launch": t_k * 1e6, "launches_timed": int(conv_launches),
                "algorithmic_flop_per_token": FLOP_PER_TOKEN_CONV}
    # the other two big kernels of the encoder (HBM bound: the fused projection / pooling tail, csrc/cnn_tail.cu); `roofline` is the
    # line of whichever launch is the longest of the step, `roofline_kernels` lists all three
    roof_all = []
    if roof is not None:
        roof_all.append(roof)
        T_tok = cfg["B"] * (cfg["C"] + cfg["S"]) * cfg["L"]
        Hp = (cfg["H"] + 15) // 16 * 16
        tj = json.load(open(tp)) if os.path.exists(tp) else {}
        for which, name, nbytes_tok, key in (
                (1, "fused tail forward (c tile -> projection MMA -> tanh -> key; softmax; pooled sum as a second MMA)",
                 2 * Hp * 2 + 4, "cnn_tail_fwd_dram_bytes_per_launch"),
                (2, "fused tail backward (pooling backward in place over the key tile, dkp Wq and dkp^T c MMAs, relu' epilogue -> dconv)",
                 3 * Hp * 2 + 32 + 4, "cnn_tail_bwd_dram_bytes_per_launch")):
            if which in ktimes:
                t_k = ktimes[which][1] * 1e-3
                alg = nbytes_tok * T_tok
                roof_all.append({"kernel": name, "bound": "hbm", "achieved": alg / t_k / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                 "frac": alg / t_k / 1e9 / pk["hbm"], "traffic": tj.get(key), "us_per_launch": t_k * 1e6,
                                 "launches_timed": int(ktimes[which][0]), "algorithmic_bytes_per_token": nbytes_tok,
                                 "peak_source": pk["source"] + " copy bandwidth"})
        roof = max(roof_all, key=lambda r: r["us_per_launch"])
    fpi = flop_per_impression(cfg)
    step_roof = {"bound": "tensor", "algorithmic_gflop_per_impression": fpi / 1e9,
                 "achieved": cfg["B"] * fpi / (ms / args.steps * 1e-3) / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s"}
    step_roof["frac"] = step_roof["achieved"] / pk["tf_sustained"]

    # ---------------- extra: in-batch unique-news dedup: the dedup plan (distinct news ids padded to a FIXED capacity +
    # slot -> distinct-news index) is made on the host with the batch (data.dedup_plan), so the step keeps fixed shapes and is
    # replayed as one CUDA graph like the headline step
    dedup_info = None
    if args.precision == "bf16" and fused and not args.ddp and not args.quick:
        n_titles = cfg["B"] * (cfg["C"] + cfg["S"])
        cap = (n_titles * 7 // 16 + 255) // 256 * 256
        while True:
            host_d = [mk(i, id_only=True, dedup_capacity=cap) for i in range(NB)]
            fits = torch.tensor([float(all("uniq_id" in b for b in host_d))], device=dev)
            if world > 1:
                dist.all_reduce(fits, op=dist.ReduceOp.MIN)
            if float(fits) > 0 or cap >= n_titles:
                break
            cap = min(n_titles, cap + 1024)
        devb_d = [{k: v.to(dev) for k, v in b.items()} for b in host_d]
        try:
            gstep_d = None if args.no_graph else trainer.GraphStep(model, opt, devb_d[0], sync)
            if gstep_d is None:
                opt.dyn = None
            f_d = gstep_d if gstep_d is not None else eager
            for s_ in range(3):
                f_d(devb_d[s_ % NB])
            ms_d, _ = timed(loop_dev(devb_d, f_d), args.steps)
            loop_d = trainer.TrainLoop(model, opt, sync, graph_step=gstep_d)
            loop_d.run(host_d, 3)
            d_ms, w_ms = timed(lambda steps: loop_d.run(host_d, steps), args.steps)
            opt.dyn = opt_dyn
            dedup_info = {"value": world * cfg["B"] * args.steps / (ms_d * 1e-3), "unit": "impressions/s", "ms_per_step": ms_d / args.steps,
                          "e2e": {"value": world * cfg["B"] * args.steps / (max(d_ms, w_ms) * 1e-3), "unit": "impressions/s",
                                  "h2d_bytes_per_step": nbytes(host_d[0])},
                          "capacity": cap, "titles_per_step": n_titles,
                          "distinct_titles_in_batch_0": int((host_d[0]["uniq_id"] != 0).sum()) + 1,
                          "step_execution": "CUDA graph" if gstep_d is not None else "eager",
                          "note": "identical log-probabilities (bit for bit), gradients summed per distinct news; every distinct news of the "
                                  "batch is encoded once (history padding = news 0), capacity slots per step"}
        except Exception as exc:                               # noqa: BLE001 -- an extra; the headline numbers stand
            if world > 1:
                raise
            opt.dyn = opt_dyn
            dedup_info = {"error": repr(exc)[:300]}

    # ---------------- second half of the BASELINE metric: evaluation news-encoded/s over the whole news set
    eval_info, eval_large = None, None
    if fused and not args.quick:
        eval_info = time_news_encoding(env, core, ids, mask, "%s-train news set" % cfg["scale"])
        # the same on the news set the sharded encoding is meant for: the MIND-large test set of BASELINE configs[3] (120,961 news,
        # Manager.py:884-914) -- at 8 GPUs the 51k-title set above is all launch and all-gather latency.  (`--config 4` adds scoring.)
        try:
            ids_t, mask_t = data.make_news_table(CONFIGS[4]["news_sets"]["large_test"], cfg["L"], seed=7)
            eval_large = time_news_encoding(env, core, ids_t, mask_t, "large-test news set")
        except Exception as exc:                               # noqa: BLE001 -- an extra; the headline numbers stand
            eval_large = {"error": repr(exc)[:300]}
    if rank == 0:
        per_step = ms / args.steps
        line = {"metric": "train_impressions_per_sec", "value": world * cfg["B"] * args.steps / (ms * 1e-3),
                "unit": "impressions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": config_dict(args, cfg), "implementation": implementation_dict(args, args.precision), "clocks": clocks,
                "e2e": {"value": world * cfg["B"] * args.steps / (ms_e2e * 1e-3), "unit": "impressions/s",
                        "h2d_bytes_per_step": nbytes(host[0]), "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches_eager), "roofline": roof, "roofline_kernels": roof_all, "step_roofline": step_roof,
                "windows": {"ms_per_step": [round(w, 5) for w in windows], "median_ms_per_step": sorted(windows)[len(windows) // 2],
                            "note": "four more timed windows of `steps` steps behind the headline region"},
                "eager": {"ms_per_step": ms_eager / args.steps, "value": world * cfg["B"] * args.steps / (ms_eager * 1e-3),
                          "launches": int(launches_eager), "note": "the same steps launched kernel by kernel (ctypes -> libmindrec.so, PDL)"},
                "token_batches": tok_info, "e2e_token_batches": e2e_tok, "eval": eval_info, "eval_large_test": eval_large, "dedup": dedup_info}
        line["implementation"]["step_execution"] = (
            "one CUDA-graph replay per step (trainer.GraphStep: forward, loss, backward, %sAdam captured once; launches counted by the "
            "library during capture)" % ("NCCL all-reduces, " if world > 1 else "")) if gstep is not None else \
            "eager launches (ctypes -> libmindrec.so) with programmatic dependent launch" + ("; graph capture failed: " + graph_error if graph_error else "")
        line["implementation"]["inputs"] = ("id-only batches (news ids, history mask, user ids, labels); token table resident in HBM, title rows gathered on "
                                    "the device (mr_gather_titles)") if fused else "int64 token batches (the reference's contract)"
        line["step_execution"] = "cuda_graph" if gstep is not None else "eager"          # (details: implementation.step_execution)
        if gstep is not None:
            line["gpu_launches"] = int(launches_eager)            # kernels per `steps` steps: a replay runs the captured launches
            line["graph_replays"] = args.steps
            line["launches_counted_in_timed_region"] = int(launches_counted)
        if world == 1:
            v, ms_c, cores, kind, sample = cpu_train_baseline(cfg, 1, 6, Bs=cfg["B"] if args.config in (2, 3) else 32, budget_s=25)
            line["cpu_baseline"] = {"value": v, "unit": "impressions/s", "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line))
    env.close([gstep, locals().get("gstep_d")])


def time_news_encoding(env, core, ids, mask, what):
    """evaluation news-encoded/s (Manager._eval_fast hot loop 1: the whole news set through encode_news, sharded over ranks +
    all-gather), timed with CUDA events around the whole table build; best of 3, max over ranks"""
    from news_recommendation_mind_b200 import evaluate as ev
    dev, world = env.dev, env.world
    ids_p, mask_p = ids.pin_memory(), mask.pin_memory()          # the host copy of the token table, pinned once
    ids_d, mask_d = ids_p.to(dev), mask_p.to(dev)

    def time_eval(i_, m_):
        with torch.no_grad():
            ev.encode_all_news(core, i_, m_)                 # warm-up (allocator, first-use initialisation)
            best = None
            for _ in range(3):
                env.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ev.encode_all_news(core, i_, m_)
                e1.record()
                torch.cuda.synchronize()
                t_ = env.max_over_ranks(e0.elapsed_time(e1))
                best = t_ if best is None else min(best, t_)
        return best
    t_eval = time_eval(ids_d, mask_d)                        # token table resident in HBM (inputs pre-staged)
    t_eval_h = time_eval(ids_p, mask_p)                      # from the pinned host table, H2D inside the timed region
    n = int(ids.shape[0])
    pk = peaks()
    flop = 2 * 3 * 300 * 150 + 2 * 150 * 150 + 4 * 150
    L = int(ids.shape[1])
    ach = n * L * flop / (t_eval * 1e-3) / 1e12
    return {"metric": "eval_news_encoded_per_sec", "value": n / (t_eval * 1e-3), "unit": "news/s", "news": n, "ms": t_eval,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"] * world, "unit": "TFLOP/s",
                         "frac": ach / (pk["tf_sustained"] * world), "algorithmic_mflop_per_news": L * flop / 1e6},
            "e2e": {"value": n / (t_eval_h * 1e-3), "unit": "news/s", "ms": t_eval_h, "h2d_bytes": int(ids.numel() * 8 * 2 // world)},
            "note": "full %s (%d titles): sharded over ranks, encoded, NCCL all-gather of the [N+1,H] table; value = token table "
                    "resident in HBM, e2e = from the pinned host table; best of 3, max over ranks" % (what, n)}


def run_eval(args):
    """BASELINE config 4: the whole fast evaluation (Manager._eval_fast, Manager.py:473-541) at full scale."""
    from news_recommendation_mind_b200 import data, evaluate as ev, ops
    cfg = CONFIGS[4]
    env = Env()
    world, rank, dev = env.world, env.rank, env.dev
    core = build_model(cfg, dev, args.precision)
    core.eval()
    if world > 1:
        for p in core.parameters():
            env.dist.broadcast(p.detach(), src=0)
        core.embedding.invalidate_shadow()
    sampler = ClockSampler(env.local) if rank == 0 else None
    if sampler:
        sampler.start()
    sets = {}
    table = None
    ids = mask = None
    for name, n_news in cfg["news_sets"].items():
        ids, mask = data.make_news_table(n_news, cfg["L"], seed=7)
        sets[name] = time_news_encoding(env, core, ids, mask, name.replace("_", "-") + " news set")
    with torch.no_grad():
        table = ev.encode_all_news(core, ids.to(dev), mask.to(dev))           # the test-set table: scored against below
    impr = data.make_eval_impressions(ids, mask, cfg["n_impr"], cfg["S"], seed=11, n_users=cfg["n_users"], with_tokens=False)
    n_impr = impr["offsets"].numel() - 1
    n_cand = int(impr["offsets"][-1])
    pin = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in impr.items()}
    res = {}

    def score_pass(impr_):
        return ev.evaluate(core, None, None, impr_, table=table, history="table")
    for label, src in (("device_resident", {k: (v.to(dev) if torch.is_tensor(v) and k != "offsets" else v) for k, v in impr.items()}), ("from_pinned_host", pin)):
        score_pass(src)                                                         # warm-up
        best, metrics = None, None
        for _ in range(max(1, min(args.steps, 3))):
            env.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            metrics = score_pass(src)                                           # ends with the .tolist() of the reduced metric sums
            e1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            t_ = env.max_over_ranks(max(e0.elapsed_time(e1), wall))
            best = t_ if best is None else min(best, t_)
        res[label] = {"ms": best, "impressions_per_sec": n_impr / (best * 1e-3), "candidates_per_sec": n_cand / (best * 1e-3), "metrics": metrics}
    clocks = sampler.stop() if sampler else None
    # the scoring kernel alone (CSR gather + dot + sigmoid), HBM roofline: table rows + user vector in, probabilities out
    H = cfg["H"]
    user = torch.randn(n_impr // world, H, device=dev)
    i0, i1 = ev.partition_bounds(n_impr, world, rank)
    offs = impr["offsets"]
    cdd = impr["cdd_id"][int(offs[i0]):int(offs[i1])].to(dev)
    loff = (offs[i0:i1 + 1] - offs[i0]).to(dev)[: user.shape[0] + 1]
    cdd = cdd[: int(loff[-1])]
    ops.score_sigmoid_gather(table, cdd, loff, user)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ops.score_sigmoid_gather(table, cdd, loff, user)
    e1.record()
    torch.cuda.synchronize()
    t_k = e0.elapsed_time(e1) / 5 * 1e-3
    bytes_alg = cdd.numel() * (H * 4 + 8 + 4) + user.numel() * 4 + loff.numel() * 8
    pk = peaks()
    if rank == 0:
        test = sets["large_test"]
        line = {"metric": "eval_news_encoded_per_sec", "value": test["value"], "unit": "news/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": test["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": eval_config(args, cfg),
                "implementation": dict(implementation_dict(args, args.precision), impressions=n_impr, candidates=n_cand,
                                       history="looked up in the news table (TwoTower.encode_user_from_table)"),
                "clocks": clocks, "e2e": {"value": test["e2e"]["value"], "unit": "news/s", "h2d_bytes_per_step": test["e2e"]["h2d_bytes"],
                                          "d2h_bytes_per_step": 0},
                "gpu_launches": None, "news_encoding": sets,
                "impression_scoring": res,
                "roofline": {"kernel": "score_sigmoid_gather (CSR gather of table rows + dot + sigmoid), this rank's %d candidates" % cdd.numel(),
                             "bound": "hbm", "achieved": bytes_alg / t_k / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": bytes_alg / t_k / 1e9 / pk["hbm"], "traffic": None, "us_per_launch": t_k * 1e6,
                             "algorithmic_bytes_per_candidate": H * 4 + 12}}
        print(json.dumps(line))
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configuration (1-based)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="time eager steps only (no trainer.GraphStep)")
    ap.add_argument("--quick", action="store_true", help="headline + e2e only (no eager / token-batch / dedup / eval extras): stability loops")
    ap.add_argument("--ddp", action="store_true", help="N > 1: wrap the model in torch DDP (the reference's scheme) instead of trainer.GradSync")
    ap.add_argument("--precision", default=os.environ.get("MINDREC_PRECISION", "bf16"), choices=["bf16", "fp32"])
    args = ap.parse_args()
    if os.environ.get("MINDREC_NO_GRAPH"):
        args.no_graph = True
    try:
        if args.impl == "reference":
            run_reference(args)
        elif args.config == 4:
            run_eval(args)
        else:
            run_train(args)
    except BaseException:                                  # noqa: BLE001
        # a failing rank must END: interpreter shutdown with a live NCCL communicator (and captured graphs holding its kernels) can
        # block for the launcher's whole time limit while the other ranks wait in a collective
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            os._exit(1)
        raise


if __name__ == "__main__":
    main()

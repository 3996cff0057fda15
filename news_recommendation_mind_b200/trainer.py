"""Training driver for the hot path: what utils/Manager.py::_train does per step
(Manager.py:636-647 -- zero_grad, forward, NLLLoss, backward, Adam.step) with the optimiser of
Manager._get_optim (Manager.py:389-413: Adam, parameters whose name matches "bert" at bert_lr,
the rest at lr), all arithmetic in libmindrec kernels.  Data-parallel training keeps the
reference's scheme: one process per GPU, torch DDP / NCCL all-reduce (mean) of the gradients
(twotower.py:49-50)."""
from __future__ import annotations

import re

import torch
import torch.distributed as dist

from . import ops
from ._lib import MR_BF16


class FusedAdam:
    """torch.optim.Adam-compatible subset (param_groups, zero_grad, step, state_dict) running
    mr_adam_step; refreshes the bf16 shadow of the token table inside the same kernel."""

    def __init__(self, model, lr=1e-4, bert_lr=6e-6, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        core = model.module if hasattr(model, "module") else model
        base, bert = [], []
        for name, p in core.named_parameters():
            (bert if re.search("bert", name) else base).append(p)
        self.param_groups = [{"params": base, "lr": lr}, {"params": bert, "lr": bert_lr}]
        self.betas, self.eps, self.grad_scale = betas, eps, grad_scale
        self.state = {}
        self.steps = 0
        self.embedding = getattr(core, "embedding", None)
        enc = getattr(core, "encoderN", None)
        self._want_shadow = self.embedding is not None and getattr(enc, "precision", None) == MR_BF16 and \
            hasattr(self.embedding, "shadow_bf16")
        # CUDA-graph mode (set by GraphStep): the step-dependent bias corrections live in device memory, refreshed by
        # begin_step() before every replay, so that the captured launch stays valid for every step
        self.dyn = None
        self._dyn_host = None
        self._dyn_ring, self._dyn_done, self._dyn_slot = None, None, 0
        self._dyn_ids = None             # ids of the parameters the (captured) launch updates, in launch order

    DYN_BLOCK = 4 + 24                   # floats per launch: {1/bc1, 1/sqrt(bc2), grad_scale, -, lr[24]} (mr_adam_step_multi_dyn)
    DYN_SLOTS = 8                        # pinned host copies of the block: how many steps the host may run ahead of the device

    def enable_device_step_scalars(self, device):
        n_params = sum(len(g["params"]) for g in self.param_groups)
        n_blocks = max(1, (n_params + 23) // 24)
        self.dyn = torch.zeros(n_blocks * self.DYN_BLOCK, dtype=torch.float32, device=device)
        # The upload is asynchronous from pinned memory: the DMA reads the host block when the copy EXECUTES, not when it is
        # queued.  A loop that replays steps without synchronising (the host is ~15x faster than a step) would overwrite a single
        # host block long before the copies of the earlier steps run, and those steps would see later steps' bias corrections and
        # learning rates.  Hence a ring of blocks, each guarded by an event recorded right behind its copy: a block is rewritten
        # only after the copy that last read it has run (which also keeps the host at most DYN_SLOTS steps ahead).
        self._dyn_ring = [torch.zeros(n_blocks * self.DYN_BLOCK, dtype=torch.float32).pin_memory() for _ in range(self.DYN_SLOTS)]
        self._dyn_done = [None] * self.DYN_SLOTS
        self._dyn_slot = 0
        self._dyn_host = self._dyn_ring[0]

    def _dyn_items(self):
        """the (parameter, lr) list in group order"""
        return [(p, g["lr"]) for g in self.param_groups for p in g["params"]]

    def _write_dyn(self):
        """upload the step-dependent scalars: bias corrections, grad_scale and the CURRENT learning rate of every tensor of the
        launch (a schedule or load_state_dict may have moved them since the capture); stream ordered, asynchronous"""
        ring, k = self._dyn_ring, self._dyn_slot
        if ring is not None:
            self._dyn_slot = (k + 1) % len(ring)
            if self._dyn_done[k] is not None:
                self._dyn_done[k].synchronize()          # the copy that last read this block has run (normally long ago)
            self._dyn_host = ring[k]
        h = self._dyn_host
        ids = self._dyn_ids
        lrs = [lr for p, lr in self._dyn_items() if (ids is None and p.requires_grad) or (ids is not None and id(p) in ids)]
        for b in range(h.numel() // self.DYN_BLOCK):
            o = b * self.DYN_BLOCK
            h[o] = 1.0 / (1.0 - self.betas[0] ** max(self.steps, 1))
            h[o + 1] = 1.0 / (1.0 - self.betas[1] ** max(self.steps, 1)) ** 0.5
            h[o + 2] = self.grad_scale
            chunk = lrs[b * 24:(b + 1) * 24]
            if chunk:
                h[o + 4:o + 4 + len(chunk)] = torch.tensor(chunk, dtype=torch.float32)
        self.dyn.copy_(h, non_blocking=True)
        if ring is not None:
            ev = self._dyn_done[k] if self._dyn_done[k] is not None else torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dyn.device))
            self._dyn_done[k] = ev

    def begin_step(self):
        """graph mode: advance the step counter and refresh the device block before the (captured) step runs"""
        self.steps += 1
        self._write_dyn()

    # ---- checkpoints: the torch.optim.Adam state-dict layout, so that Manager.save / Manager.load (utils/Manager.py:289-343)
    # work unchanged and a checkpoint written with the reference's optim.Adam resumes here (and vice versa) ---------------
    def state_dict(self):
        state, groups, idx = {}, [], 0
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                st = self.state.get(p)
                if st is not None:
                    state[idx] = {"step": torch.tensor(float(self.steps)), "exp_avg": st[0], "exp_avg_sq": st[1]}
                ids.append(idx)
                idx += 1
            groups.append({"lr": g["lr"], "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                           "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                           "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"]) for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        params = [p for g in self.param_groups for p in g["params"]]
        flat_ids = [i for g in groups for i in g["params"]]
        for g, loaded in zip(self.param_groups, groups):
            g["lr"] = loaded["lr"]
        self.betas, self.eps = tuple(groups[0].get("betas", self.betas)), groups[0].get("eps", self.eps)
        self.state, steps = {}, 0
        for p, i in zip(params, flat_ids):
            st = sd["state"].get(i)
            if st is None:
                continue
            if st["exp_avg"].shape != p.shape:
                raise ValueError("optimizer state of parameter %d has shape %s, expected %s" % (i, tuple(st["exp_avg"].shape), tuple(p.shape)))
            self.state[p] = (st["exp_avg"].to(device=p.device, dtype=torch.float32).clone().contiguous(),
                             st["exp_avg_sq"].to(device=p.device, dtype=torch.float32).clone().contiguous())
            steps = max(steps, int(float(st["step"])))
        self.steps = steps

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self):
        """One Adam update of every parameter that has a gradient: a single mr_adam_step_multi launch (chunks of 24
        tensors), which also rewrites the bf16 shadow of the token table."""
        import ctypes
        from . import _lib
        if self.dyn is None:
            self.steps += 1
        items = []
        for idx, (p, lr) in enumerate(self._dyn_items()):
            if p.grad is None:
                continue
            st = self.state.get(p)
            if st is None:
                st = self.state[p] = (torch.zeros_like(p), torch.zeros_like(p))
            grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            if not p.data.is_contiguous() or grad.dtype != torch.float32 or p.dtype != torch.float32:
                raise RuntimeError("FusedAdam needs contiguous fp32 parameters and gradients")
            items.append((p, grad, st[0], st[1], lr))
        if self.dyn is not None:
            ids = {id(it[0]) for it in items}
            if ids != self._dyn_ids:     # first step in graph mode (eager warm-up): the launch's tensor list is known now
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("FusedAdam: the set of parameters with gradients changed between warm-up and capture")
                self._dyn_ids = ids
                self._write_dyn()
        lib = _lib.load()
        for c0 in range(0, len(items), 24):
            chunk = items[c0:c0 + 24]
            n = len(chunk)
            PA = ctypes.c_void_p * n
            shadow, shadow_idx, row_len, shadow_ld = None, -1, 0, 0
            for i, it in enumerate(chunk):
                if self._want_shadow and it[0] is self.embedding.weight:
                    shadow = self.embedding.shadow_bf16()
                    shadow_idx, row_len, shadow_ld = i, it[0].shape[-1], shadow.shape[-1]
            dev = chunk[0][0].device
            dyn = None if self.dyn is None else self.dyn[(c0 // 24) * self.DYN_BLOCK:(c0 // 24 + 1) * self.DYN_BLOCK]
            _lib.check(lib.mr_adam_step_multi_dyn(
                n, PA(*[it[0].data_ptr() for it in chunk]), PA(*[it[1].data_ptr() for it in chunk]),
                PA(*[it[2].data_ptr() for it in chunk]), PA(*[it[3].data_ptr() for it in chunk]),
                (ctypes.c_int64 * n)(*[it[0].numel() for it in chunk]), (ctypes.c_double * n)(*[float(it[4]) for it in chunk]),
                max(self.steps, 1), self.betas[0], self.betas[1], self.eps, self.grad_scale, shadow_idx, _lib.ptr(shadow), row_len,
                shadow_ld, _lib.ptr(dyn), _lib.stream_ptr(dev)), "mr_adam_step_multi")
            if shadow is not None:
                self.embedding.mark_shadow_fresh(shadow)


class LinearWarmupSchedule:
    """transformers.get_linear_schedule_with_warmup as Manager._get_optim uses it (utils/Manager.py:414-420): every
    group's learning rate = its initial value x (step / warmup while step < warmup, else the linear decay to 0 at
    `total`).  Call step() after every optimiser step."""

    def __init__(self, optimizer, num_warmup_steps, num_training_steps):
        self.optimizer, self.warmup, self.total = optimizer, int(num_warmup_steps), int(num_training_steps)
        self.base = [g["lr"] for g in optimizer.param_groups]
        self.last = 0
        self._apply()

    def factor(self, step):
        if step < self.warmup:
            return float(step) / float(max(1, self.warmup))
        return max(0.0, float(self.total - step) / float(max(1, self.total - self.warmup)))

    def _apply(self):
        f = self.factor(self.last)
        for g, b in zip(self.optimizer.param_groups, self.base):
            g["lr"] = b * f

    def step(self):
        self.last += 1
        self._apply()


class GradSync:
    """Data-parallel gradient averaging for the package's own training loop (the reference wraps the model in DDP,
    twotower.py:49-50, which also works with these modules; this does the same mean with less traffic on the step's
    critical path):
      * the 36.6 MB table gradient is all-reduced (SUM) IN PLACE from the moment the encoder backward has produced it
        -- i.e. while the filter-gradient GEMMs still run -- instead of being copied into and out of a bucket after the
        whole backward;
      * the small dense gradients go through ONE static flat buffer and one all-reduce;
      * the 1/world factor is folded into the Adam kernel (optimizer.grad_scale).
    Every collective is issued with async_op=True: ProcessGroupNCCL then runs all of them on its ONE internal stream, in
    issue order, which is the same on every rank (table gradient(s) first, flat buffer last).  Two collectives of one
    communicator are therefore never in flight on two streams (NCCL gives no ordering guarantee across streams; with one
    of them issued synchronously on the compute stream the two kernels were independent, which is unsafe in eager mode and
    dead-locks as two unordered branches of a captured CUDA graph)."""

    def __init__(self, model, optimizer, group=None, prewarm=8):
        self.model, self.optimizer, self.group = model, optimizer, group
        self.world = dist.get_world_size(group)
        optimizer.grad_scale = 1.0 / self.world
        self.params = [p for p in model.parameters() if p.requires_grad]
        self._events, self._pending, self._extra = {}, [], []
        self._flat, self._flat_key = None, None
        emb = getattr(model, "embedding", None)
        self.table_param = emb.weight if emb is not None and hasattr(emb, "weight") else None
        ops.TABLE_GRAD_HOOK = self
        # same initial weights on every rank, as DDP's constructor does
        src = dist.get_global_rank(group, 0) if group is not None else 0
        with torch.no_grad():
            for p in model.parameters():
                dist.broadcast(p.detach(), src=src, group=group)
        if emb is not None and hasattr(emb, "invalidate_shadow"):
            emb.invalidate_shadow()          # the bf16 shadow may have been built from the pre-broadcast table
        if prewarm:
            self.prewarm(prewarm)

    def prewarm(self, rounds=8):
        """Set-up, not training: runs the step's two collectives `rounds` times on scratch buffers of the real sizes so that
        NCCL's lazily created channels / connections / algorithm choices for these message sizes exist before the first
        timed step (a training loop otherwise pays for them during its first ~10 steps)."""
        dev = self.params[0].device
        big = torch.zeros(self.table_param.numel() if self.table_param is not None else 1 << 20, dtype=torch.float32, device=dev)
        small = torch.zeros(sum((p.numel() + 3) // 4 * 4 for p in self.params if p is not self.table_param) or 4,
                            dtype=torch.float32, device=dev)
        for _ in range(rounds):
            w1 = dist.all_reduce(big, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            w2 = dist.all_reduce(small, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            w1.wait()
            w2.wait()
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    # -- hook protocol used by ops.NewsCNN.backward -------------------------------------------------------------
    def event(self, device):
        key = torch.device(device).index
        if key not in self._events:
            ev = torch.cuda.Event()
            ev.record()                      # materialises the cudaEvent_t so that the C library can record it
            self._events[key] = ev
        return self._events[key]

    def __call__(self, d_table, ev):
        tp = self.table_param
        if tp is None or d_table.shape != tp.shape:
            return False                     # not ours: leave it to autograd
        side = ops.side_stream(d_table.device)
        side.wait_event(ev)                  # d_table is complete at `ev`; the rest of the backward keeps running
        with torch.cuda.stream(side):
            work = dist.all_reduce(d_table, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append(work)
        if tp.grad is None:
            tp.grad = d_table                # assigned here, not through AccumulateGrad (which would clone it, see ops.TABLE_GRAD_HOOK)
        else:
            # a second encoder backward in the same step (encode_news and encode_user called separately, as the reference's
            # forward does): this contribution gets its own all-reduce and is added to .grad in finish(), after both are done
            self._extra.append(d_table)
        return True

    # -- after loss.backward() ----------------------------------------------------------------------------------
    def finish(self):
        own = {t.data_ptr() for t in self._extra}
        if self.table_param is not None and self.table_param.grad is not None and self._pending:
            own.add(self.table_param.grad.data_ptr())
        rest = [p for p in self.params if p.grad is not None and p.grad.data_ptr() not in own]
        if rest:
            # one static flat buffer, every segment starting on a 16-byte boundary (the Adam kernel reads float4)
            key = tuple((id(p), p.grad.numel()) for p in rest)
            if self._flat_key != key:
                total = sum((n + 3) // 4 * 4 for _, n in key)
                self._flat = torch.zeros(total, dtype=torch.float32, device=rest[0].grad.device)
                self._flat_key = key
            views, o = [], 0
            for p in rest:
                n = p.grad.numel()
                views.append(self._flat[o:o + n])
                o += (n + 3) // 4 * 4
            torch._foreach_copy_(views, [p.grad.reshape(-1) for p in rest])
            self._pending.append(dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for p, v in zip(rest, views):
                p.grad = v.view_as(p.grad)
        for work in self._pending:
            work.wait()                      # the compute stream waits for the collective (no host synchronisation)
        for t in self._extra:
            self.table_param.grad.add_(t)
        self._pending, self._extra = [], []

    def close(self):
        if ops.TABLE_GRAD_HOOK is self:
            ops.TABLE_GRAD_HOOK = None
        self.optimizer.grad_scale = 1.0


def train_step(model, x, optimizer, sync=None):
    """One Manager._train iteration; returns the (device) loss tensor without synchronising.  `sync` (a GradSync)
    averages the gradients over the data-parallel group before the optimiser step."""
    optimizer.zero_grad(set_to_none=True)
    logp = model(x)[0]
    core = model.module if hasattr(model, "module") else model
    loss = ops.NLLMean.apply(logp, x["label"].to(core.device, non_blocking=True))
    loss.backward()
    if sync is not None:
        sync.finish()
    optimizer.step()
    return loss


def to_device(x, device):
    return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) and k != "his_mask" else v) for k, v in x.items()}


class BatchPrefetcher:
    """Host -> device staging of the NEXT batch on a side stream while the current step computes.

    The reference copies every field synchronously inside the model (TwoTower.py:24-27,38-41: ``.to(self.device)``
    on pageable DataLoader tensors).  Here the int64 id / mask tensors of batch i+1 (7.2 MB at the MIND-small
    shape) cross PCIe during step i.  ``his_mask`` is staged too: left on the host (as the reference does, RNN.py:65)
    its lengths would reach the device through a pageable, stream-ordered copy that stalls the host until the
    previous step has drained."""

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.bufs = [None] * depth          # static device copies of the batch dict (no allocator traffic per step)
        self.free = [None] * depth          # event: the step that consumed buffer k has been queued and finished
        self.count = 0

    def _fits(self, buf, x):
        return buf is not None and all(torch.is_tensor(v) == (k in buf) and (not torch.is_tensor(v) or (
            buf[k].shape == v.shape and buf[k].dtype == v.dtype)) for k, v in x.items())

    def stage(self, x):
        k = self.count % self.depth
        self.count += 1
        if not self._fits(self.bufs[k], x):
            self.bufs[k] = {key: torch.empty(v.shape, dtype=v.dtype, device=self.device) for key, v in x.items() if torch.is_tensor(v)}
            self.free[k] = None
        with torch.cuda.stream(self.stream):
            if self.free[k] is not None:
                self.stream.wait_event(self.free[k])
            else:
                self.stream.wait_stream(torch.cuda.current_stream(self.device))     # fresh buffers: allocated on the compute stream
            for key, dst in self.bufs[k].items():
                dst.copy_(x[key], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        extra = {key: v for key, v in x.items() if not torch.is_tensor(v)}
        return k, ev, extra

    def take(self, staged):
        k, ev, extra = staged
        torch.cuda.current_stream(self.device).wait_event(ev)
        out = dict(self.bufs[k])
        out.update(extra)
        return out

    def release(self, staged):
        """call after the step that consumed `staged` has been queued: its buffer may be overwritten once that step is done"""
        k = staged[0]
        self.free[k] = torch.cuda.Event()
        self.free[k].record(torch.cuda.current_stream(self.device))


class GraphStep:
    """The whole training step (zero_grad, forward, NLLLoss, backward, gradient all-reduce, Adam) captured ONCE as a CUDA
    graph and replayed: the host then issues one graph launch per step instead of ~75 kernel launches through Python /
    ctypes (1.2-1.5 ms of host work against 1.5 ms of device work).  Inputs are copied into static device tensors; every
    step-dependent optimiser scalar (Adam bias corrections, the learning rates a schedule moves, grad_scale) comes from a
    device block that FusedAdam.begin_step() refreshes before each replay, so LinearWarmupSchedule / load_state_dict keep
    working; the loss is a static device scalar.  Parameters, Adam moments and the step counter are snapshotted before the
    three warm-up steps the capture needs and restored afterwards: constructing a GraphStep does not train.

    With a GradSync the NCCL all-reduces are captured too (thread-local capture mode; all of them on ProcessGroupNCCL's one
    internal stream, i.e. ONE dependency chain inside the graph -- see GradSync).  Every rank must construct the GraphStep
    at the same point.  Every call must bring tensors of the captured shapes."""

    def __init__(self, model, optimizer, example, sync=None, warmup_steps=3):
        core = model.module if hasattr(model, "module") else model
        self.model, self.optimizer, self.device, self.sync = model, optimizer, torch.device(core.device), sync
        self.static_x = {k: (v.to(self.device).clone() if torch.is_tensor(v) else v) for k, v in example.items()}
        if optimizer.dyn is None:
            optimizer.enable_device_step_scalars(self.device)
        # ---- snapshot: the warm-up steps below are real optimiser steps on the example batch
        params = [p for g in optimizer.param_groups for p in g["params"]]
        snap_p = [p.detach().clone() for p in params]
        snap_state = {p: (st[0].clone(), st[1].clone()) for p, st in optimizer.state.items()}
        snap_steps = optimizer.steps
        # warm-up on a side stream (allocator state, lazy initialisation), as torch.cuda.graphs requires
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup_steps):
                optimizer.begin_step()
                train_step(model, self.static_x, optimizer, sync)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        mode = {"capture_error_mode": "thread_local"} if sync is not None else {}
        with torch.cuda.graph(self.graph, **mode):          # recorded, not executed: the step counter is not advanced here
            self.loss = train_step(model, self.static_x, optimizer, sync)
        torch.cuda.synchronize(self.device)
        # ---- restore IN PLACE (the graph holds the addresses of the parameters and of the moment tensors)
        with torch.no_grad():
            for p, old in zip(params, snap_p):
                p.copy_(old)
            for p, st in optimizer.state.items():
                old = snap_state.get(p)
                if old is None:
                    st[0].zero_()
                    st[1].zero_()
                else:
                    st[0].copy_(old[0])
                    st[1].copy_(old[1])
        optimizer.steps = snap_steps
        emb = getattr(core, "embedding", None)
        if emb is not None and hasattr(emb, "refresh_shadow_inplace"):
            emb.refresh_shadow_inplace()                     # the captured forward gathers from this very buffer
        torch.cuda.synchronize(self.device)
        self.warmup_steps = warmup_steps

    def __call__(self, x):
        for k, dst in self.static_x.items():
            if torch.is_tensor(dst):
                dst.copy_(x[k], non_blocking=True)
        self.optimizer.begin_step()
        self.graph.replay()
        return self.loss

    def close(self):
        """Release the captured graph.  REQUIRED before torch.distributed.destroy_process_group() when a GradSync was captured:
        a live CUDA graph keeps the communicator's captured kernels referenced and the communicator teardown then hangs."""
        if self.graph is not None:
            torch.cuda.synchronize(self.device)
            self.graph.reset()
            self.graph = None


class TrainLoop:
    """The training loop over host (pinned) batches: input prefetch + lagged loss read.  The loss of step i is copied
    to pinned host memory right behind step i on the compute stream (+ an event) and read by the host after step i+1
    has been queued, so the device never waits for the host; the last loss is read at the end of run().  The staging
    buffers, the pinned loss slots and the events are created once (cudaHostAlloc is a multi-millisecond, device-
    synchronising call)."""

    def __init__(self, model, optimizer, sync=None, graph_step=None):
        self.model, self.optimizer, self.sync, self.graph_step = model, optimizer, sync, graph_step
        core = model.module if hasattr(model, "module") else model
        self.prefetch = BatchPrefetcher(core.device)
        self.slots = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event() for _ in range(2)]

    def run(self, host_batches, steps, on_loss=None):
        pf, slots, events = self.prefetch, self.slots, self.events
        n = len(host_batches)
        losses = []

        def read(i):
            events[i % 2].synchronize()
            losses.append(float(slots[i % 2][0]))
            if on_loss is not None:
                on_loss(i, losses[-1])

        staged = pf.stage(host_batches[0]) if steps > 0 else None
        for s in range(steps):
            cur = staged
            x = pf.take(cur)
            if s + 1 < steps:
                staged = pf.stage(host_batches[(s + 1) % n])
            loss = self.graph_step(x) if self.graph_step is not None else train_step(self.model, x, self.optimizer, self.sync)
            pf.release(cur)
            slots[s % 2].copy_(loss.detach().reshape(1), non_blocking=True)        # device -> host read of the step's result
            events[s % 2].record()
            if s >= 1:
                read(s - 1)
        if steps >= 1:
            read(steps - 1)
        return losses


def run_steps(model, host_batches, optimizer, steps, on_loss=None):
    """One-shot convenience wrapper around TrainLoop (pays its set-up cost on every call)."""
    return TrainLoop(model, optimizer).run(host_batches, steps, on_loss)

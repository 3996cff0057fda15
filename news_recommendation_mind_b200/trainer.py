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

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self):
        self.steps += 1
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self.state.get(p)
                if st is None:
                    st = self.state[p] = (torch.zeros_like(p), torch.zeros_like(p))
                shadow = None
                if self._want_shadow and p is self.embedding.weight:
                    shadow = self.embedding.shadow_bf16()
                ops.adam_step(p.data, p.grad.contiguous(), st[0], st[1], self.steps, g["lr"], self.betas[0], self.betas[1],
                              self.eps, self.grad_scale, shadow)
                if shadow is not None:
                    self.embedding.mark_shadow_fresh(shadow)


def train_step(model, x, optimizer):
    """One Manager._train iteration; returns the (device) loss tensor without synchronising."""
    optimizer.zero_grad(set_to_none=True)
    logp = model(x)[0]
    core = model.module if hasattr(model, "module") else model
    loss = ops.NLLMean.apply(logp, x["label"].to(core.device, non_blocking=True))
    loss.backward()
    optimizer.step()
    return loss


def to_device(x, device):
    return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) and k != "his_mask" else v) for k, v in x.items()}

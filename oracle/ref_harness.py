"""Drives the UNMODIFIED reference (oracle/_ref, staged by oracle/build_ref.py; or /root/reference in the build container)
on the host cores -- TEST / BENCH INFRASTRUCTURE ONLY.

What is the reference's and what is ours:
  * models.TwoTower, CNN_Encoder, RNN_User_Encoder, ... , Attention.py: the reference's files, imported as they are;
  * the harness shims of SURVEY.md 8(c), none of which edits a reference file (see oracle/make_golden.py, whose builders are
    reused here): torch-2.x signature of _softmax_backward_data, BERT_Embedding built without the network download
    (nn.Embedding(V, E, padding_idx=0) under the reference's attribute name), intended-semantics adapters for the
    MHA / LSTUR user encoders that are broken as shipped;
  * the loop below restates utils/Manager.py:636-647 (zero_grad, forward, NLLLoss, backward, optimizer.step) and
    Manager._get_optim (Manager.py:389-413: Adam, "bert" parameters at bert_lr) -- Manager itself needs the MIND files on
    disk and a CLI, so it is not instantiated.
"""
from __future__ import annotations

import os
import sys
import time

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """oracle/_ref when staged (it is what travels to the GPU box), else the read-only mount, else None."""
    from oracle import build_ref
    if build_ref.available():
        return build_ref.DEST
    if os.path.isdir(build_ref.SRC):
        return build_ref.SRC
    return None


_R = None


def load():
    """-> namespace of the reference's classes (with shim 1 applied); raises when the reference is not available."""
    global _R
    if _R is not None:
        return _R
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not staged: run oracle/build_ref.py where /root/reference exists")
    from oracle import make_golden as G
    G.REF = root
    _R = G._import_reference()
    _R.root = root
    return _R


def build_model(encoder_n="cnn", encoder_u="lstm", V=30522, E=300, H=150, C=5, S=50, L=32, hn=10, n_users=40, seed=42,
                state=None, dropout_p=0.2):
    from oracle import make_golden as G
    R = load()
    torch.manual_seed(seed)
    man = G.fake_manager(encoderN=encoder_n, encoderU=encoder_u, cdd_size=C, his_size=S, signal_length=L, bert_dim=E,
                         hidden_dim=H, head_num=hn, n_users=n_users, dropout_p=dropout_p)
    model = G.build_reference_model(R, man, V)
    if state is not None:
        model.load_state_dict(state, strict=False)
    return model


def make_optimizer(model, lr=1e-4, bert_lr=6e-6):
    """Manager._get_optim (Manager.py:389-413)."""
    import re
    base, bert = [], []
    for name, p in model.named_parameters():
        (bert if re.search("bert", name) else base).append(p)
    return torch.optim.Adam([{"params": base, "lr": lr}, {"params": bert, "lr": bert_lr}])


def train_step(model, optimizer, x, loss_func=None):
    """one iteration of Manager._train (Manager.py:636-647)"""
    loss_func = loss_func or nn.NLLLoss()
    optimizer.zero_grad(set_to_none=True)
    loss = loss_func(model(x)[0], x["label"].to(model.device))
    loss.backward()
    optimizer.step()
    return loss


def time_training(batches, warmup, steps, threads=None, **model_kw):
    """-> (seconds per timed step list, cores) for the reference model over `batches` (cycled)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_model(**model_kw)
    model.train()
    opt = make_optimizer(model)
    ts = []
    for s in range(warmup + steps):
        x = batches[s % len(batches)]
        t0 = time.perf_counter()
        train_step(model, opt, x)
        if s >= warmup:
            ts.append(time.perf_counter() - t0)
    return ts, torch.get_num_threads()

"""Live cross-check of the oracle against the reference ITSELF (oracle/_ref staged from /root/reference, or the mount), on seeds
that are NOT in the committed golden files: guards against goldens that only pin the cases they were generated for.
Skipped where the reference is not available (it never enters the repository)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import ref_harness as RH            # noqa: E402
from oracle import twotower_oracle as O         # noqa: E402
from helpers import random_batch                # noqa: E402

pytestmark = pytest.mark.skipif(RH.reference_root() is None, reason="reference not staged / mounted here")


@pytest.mark.parametrize("encn,encu,seed", [("cnn", "lstm", 101), ("cnn", "gru", 102), ("cnn", "attn", 103), ("cnn", "avg", 104),
                                            ("cnn", "mha", 105), ("mha", "lstm", 106)])
def test_oracle_matches_the_live_reference_on_fresh_seeds(encn, encu, seed):
    B, C, S, L, E, H, V, hn = 5, 4, 7, 12, 40, 20, 300, 5
    torch.set_num_threads(4)
    model = RH.build_model(encn, encu, V=V, E=E, H=H, C=C, S=S, L=L, hn=hn, n_users=40, seed=seed, dropout_p=0.0)
    model.train()
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():                                   # non-trivial values everywhere (biases and the padding row included)
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
    x = random_batch(gen, B, C, S, L, V)
    logp = model(x)[0]
    loss = torch.nn.NLLLoss()(logp, x["label"])
    loss.backward()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ours = O.forward(params, x, True, encoder_n=encn, encoder_u=encu, head_num=hn)
    O.nll_loss(ours, x["label"]).backward()
    assert torch.allclose(ours, logp.detach(), rtol=1e-5, atol=1e-6), float((ours - logp).abs().max())
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        g = params[name].grad
        assert g is not None, name
        scale = float(p.grad.abs().max()) + 1e-12
        assert float((g - p.grad).abs().max()) <= 2e-5 * scale + 1e-7, (name, float((g - p.grad).abs().max()), scale)


def test_metrics_oracle_matches_the_live_cal_metric():
    """cal_metric (Manager.py:1276-1344: sklearn AUC, mrr / dcg through np.argsort(score)[::-1]) on fresh impressions against
    oracle/metrics_oracle.py: all four metrics on tie-free scores, AUC also on heavily tied ones.

    MRR / nDCG of TIED scores are not defined by the reference: np.argsort's default sort is unstable (numpy >= 2 sorts float64
    with a SIMD network even for 5 elements -- probed on this build --, numpy 1.x insertion-sorts up to 16 elements and
    introsorts beyond), so the order among equal scores depends on the numpy build and the CPU.  The oracle and mr_rank_metrics
    use the numpy-1.x small-array behaviour everywhere (stable sort, reversed: the later position first); AUC is order free."""
    import numpy as np
    from oracle import metrics_oracle as M
    root = RH.reference_root()
    sys.path.insert(0, root)
    try:
        from utils.Manager import cal_metric
    finally:
        sys.path.remove(root)
    rng = np.random.default_rng(2024)

    def impressions(count, lo, hi, tied):
        labels, preds = [], []
        for _ in range(count):
            n = int(rng.integers(lo, hi))
            y = (rng.random(n) < 0.12).astype(np.float64)
            y[int(rng.integers(0, n))] = 1.0
            if y.sum() == n:
                y[0] = 0.0
            p = np.round(rng.random(n) * 5) / 5 if tied else rng.permutation(n) / n + rng.random() * 1e-3
            labels.append(y.tolist())
            preds.append((1.0 / (1.0 + np.exp(-p))).tolist())
        return labels, preds
    labels, preds = impressions(300, 2, 300, tied=False)
    ref = cal_metric(labels, preds, ["auc", "mean_mrr", "ndcg@5", "ndcg@10"])
    ours = M.ranking_metrics(labels, preds)
    assert ours == {k: float(ref[k]) for k in ours}, (ours, ref)
    labels, preds = impressions(200, 2, 200, tied=True)
    assert M.ranking_metrics(labels, preds)["auc"] == float(cal_metric(labels, preds, ["auc"])["auc"])

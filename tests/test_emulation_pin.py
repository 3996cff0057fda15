"""CPU pin of the bf16 emulation the GPU tests hold the tensor-core kernels to (tests/test_gpu_tc.py:
cnn_encoder_bf16_emulation / cnn_encoder_bf16_backward_emulation).  The emulation restates the encoder's forward and backward by
hand with bf16 rounding at the kernels' storage points; with the rounding function replaced by the identity it must BE the
oracle (oracle.twotower_oracle.cnn_news_encoder + autograd, itself pinned to the reference), in both backward forms
(per-token and token-grouped).  What remains specific to the kernels is then only WHERE they round, not the mathematics."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import twotower_oracle as O   # noqa: E402


@pytest.mark.parametrize("grouped", [False, True])
@pytest.mark.parametrize("N,L,E,H,V", [(7, 12, 24, 16, 50), (5, 32, 40, 20, 90)])
def test_emulation_without_rounding_is_the_oracle(monkeypatch, N, L, E, H, V, grouped):
    import test_gpu_tc as T
    monkeypatch.setattr(T, "_bf", lambda t: t.detach().cpu().double())
    gen = torch.Generator().manual_seed(N * 100 + L)
    table = torch.randn(V, E, generator=gen, dtype=torch.float64) * 0.3
    ids = torch.randint(1, V, (N, L), generator=gen)
    ln = torch.randint(2, L + 1, (N,), generator=gen)
    ln[0] = 0                                                       # an all-masked title: zero vector, zero gradients through the pooling
    mask = (torch.arange(L)[None, :] < ln[:, None]).long()
    ids = ids * mask                                                # padding positions hold token 0 (and are convolved, CNN.py:41)
    conv_w = torch.randn(H, E, 3, generator=gen, dtype=torch.float64) * 0.1
    conv_b = torch.randn(H, generator=gen, dtype=torch.float64) * 0.1
    proj_w = torch.randn(H, H, generator=gen, dtype=torch.float64) * 0.2
    proj_b = torch.randn(H, generator=gen, dtype=torch.float64) * 0.1
    query = torch.randn(1, H, generator=gen, dtype=torch.float64)
    g = torch.randn(N, H, generator=gen, dtype=torch.float64)
    # ---- oracle, float64, autograd
    P = [t.clone().requires_grad_(True) for t in (table, conv_w, conv_b, proj_w, proj_b, query)]
    c_o, news_o = O.cnn_news_encoder(P[0][ids], mask, *P[1:])
    (news_o * g).sum().backward()
    d_table_o = P[0].grad.clone()
    d_table_o[0] = 0                                                # BertEmbeddings' padding_idx = 0
    # ---- emulation with the rounding switched off (its `.float()` casts at the storage points remain: fp32-level agreement)
    c_e, news_e, p_e = T.cnn_encoder_bf16_emulation(table, ids, mask, conv_w, conv_b, proj_w, proj_b, query)
    assert float(news_e[0].abs().max()) == 0.0
    torch.testing.assert_close(c_e, c_o.detach(), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(news_e, news_o.detach(), rtol=1e-6, atol=1e-6)
    grads = T.cnn_encoder_bf16_backward_emulation(table, ids, mask, conv_w, conv_b, proj_w, proj_b, query, g, grouped=grouped)
    want = {"table": d_table_o, "cnn.weight": P[1].grad, "cnn.bias": P[2].grad, "wordQueryProject.weight": P[3].grad,
            "wordQueryProject.bias": P[4].grad, "query_words": P[5].grad}
    assert set(grads) == set(want)
    for k, w in want.items():
        scale = float(w.abs().max()) + 1e-30
        err = float((grads[k].reshape(w.shape) - w).abs().max())
        assert err <= 2e-6 * scale, (k, err, scale)

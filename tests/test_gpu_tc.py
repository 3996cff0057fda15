"""Pins the tcgen05 shared-memory descriptor conventions of csrc/tc05.cuh on hardware: one MMA tile
through mr_tc_selftest for every operand majorness / row shift the production kernels use, against
a bf16-rounded fp32 matmul (tolerance 1e-5 relative: only the accumulation order differs)."""
import pytest
import torch

from news_recommendation_mind_b200 import _lib
from news_recommendation_mind_b200._lib import check, ptr, stream_ptr

pytestmark = pytest.mark.gpu


def run_tile(a, b, a_mn, b_mn, N, K, shift, halo, swap=0):
    lib = _lib.load()
    d = torch.empty(128, N, dtype=torch.float32, device="cuda")
    check(lib.mr_tc_selftest(ptr(a), a.shape[0], a.shape[1], ptr(b), b.shape[0], b.shape[1], ptr(d), a_mn, b_mn, N, K,
                             shift, halo, swap, stream_ptr("cuda")), "mr_tc_selftest")
    torch.cuda.synchronize()
    return d


def expected(a, b, a_mn, b_mn, N, K, shift):
    ar = a.bfloat16().float()
    br = b.bfloat16().float()
    if a_mn:                      # rows = K index
        A = torch.zeros(K, 128, device="cuda")
        for k in range(K):
            if 0 <= k + shift < ar.shape[0]:
                A[k] = ar[k + shift]
        A = A.t()
    else:
        A = torch.zeros(128, K, device="cuda")
        for m in range(128):
            if 0 <= m + shift < 128:
                A[m] = ar[m + shift, :K]
    B = br[:K, :N].t() if b_mn else br[:N, :K]
    return A.double() @ B.double().t()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("N,K,shift", [(160, 304, 0), (160, 304, -4), (160, 304, 4), (144, 160, 1), (256, 64, 0), (16, 16, 0)])
def test_umma_tile(a_mn, b_mn, N, K, shift):
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + K + shift + 7 * a_mn + 3 * b_mn)
    a = torch.randn((K, 128) if a_mn else (128, K), generator=g, device="cuda")
    b = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda")
    d = run_tile(a, b, a_mn, b_mn, N, K, shift, 8)
    ref = expected(a, b, a_mn, b_mn, N, K, shift)
    err = float((d.double() - ref).norm() / ref.norm())
    assert err < 1e-5, err


# ------------------------------------------------------------------------------------------------
# bf16 (tcgen05) news encoder against the oracle
# ------------------------------------------------------------------------------------------------
def _bf(t):
    return t.detach().cpu().bfloat16().double()


def cnn_encoder_bf16_emulation(table, ids, mask, conv_w, conv_b, proj_w, proj_b, query):
    """The oracle's CNN news encoder (oracle.twotower_oracle.cnn_news_encoder, CNN.py:30-51) evaluated in
    float64 with bf16 rounding applied exactly where the MR_BF16 kernels round: table, conv/proj weights,
    the stored c and key tensors."""
    import math
    from oracle import twotower_oracle as O
    x = _bf(table)[ids.cpu()]
    lead = x.shape[:-2]
    L, E = x.shape[-2:]
    H = conv_w.shape[0]
    xr = x.reshape(-1, L, E)
    xpad = torch.nn.functional.pad(xr, (0, 0, 1, 1))
    cw = _bf(conv_w)
    c = conv_b.detach().cpu().double().view(1, 1, H).expand(xr.shape[0], L, H).clone()
    for tap in range(3):
        c = c + xpad[:, tap:tap + L, :] @ cw[:, :, tap].t()
    c = _bf(torch.relu(c).float())
    key = _bf(torch.tanh(c @ _bf(proj_w).t() + proj_b.detach().cpu().double()).float())
    s = (key @ query.detach().cpu().double().view(H, 1)).squeeze(-1) / math.sqrt(H)
    p = O.masked_softmax(s, mask.cpu().reshape(-1, L))
    news = (p.unsqueeze(-1) * c).sum(1)
    return c.view(*lead, L, H), news.view(*lead, H), p


@pytest.mark.parametrize("N,L,E,H", [(37, 32, 300, 150), (4, 32, 300, 150), (1, 32, 300, 150), (23, 30, 300, 150),
                                      (9, 48, 64, 32), (130, 20, 768, 150), (515, 32, 300, 150)])
def test_news_cnn_bf16_forward(N, L, E, H):
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import manager_for, rel_err
    import news_recommendation_mind_b200 as mr
    from oracle import twotower_oracle as O
    torch.manual_seed(N + L)
    V = 997
    man = manager_for("cnn", "lstm", 5, 50, L, E, H, 10, precision="bf16")
    emb = mr.BERT_Embedding(man, vocab_size=V).cuda()
    enc = mr.CNN_Encoder(man).cuda()
    with torch.no_grad():
        emb.weight.normal_(0, 0.3)
        enc.cnn.bias.normal_(0, 0.1)
    gen = torch.Generator().manual_seed(1)
    ln = torch.randint(0, L + 1, (N,), generator=gen)
    ln[0] = L
    ids = torch.randint(1, V, (N, L), generator=gen)
    mask = (torch.arange(L)[None, :] < ln[:, None]).long()
    ids = ids * mask
    with torch.no_grad():
        news = enc.encode_ids(emb, ids.cuda(), mask.cuda())
    c_e, news_e, _ = cnn_encoder_bf16_emulation(emb.weight, ids, mask, enc.cnn.weight, enc.cnn.bias, enc.wordQueryProject.weight,
                                                enc.wordQueryProject.bias, enc.query_words)
    err = rel_err(news, news_e)
    _, news_f = O.cnn_news_encoder(O.embed_tokens(emb.weight.detach().cpu(), ids), mask, enc.cnn.weight.detach().cpu(),
                                   enc.cnn.bias.detach().cpu(), enc.wordQueryProject.weight.detach().cpu(),
                                   enc.wordQueryProject.bias.detach().cpu(), enc.query_words.detach().cpu())
    err32 = rel_err(news, news_f)
    print("N=%d L=%d E=%d H=%d: vs bf16-emulated oracle %.3e, vs fp32 oracle %.3e" % (N, L, E, H, err, err32))
    assert err < 2e-3, err                      # same roundings, only accumulation order + tanh.approx differ
    assert err32 < 1e-2, err32                  # bf16 operand rounding vs the fp32 reference
    # all-masked titles -> exact zeros (XSoftmax semantics, Attention.py:66-74)
    dead = (ln == 0)
    if dead.any():
        assert float(news[dead.cuda()].abs().max()) == 0.0

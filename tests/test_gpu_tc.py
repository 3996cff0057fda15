"""Pins the tcgen05 shared-memory descriptor conventions of csrc/tc05.cuh on hardware: one MMA tile
through mr_tc_selftest for every operand majorness / row shift the production kernels use, against
a bf16-rounded fp32 matmul (tolerance 1e-5 relative: only the accumulation order differs)."""
import pytest
import torch

from news_recommendation_mind_b200 import _lib
from news_recommendation_mind_b200._lib import check, ptr, stream_ptr

pytestmark = pytest.mark.gpu


def run_tile(a, b, a_mn, b_mn, N, K, shift, halo, swap=0, a_layout=0, b_layout=0, base_off_mode=0):
    lib = _lib.load()
    d = torch.empty(128, N, dtype=torch.float32, device="cuda")
    check(lib.mr_tc_selftest(ptr(a), a.shape[0], a.shape[1], ptr(b), b.shape[0], b.shape[1], ptr(d), a_mn, b_mn, N, K,
                             shift, halo, swap, a_layout, b_layout, base_off_mode, stream_ptr("cuda")), "mr_tc_selftest")
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


@pytest.mark.parametrize("a_mn,b_mn,a_layout,b_layout,N,K", [
    (0, 0, 4, 0, 160, 160),      # cnn_tail D1: A = dkp, K-major rows of 64 B (SWIZZLE_64B blocks of 32 columns); B = Wq panels
    (1, 1, 4, 4, 160, 128),      # cnn_tail D2: both operands MN-major views of SWIZZLE_64B tiles (K = token rows)
    (0, 0, 2, 0, 160, 128),      # tap GEMM: A K-major SWIZZLE_128B
    (1, 1, 2, 4, 160, 128),      # token-reduction GEMM: P SWIZZLE_128B, Q SWIZZLE_64B, both MN-major
    (1, 1, 4, 4, 32, 64),
    (0, 0, 8, 0, 16, 160),       # rnn_tc: A (W_hh) read from tensor memory, B = hidden states (16 rows), panel layout
    (0, 1, 8, 0, 64, 160),
])
def test_umma_tile_swizzled_layouts(a_mn, b_mn, a_layout, b_layout, N, K):
    g = torch.Generator(device="cuda").manual_seed(N + 31 * K + a_layout + 5 * b_layout)
    a = torch.randn((K, 128) if a_mn else (128, K), generator=g, device="cuda")
    b = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda")
    d = run_tile(a, b, a_mn, b_mn, N, K, 0, 0, 0, a_layout, b_layout)
    ref = expected(a, b, a_mn, b_mn, N, K, 0)
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


@pytest.mark.parametrize("B,C,S,L,E,H", [(8, 5, 50, 32, 300, 150), (3, 2, 7, 30, 300, 150), (2, 3, 5, 48, 64, 32), (2, 2, 3, 20, 768, 150)])
def test_twotower_bf16_training_step_vs_fp32_oracle(B, C, S, L, E, H):
    """Whole TwoTower training step (CNN news encoder on tcgen05 in bf16, LSTM user encoder) against the fp32
    oracle.  north_star bound for bf16 logits is 1e-3 relative.  Gradients of the news encoder are sums with heavy
    cancellation (softmax backward sums to zero over each title), which amplifies the 2^-9 operand rounding of
    bf16 to a few percent relative L2 against the fp32 oracle; the tight check of the backward kernels is
    test_news_cnn_bf16_backward (same rounding points, <= 5e-3)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import build_model, manager_for, rel_err, random_batch
    from oracle import twotower_oracle as O
    gen = torch.Generator().manual_seed(42)
    torch.manual_seed(42)
    V = 3000
    man = manager_for("cnn", "lstm", C, S, L, E, H, 10, precision="bf16")
    model = build_model(man, V)
    with torch.no_grad():
        model.embedding.weight.normal_(0, 0.3)
    x = random_batch(gen, B, C, S, L, V)
    params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ref = O.forward(params, x, True, encoder_n="cnn", encoder_u="lstm")
    O.nll_loss(ref, x["label"]).backward()
    model.train()
    logp = model(x)[0]
    torch.nn.NLLLoss()(logp, x["label"].cuda()).backward()
    torch.cuda.synchronize()
    e_logit = rel_err(logp, ref)
    print("logp rel err %.3e" % e_logit)
    worst = 0.0
    for k, p in model.named_parameters():
        e = rel_err(p.grad, params[k].grad)
        print("  grad %-40s rel err %.3e" % (k, e))
        worst = max(worst, e)
    assert e_logit < 1e-3, e_logit
    assert worst < 0.15, worst
    assert float(model.embedding.weight.grad[0].abs().max()) == 0.0


def test_train_loop_prefetch_matches_plain_steps():
    """trainer.TrainLoop (host batches staged on a side stream into static buffers, losses read with one step of lag)
    produces bit-identical losses and parameters to calling train_step on device-resident copies of the same batches."""
    import sys, os, copy
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import build_model, manager_for, random_batch
    from news_recommendation_mind_b200 import trainer
    B, C, S, L, E, H, V = 6, 5, 9, 32, 300, 150, 500
    gen = torch.Generator().manual_seed(3)
    host = [{k: v.pin_memory() for k, v in random_batch(gen, B, C, S, L, V).items()} for _ in range(3)]
    torch.manual_seed(5)
    man = manager_for("cnn", "lstm", C, S, L, E, H, 10, precision="bf16")
    m1 = build_model(man, V)
    m2 = copy.deepcopy(m1)
    o1, o2 = trainer.FusedAdam(m1, lr=1e-3, bert_lr=1e-4), trainer.FusedAdam(m2, lr=1e-3, bert_lr=1e-4)
    seen = []
    got = trainer.TrainLoop(m1, o1).run(host, 7, on_loss=lambda i, v: seen.append(i))
    exp = [float(trainer.train_step(m2, {k: v.cuda() for k, v in host[s % 3].items()}, o2)) for s in range(7)]
    assert seen == list(range(7))
    assert got == exp, (got, exp)
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), k


def cnn_encoder_bf16_backward_emulation(table, ids, mask, conv_w, conv_b, proj_w, proj_b, query, g, grouped=False):
    """float64 restatement of the backward of oracle.twotower_oracle.cnn_news_encoder (autograd of CNN.py:30-51,
    softmax backward of Attention.py:77-80) with bf16 rounding at the points where the MR_BF16 kernels store
    bf16: c, key, dkey_pre, p*d_news, dconv, d_emb."""
    import math
    c, news, p = cnn_encoder_bf16_emulation(table, ids, mask, conv_w, conv_b, proj_w, proj_b, query)
    N, L = ids.shape
    H, E = conv_w.shape[0], conv_w.shape[1]
    c = c.reshape(N, L, H)
    x = _bf(table)[ids.cpu()].reshape(N, L, E)
    wq = _bf(proj_w)
    key = _bf(torch.tanh(c @ wq.t() + proj_b.detach().cpu().double()).float())
    q = query.detach().cpu().double().view(H)
    g = g.detach().cpu().double().view(N, 1, H)
    dp = (g * c).sum(-1)
    dot = (p * dp).sum(-1, keepdim=True)
    ds = p * (dp - dot) / math.sqrt(H)
    d_q = (ds.unsqueeze(-1) * key).sum((0, 1))
    dkp_f = ds.unsqueeze(-1) * q * (1 - key * key)
    dkp = _bf(dkp_f.float())
    dcp = p.unsqueeze(-1) * g
    if L > 32 or H > 160:                      # generic pooling path: p * d_news is stored as bf16 (fast path: formed in fp32
        dcp = _bf(dcp.float())                 # inside the RELUGRAD_POOL epilogue, never stored)
    dconv_f = (c > 0).double() * (dcp + dkp @ wq)
    dconv = _bf(dconv_f.float())
    d_proj_b = dkp_f.sum((0, 1))               # bias gradients are summed in fp32 before the bf16 rounding of the stored tensor
    d_proj_w = torch.einsum("nlh,nlk->hk", dkp, c)
    d_conv_b = dconv_f.sum((0, 1))
    gpad = torch.nn.functional.pad(dconv, (0, 0, 1, 1))
    cw = _bf(conv_w)
    if grouped:
        # token-grouped backward (mr_news_cnn_bwd_table): S[v, tap] = bf16(sum_{t: ids[t]=v} dconv[t+1-tap]) is the rounding
        # point; the table and filter gradients are GEMMs over vocabulary rows
        V = table.shape[0]
        flat = ids.cpu().reshape(-1)
        S = [_bf(torch.zeros(V, H, dtype=torch.float64).index_add_(0, flat, gpad[:, 2 - tap:2 - tap + L].reshape(-1, H)).float())
             for tap in range(3)]
        tb = _bf(table)
        d_conv_w = torch.stack([S[tap].t() @ tb for tap in range(3)], dim=-1)
        d_table = sum(S[tap] @ cw[:, :, tap] for tap in range(3))
    else:
        xpad = torch.nn.functional.pad(x, (0, 0, 1, 1))
        d_conv_w = torch.stack([torch.einsum("nlh,nle->he", dconv, xpad[:, tap:tap + L]) for tap in range(3)], dim=-1)
        d_x = sum(gpad[:, 2 - tap:2 - tap + L] @ cw[:, :, tap] for tap in range(3))
        d_x = _bf(d_x.float())
        d_table = torch.zeros(table.shape, dtype=torch.float64).index_add_(0, ids.cpu().reshape(-1), d_x.reshape(-1, E))
    d_table[0] = 0
    return {"cnn.weight": d_conv_w, "cnn.bias": d_conv_b, "wordQueryProject.weight": d_proj_w,
            "wordQueryProject.bias": d_proj_b, "query_words": d_q.view(1, H), "table": d_table}


@pytest.mark.parametrize("grouped", [True, False])
@pytest.mark.parametrize("N,L,E,H", [(37, 32, 300, 150), (23, 30, 300, 150), (9, 48, 64, 32), (130, 20, 768, 150), (515, 32, 300, 150)])
def test_news_cnn_bf16_backward(N, L, E, H, grouped, monkeypatch):
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import manager_for, rel_err
    import news_recommendation_mind_b200 as mr
    from news_recommendation_mind_b200 import ops
    monkeypatch.setattr(ops, "GROUPED_TABLE_GRAD", grouped)      # token-grouped table / filter gradient vs per-token d_emb + seg-reduce
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
    g = torch.randn(N, H, generator=gen)
    news = enc.encode_ids(emb, ids.cuda(), mask.cuda())
    (news * g.cuda()).sum().backward()
    torch.cuda.synchronize()
    exp = cnn_encoder_bf16_backward_emulation(emb.weight, ids, mask, enc.cnn.weight, enc.cnn.bias, enc.wordQueryProject.weight,
                                              enc.wordQueryProject.bias, enc.query_words, g, grouped=grouped)
    worst = 0.0
    for k, p in enc.named_parameters():
        e = rel_err(p.grad, exp[k])
        print("  %-28s %.3e" % (k, e))
        worst = max(worst, e)
    e = rel_err(emb.weight.grad, exp["table"])
    print("  %-28s %.3e" % ("table", e))
    worst = max(worst, e)
    assert float(emb.weight.grad[0].abs().max()) == 0.0          # padding_idx = 0 row (BERT.py:16-21)
    # same rounding points, so only accumulation order, tanh.approx and bf16 ties differ
    assert worst < 5e-3, worst


@pytest.mark.parametrize("kind,B,S,H,rev", [("lstm", 256, 50, 150, False), ("gru", 37, 20, 150, False), ("lstm", 5, 7, 64, True),
                                             ("gru", 300, 11, 96, True), ("lstm", 700, 9, 150, False)])
def test_rnn_user_encoder_resident_weights(kind, B, S, H, rev):
    """MR_BF16 recurrent user encoder (persistent kernel with W_hh resident in shared memory as bf16, input
    projection and weight-gradient GEMMs on the tensor-core kernels) against the oracle's LSTM / GRU
    (RNN.py:50-73) evaluated with the forward operand roundings the kernels apply (x, W_ih, W_hh to bf16;
    state, gates and accumulation fp32): outputs agree to 1e-4.  The backward GEMMs additionally round the
    gate gradients and h_{s-1} to bf16 (2^-9 relative per operand): gradients within 1e-2 relative L2."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import manager_for, rel_err
    import news_recommendation_mind_b200 as mr
    from oracle import twotower_oracle as O
    torch.manual_seed(B + S)
    man = manager_for("cnn", kind, 5, S, 32, 300, H, 10, precision="bf16", descend_history=rev)
    enc = mr.RNN_User_Encoder(man).cuda()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(B, S, H, generator=gen) * 0.5
    ln = torch.randint(1, S + 1, (B,), generator=gen)
    ln[0] = S
    his_mask = (torch.arange(S)[None, :] < ln[:, None]).double().unsqueeze(-1)
    g = torch.randn(B, 1, H, generator=gen)
    xc = x.cuda().requires_grad_(True)
    out = enc(xc, his_mask=his_mask)
    (out * g.cuda()).sum().backward()
    torch.cuda.synchronize()
    p = {k: v.detach().cpu().double().requires_grad_(True) for k, v in enc.rnn.named_parameters()}
    whh = p["weight_hh_l0"].detach().float().bfloat16().double().requires_grad_(True)
    wih = p["weight_ih_l0"].detach().float().bfloat16().double().requires_grad_(True)
    xo = x.bfloat16().double().requires_grad_(True)
    fn = O.lstm_user_encoder if kind == "lstm" else O.gru_user_encoder
    ref = fn(xo, his_mask, wih, whh, p["bias_ih_l0"], p["bias_hh_l0"], descend_history=rev)
    (ref * g.double()).sum().backward()
    errs = {"out": rel_err(out, ref), "d_x": rel_err(xc.grad, xo.grad),
            "d_w_ih": rel_err(enc.rnn.weight_ih_l0.grad, wih.grad),
            "d_w_hh": rel_err(enc.rnn.weight_hh_l0.grad, whh.grad),
            "d_b_ih": rel_err(enc.rnn.bias_ih_l0.grad, p["bias_ih_l0"].grad),
            "d_b_hh": rel_err(enc.rnn.bias_hh_l0.grad, p["bias_hh_l0"].grad)}
    print(kind, B, S, H, {k: "%.2e" % v for k, v in errs.items()})
    assert errs["out"] < 1e-4, errs
    assert max(errs.values()) < 1e-2, errs


@pytest.mark.parametrize("precision,gtol", [("fp32", 1e-4), ("bf16", 0.15)])
def test_dedup_titles_is_exact_in_forward_and_sums_gradients(precision, gtol):
    """In-batch unique-news dedup (opt-in) must give the same log-probabilities bit for bit (the encoder is batch
    invariant: a title's vector does not depend on where in the batch it sits) and the same gradients up to
    summation order (fp32: 1e-4; bf16 rounds the per-token gradients after the slot sum instead of before it, so
    the cancellation-heavy encoder gradients move by a few percent, as they do against the fp32 oracle)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import build_model, manager_for, rel_err
    from news_recommendation_mind_b200 import data
    torch.manual_seed(3)
    C, S, L, E, H, V = 5, 20, 32, 300, 150, 30522
    news_ids, news_mask = data.make_news_table(500, L, seed=2)
    x = data.make_train_batch(news_ids, news_mask, 16, C, S, seed=4)
    outs = []
    for dedup in (False, True):
        man = manager_for("cnn", "lstm", C, S, L, E, H, 10, precision=precision)
        man.dedup_titles = dedup
        torch.manual_seed(5)
        model = build_model(man, V)
        with torch.no_grad():
            model.embedding.weight.normal_(0, 0.3)
        model.train()
        logp = model(x)[0]
        torch.nn.NLLLoss()(logp, x["label"].cuda()).backward()
        outs.append((logp.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    assert torch.equal(outs[0][0], outs[1][0])
    for k in outs[0][1]:
        assert rel_err(outs[1][1][k], outs[0][1][k]) < gtol, (k, rel_err(outs[1][1][k], outs[0][1][k]))


def test_full_size_step_is_bit_reproducible_and_paths_agree(monkeypatch):
    """BASELINE config 1/2 at full size (B=256, C=5, S=50, L=32, E=300, H=150, V=30522; 450,560 tokens) through
    size-independent properties: (1) two runs of the same training step give bit-identical loss and gradients (no
    floating-point atomics anywhere: sorted segmented reductions, fixed-order partial sums); (2) the token-grouped
    backward and the per-token d_emb + segmented-reduction backward agree (same sums, different association and
    rounding point); (3) the padding row of the table gets no gradient (BERT.py:16-21); (4) checksum: the column sums
    of the table gradient of the two paths agree."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    sys.path.insert(0, os.path.dirname(os.path.dirname(__file__)))
    from helpers import rel_err
    import bench
    import news_recommendation_mind_b200 as mr
    from news_recommendation_mind_b200 import data, ops
    CFG = bench.CFG
    torch.manual_seed(42)
    man = bench.manager_ns("cuda:0", "bf16")
    model = mr.TwoTower(man, mr.BERT_Embedding(man, vocab_size=CFG["V"]), mr.CNN_Encoder(man), mr.RNN_User_Encoder(man)).to("cuda:0")
    ids, mask = data.make_news_table(CFG["n_news"], CFG["L"])
    x = {k: v.cuda() for k, v in data.make_train_batch(ids, mask, CFG["B"], CFG["C"], CFG["S"], seed=7).items()}

    def run():
        model.zero_grad(set_to_none=True)
        loss = ops.NLLMean.apply(model(x)[0], x["label"])
        loss.backward()
        torch.cuda.synchronize()
        return loss.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    monkeypatch.setattr(ops, "GROUPED_TABLE_GRAD", True)
    l1, g1 = run()
    l2, g2 = run()
    assert torch.equal(l1, l2)
    for k in g1:
        assert torch.equal(g1[k], g2[k]), k                       # (1) bit reproducible
    monkeypatch.setattr(ops, "GROUPED_TABLE_GRAD", False)
    l3, g3 = run()
    assert torch.equal(l1, l3)                                    # the forward is shared
    tab = "embedding.bert_word_embedding.weight"
    for k in g1:
        e = rel_err(g1[k], g3[k])
        print("  %-44s grouped vs per-token %.3e" % (k, e))
        assert e < (2e-2 if k in (tab, "encoderN.cnn.weight") else 1e-6), (k, e)     # (2) only the two regrouped gradients differ
    assert float(g1[tab][0].abs().max()) == 0.0 and float(g3[tab][0].abs().max()) == 0.0                  # (3)
    cs1, cs3 = g1[tab].double().sum(0), g3[tab].double().sum(0)
    assert float((cs1 - cs3).norm() / cs3.norm()) < 2e-2                                                  # (4)
    assert bool(torch.isfinite(g1[tab]).all())


def test_graph_step_matches_eager_steps():
    """trainer.GraphStep (the training step captured as one CUDA graph; Adam bias corrections, learning rates and grad_scale
    read from device memory) replays to bit-identical losses and parameters as eager train_step calls on the same batches,
    WITH a LinearWarmupSchedule moving the learning rates every step (the captured launch must not freeze them), and
    constructing it does not train: parameters, moments and step counter are as before."""
    import sys, os, copy
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import build_model, manager_for, random_batch
    from news_recommendation_mind_b200 import trainer
    B, C, S, L, E, H, V = 6, 5, 9, 32, 300, 150, 500
    gen = torch.Generator().manual_seed(4)
    batches = [{k: v.cuda() for k, v in random_batch(gen, B, C, S, L, V).items()} for _ in range(3)]
    torch.manual_seed(6)
    man = manager_for("cnn", "lstm", C, S, L, E, H, 10, precision="bf16")
    m1 = build_model(man, V)
    m2 = copy.deepcopy(m1)
    o1, o2 = trainer.FusedAdam(m1, lr=1e-3, bert_lr=1e-4), trainer.FusedAdam(m2, lr=1e-3, bert_lr=1e-4)
    s1 = trainer.LinearWarmupSchedule(o1, 3, 10)            # built FIRST: lr = 0 at capture time (the ADVICE r1 failure mode)
    s2 = trainer.LinearWarmupSchedule(o2, 3, 10)
    gs = trainer.GraphStep(m1, o1, batches[0])              # warm-up steps + capture, then everything restored
    assert o1.steps == 0
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), ("constructing a GraphStep must not train", k)
    o2.enable_device_step_scalars("cuda:0")                 # same arithmetic for the step scalars as the graph path
    for s in range(6):
        got = float(gs(batches[s % 3]))
        s1.step()
        o2.begin_step()
        exp = float(trainer.train_step(m2, batches[s % 3], o2))
        s2.step()
        assert got == exp, (s, got, exp)
    assert o1.steps == o2.steps == 6
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), k
    assert o1.param_groups[0]["lr"] != 1e-3                 # the schedule is live (6th step of a 3-warm-up / 10-total ramp)


def test_graph_step_learns_under_warmup_schedule():
    """With the schedule built before the capture the captured learning rate used to be 0 for ever: the parameters must
    change once the ramp has left 0."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import build_model, manager_for, random_batch
    from news_recommendation_mind_b200 import trainer
    B, C, S, L, E, H, V = 4, 5, 6, 32, 300, 150, 300
    gen = torch.Generator().manual_seed(1)
    x = {k: v.cuda() for k, v in random_batch(gen, B, C, S, L, V).items()}
    torch.manual_seed(2)
    m = build_model(manager_for("cnn", "lstm", C, S, L, E, H, 10, precision="bf16"), V)
    o = trainer.FusedAdam(m, lr=1e-2, bert_lr=1e-3)
    sched = trainer.LinearWarmupSchedule(o, 2, 8)
    gs = trainer.GraphStep(m, o, x)
    before = m.encoderN.cnn.weight.detach().clone()
    gs(x); sched.step()                                     # lr factor 0: nothing moves
    assert torch.equal(before, m.encoderN.cnn.weight)
    gs(x); sched.step()                                     # factor 1/2
    assert not torch.equal(before, m.encoderN.cnn.weight)

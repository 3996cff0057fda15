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

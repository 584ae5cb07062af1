"""tcgen05 / TMEM plumbing (supernormal_b200/csrc/umma.cuh): one CTA computes D[128,N] = A[128,K] B[N,K]^T with kind::tf32
from thread-written shared-memory operands and reads the accumulator back thread-per-row.  Compared with an fp64 matmul of
the TF32-rounded inputs (the products are then exact; only fp32 accumulation order differs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("K,N,lbo", [(8, 64, 128), (40, 64, 128), (64, 32, 128), (128, 64, 128), (32, 32, 128), (40, 64, 192), (64, 32, 192), (8, 64, 192)])
def test_umma_selftest(cuda, K, N, lbo):
    from supernormal_b200._lib import call, ptr
    g = torch.Generator(device=cuda).manual_seed(K * 100 + N)
    A = _tf32(torch.randn(128, K, device=cuda, generator=g))
    B = _tf32(torch.randn(N, K, device=cuda, generator=g))
    D = torch.full((128, N), float("nan"), device=cuda)
    err = torch.zeros(1, dtype=torch.int32, device=cuda)
    if lbo == 128:
        call("snb_umma_selftest", ptr(A), ptr(B), ptr(D), K, N, ptr(err))
    else:   # A tile with a padded leading-dimension byte offset (the layout of the split backward's point tiles)
        call("snb_umma_selftest_lbo", ptr(A), ptr(B), ptr(D), K, N, lbo, ptr(err))
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "tcgen05.mma never committed"
    ref = A.double() @ B.double().T
    assert torch.isfinite(D).all()
    assert (D.double() - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())

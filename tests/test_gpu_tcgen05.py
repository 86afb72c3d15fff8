"""GPU tests of the hand-written tcgen05 path: descriptor / TMEM-layout self-test and the fused field MLP."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _selftest(mode, A, B, M, N, K):
    from ngp_b200 import _cabi as c
    D = torch.full((M if mode == 1 else 128, N), float("nan"), device=DEV)
    c.call("ngp_tc_selftest", D.device, mode, c.ptr(A), c.ptr(B), c.ptr(D), M, N, K)
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("N,K", [(64, 32), (64, 64), (16, 64), (32, 16), (128, 128)])
def test_umma_k_major(N, K):
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).half().to(DEV)
    B = torch.randn(N, K, generator=g).half().to(DEV)
    D = _selftest(0, A, B, 128, N, K)
    want = A.float() @ B.float().T
    assert torch.allclose(D, want, rtol=1e-3, atol=1e-3), (D - want).abs().max().item()


@pytest.mark.parametrize("M,N", [(64, 16), (64, 64), (64, 32), (128, 64)])
def test_umma_mn_major_weight_gradient_shape(M, N):
    g = torch.Generator().manual_seed(M * 1000 + N)
    A = torch.randn(128, M, generator=g).half().to(DEV)    # [samples, M]
    B = torch.randn(128, N, generator=g).half().to(DEV)    # [samples, N]
    D = _selftest(1, A, B, M, N, 128)
    want = A.float().T @ B.float()
    assert torch.allclose(D, want, rtol=1e-3, atol=2e-3), (D - want).abs().max().item()


@pytest.mark.parametrize("N,K", [(64, 16), (64, 64), (32, 64)])
def test_umma_data_gradient_shape(N, K):
    g = torch.Generator().manual_seed(N * 77 + K)
    A = torch.randn(128, K, generator=g).half().to(DEV)    # upstream grads [samples, out]
    B = torch.randn(K, N, generator=g).half().to(DEV)      # weights [out, in]
    D = _selftest(2, A, B, 128, N, K)
    want = A.float() @ B.float()
    assert torch.allclose(D, want, rtol=1e-3, atol=2e-3), (D - want).abs().max().item()

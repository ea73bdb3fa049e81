"""GPU parity of the time-batched contractions (input projection, dX, weight gradients) through the C ABI
against torch matmul in fp64.  fp32 path: 1e-4 normwise (north_star); bf16 projection mode: 2e-2."""
import pytest
import torch

from parity_util import relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("M,N,K", [(37, 72, 14), (24576, 72, 24), (1000, 192, 64), (513, 384, 128), (300, 14, 56)])
def test_proj_dgrad_wgrad_fp32(M, N, K):
    from timegan_b200 import ops, _lib
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    dG = torch.randn(M, N, generator=g)
    Ad, Wd, bd, dGd = A.to(DEV), W.to(DEV), b.to(DEV), dG.to(DEV)
    C = ops.proj(Ad, Wd, bd, mode=_lib.PROJ_FP32)
    assert relerr(C, A.double() @ W.double().T + b.double()) < 1e-5
    dX = ops.dgrad(dGd, Wd)
    assert relerr(dX, dG.double() @ W.double()) < 1e-5
    dW = torch.empty(N, K, device=DEV); db = torch.empty(N, device=DEV)
    ops.wgrad(dGd, Ad, dW, db, N)
    assert relerr(dW, dG.double().T @ A.double()) < 1e-5
    assert relerr(db, dG.double().sum(0)) < 1e-5
    # accumulate
    ops.wgrad(dGd, Ad, dW, db, N, accumulate=True)
    assert relerr(dW, 2 * (dG.double().T @ A.double())) < 1e-5


def test_wgrad_shifted_rows():
    """dW_hh = sum_t dGH_t^T h_{t-1} with h_{-1} = 0, reading h_{t-1} from the layer output."""
    from timegan_b200 import ops
    B, T, H = 5, 33, 24
    g = torch.Generator().manual_seed(0)
    y = torch.randn(B, T, H, generator=g)
    dg = torch.randn(B, T, 3 * H, generator=g)
    hprev = torch.cat([torch.zeros(B, 1, H), y[:, :-1]], 1)
    ref = dg.double().reshape(-1, 3 * H).T @ hprev.double().reshape(-1, H)
    dW = torch.empty(3 * H, H, device=DEV)
    ops.wgrad(dg.to(DEV).view(B * T, 3 * H), y.to(DEV).view(B * T, H), dW, None, 3 * H, shift_T=T)
    assert relerr(dW, ref) < 1e-5


@pytest.mark.parametrize("M,N,K", [(24576, 72, 24), (4096, 192, 64), (1000, 384, 128), (777, 72, 14)])
def test_proj_bf16_mode(M, N, K):
    from timegan_b200 import ops, _lib
    g = torch.Generator().manual_seed(M + 1)
    A = torch.rand(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    C = ops.proj(A.to(DEV), W.to(DEV), b.to(DEV), mode=_lib.PROJ_BF16)
    assert relerr(C, A.double() @ W.double().T + b.double()) < 2e-2

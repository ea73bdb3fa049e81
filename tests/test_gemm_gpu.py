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


TC_SHAPES = [(24576, 72, 24), (4096, 192, 64), (1000, 384, 128), (777, 72, 14), (128, 16, 4), (5000, 168, 28),
             (3000, 64, 192), (1300, 24, 72), (196608, 192, 64)]


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
@pytest.mark.parametrize("mode,tol", [("tf32", 2e-2), ("tf32x3", 2e-6)])
def test_proj_tensor_core_modes(M, N, K, mode, tol):
    """tcgen05 projection: one TF32 pass (the 2e-2 'bf16-class' mode; measured ~5e-4) and 3xTF32 (fp32 parity).
    Shapes the tile cannot take (K % 4 != 0) fall back to the FFMA kernel inside tg_proj and still pass."""
    from timegan_b200 import ops, _lib
    g = torch.Generator().manual_seed(M + 1)
    A = torch.rand(M, K, generator=g) * 2 - 0.7
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = A.double() @ W.double().T + b.double()
    C = ops.proj(A.to(DEV), W.to(DEV), b.to(DEV), mode=ops._MODES[mode])
    err = relerr(C, ref)
    assert err < tol, err
    assert err < 2e-3          # even the single pass is far inside its 2e-2 budget
    # accumulate and no-bias variants
    C2 = ops.proj(A.to(DEV), W.to(DEV), None, out2d=C.clone(), mode=ops._MODES[mode], accumulate=True)
    assert relerr(C2, 2 * ref - b.double()) < max(tol, 2e-6) * 2


@pytest.mark.parametrize("M,N,K", [(4096, 64, 192), (2000, 24, 72), (1500, 28, 168)])
def test_dgrad_through_tensor_cores(M, N, K):
    from timegan_b200 import ops
    g = torch.Generator().manual_seed(K)
    dG = torch.randn(M, K, generator=g)
    W = torch.randn(K, N, generator=g) / K ** 0.5
    ref = dG.double() @ W.double()
    old = ops.get_proj_mode()
    try:
        ops.set_proj_mode("tf32x3")
        assert relerr(ops.dgrad(dG.to(DEV), W.to(DEV)), ref) < 2e-6
        ops.set_proj_mode("tf32")
        assert relerr(ops.dgrad(dG.to(DEV), W.to(DEV)), ref) < 2e-3
    finally:
        ops.set_proj_mode(old)


@pytest.mark.parametrize("M,N,K,T", [(24576, 72, 24, 768), (6000, 192, 64, 200), (4096, 128, 64, 0), (3000, 168, 56, 100),
                                     (196608, 192, 64, 768), (5000, 64, 64, 50), (2048, 384, 128, 64), (1000, 24, 24, 0)])
@pytest.mark.parametrize("mode,tol", [("tf32", 3e-3), ("tf32x3", 2e-5)])
def test_wgrad_tensor_core_modes(M, N, K, T, mode, tol):
    """tcgen05 weight gradients (MN-major operands) incl. the h_{t-1} row shift and the fused bias gradient,
    from a strided column slice of a wider dG (as the BPTT kernels' outputs are consumed)."""
    from timegan_b200 import ops
    g = torch.Generator().manual_seed(N + K)
    ldg = N + 64
    dG_full = torch.randn(M, ldg, generator=g)
    A = torch.rand(M, K, generator=g) - 0.3
    dG = dG_full[:, :N]
    if T > 0:
        B = M // T
        Ash = torch.cat([torch.zeros(B, 1, K), A.view(B, T, K)[:, :-1]], 1).reshape(M, K)
    else:
        Ash = A
    ref_w = dG.double().T @ Ash.double()
    ref_b = dG.double().sum(0)
    dGd = dG_full.to(DEV)[:, :N]
    dW = torch.full((N, K), 7.0, device=DEV)
    db = torch.full((N,), 7.0, device=DEV)
    ops.wgrad(dGd, A.to(DEV), dW, db, N, shift_T=T, mode=ops._MODES[mode])
    assert relerr(dW, ref_w) < tol, relerr(dW, ref_w)
    assert relerr(db, ref_b) < 1e-5
    ops.wgrad(dGd, A.to(DEV), dW, None, N, shift_T=T, accumulate=True, mode=ops._MODES[mode])
    assert relerr(dW, 2 * ref_w) < tol
    assert relerr(db, ref_b) < 1e-5      # untouched when db is not requested


@pytest.mark.parametrize("M,N,K", [(1000, 192, 64), (1061, 384, 128), (777, 192, 16), (4096, 384, 64), (128, 192, 128),
                                   (196608, 192, 64), (113664, 384, 128)])
def test_proj_bf16_operands_and_bf16_result(M, N, K):
    """tg_proj_bf16 (the 'bf16 input projections' of BASELINE config c3): fp32 activations converted to bf16 inside the
    kernel, bf16 W, fp32 accumulation, bf16 result.  Against the same contraction of the bf16-ROUNDED operands in fp64,
    rounded to bf16: at most one bf16 ulp apart (accumulation order), i.e. <= 2^-7 relative per element."""
    from timegan_b200._lib import lib, check, ptr, stream_ptr
    g = torch.Generator().manual_seed(M + K)
    A = torch.rand(M, K, generator=g) * 2 - 0.7
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    Ad, bd = A.to(DEV), b.to(DEV)
    W16 = W.to(DEV).to(torch.bfloat16)
    C16 = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    check(lib.tg_proj_bf16(stream_ptr(), ptr(Ad), K, ptr(W16), K, ptr(bd), ptr(C16), N, M, N, K), "tg_proj_bf16")
    ref = A.to(torch.bfloat16).double() @ W.to(torch.bfloat16).double().T + b.double()
    got = C16.float().cpu().double()
    assert torch.isfinite(got).all()
    err = (got - ref).abs() / (ref.abs() + 1e-2)
    assert err.max().item() < 2.0 ** -7, err.max().item()
    assert relerr(got, ref) < 3e-3
    # without bias, into a wider output (ldc > N is not used by the step, so only the plain layout is exercised)
    C2 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    check(lib.tg_proj_bf16(stream_ptr(), ptr(Ad), K, ptr(W16), K, None, ptr(C2), N, M, N, K), "tg_proj_bf16")
    assert relerr(C2.float().cpu().double(), ref - b.double()) < 3e-3

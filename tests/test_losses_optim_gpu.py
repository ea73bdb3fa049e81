"""GPU parity of the fused loss / clip+Adam / noise kernels against the reference's torch expressions
(train_timegan.py:70-126 losses, :141-142 clip+Adam, :40-47,64-65 noise).  Tolerance 1e-4 normwise."""
import math

import pytest
import torch

from parity_util import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


# ---- the reference's loss expressions, restated in plain torch (tt:72-74, 79-80, 82-126) -----------------
def ref_recon(x, xt, eps=1e-8):
    return 10.0 * torch.sqrt(torch.mean((x - xt) ** 2) + eps)


def ref_diff1(h):
    return torch.mean((h[:, 1:, :] - h[:, :-1, :]) ** 2)


def ref_cov(x):
    B, T, C = x.shape
    X = x.reshape(B * T, C)
    X = X - X.mean(dim=0, keepdim=True)
    return (X.t() @ X) / (X.size(0) - 1)


def ref_cov_term(xg, xr):
    cr = ref_cov(xr.detach())
    return torch.norm(ref_cov(xg) - cr, p="fro") / (cr.numel() ** 0.5)


def ref_acf_term(xg, xr, max_lag):
    B, T, C = xg.shape
    max_lag = max(1, min(max_lag, T - 1))

    def acf_all(x):
        xz = (x - x.mean(dim=(0, 1), keepdim=True)) / (x.std(dim=(0, 1), keepdim=True) + 1e-8)
        return torch.stack([(xz[:, :-l, :] * xz[:, l:, :]).mean(dim=(0, 1)) for l in range(1, max_lag + 1)], 0)

    return torch.mean(torch.abs(acf_all(xg) - acf_all(xr).detach()))


@pytest.mark.parametrize("shape", [(3, 17, 14), (32, 768, 14), (5, 100, 7)])
def test_recon_loss(shape):
    from timegan_b200 import losses
    g = torch.Generator().manual_seed(0)
    x = torch.rand(shape, generator=g)
    xt = torch.rand(shape, generator=g).requires_grad_(True)
    ref = ref_recon(x, xt)
    ref.backward()
    xo = xt.detach().to(DEV).requires_grad_(True)
    out = losses.recon_loss(x.to(DEV), xo)
    out.backward()
    assert abs(out.item() - ref.item()) <= TOL * abs(ref.item())
    assert relerr(xo.grad, xt.grad) < TOL


@pytest.mark.parametrize("shape", [(3, 17, 24), (8, 767, 24)])
def test_mse_and_first_difference(shape):
    from timegan_b200 import losses
    g = torch.Generator().manual_seed(1)
    a = torch.rand(shape, generator=g).requires_grad_(True)
    b = torch.rand(shape, generator=g)
    ref = torch.mean((a - b) ** 2)
    ref.backward()
    ao = a.detach().to(DEV).requires_grad_(True)
    out = losses.mse_loss(ao, b.to(DEV))
    out.backward()
    assert abs(out.item() - ref.item()) <= TOL * abs(ref.item())
    assert relerr(ao.grad, a.grad) < TOL
    h = torch.rand(shape, generator=g).requires_grad_(True)
    r2 = ref_diff1(h)
    r2.backward()
    ho = h.detach().to(DEV).requires_grad_(True)
    o2 = losses.sup_loss_fake(ho)
    o2.backward()
    assert abs(o2.item() - r2.item()) <= TOL * abs(r2.item())
    assert relerr(ho.grad, h.grad) < TOL


def test_bce_matches_nn_bceloss():
    from timegan_b200 import losses
    g = torch.Generator().manual_seed(2)
    p = torch.rand(32, 1, generator=g).requires_grad_(True)
    y = torch.rand(32, 1, generator=g)
    ref = torch.nn.BCELoss()(p, y)
    ref.backward()
    po = p.detach().to(DEV).requires_grad_(True)
    out = losses.bce(po, y.to(DEV))
    out.backward()
    assert abs(out.item() - ref.item()) <= 1e-6
    assert relerr(po.grad, p.grad) < 1e-5
    # saturated probabilities: log clamp at -100 like ATen
    ps = torch.tensor([[0.0], [1.0]])
    ys = torch.tensor([[1.0], [0.0]])
    assert abs(losses.bce(ps.to(DEV), ys.to(DEV)).item() - torch.nn.BCELoss()(ps, ys).item()) < 1e-4


@pytest.mark.parametrize("B,T,C,L", [(3, 40, 14, 8), (32, 768, 14, 64), (4, 100, 14, 200)])
def test_cov_and_acf_terms(B, T, C, L):
    from timegan_b200 import losses
    g = torch.Generator().manual_seed(3)
    # autocorrelated series so the ACF term is not ~0
    xr = torch.cumsum(torch.randn(B, T, C, generator=g), 1) * 0.05 + 0.5
    xg = (torch.cumsum(torch.randn(B, T, C, generator=g), 1) * 0.03 + 0.4).requires_grad_(True)
    cov_ref = ref_cov_term(xg, xr)
    acf_ref = ref_acf_term(xg, xr, L)
    (0.7 * cov_ref + 1.3 * acf_ref).backward()
    xo = xg.detach().to(DEV).requires_grad_(True)
    cov, acf = losses.cov_acf_losses(xo, xr.to(DEV), L)
    (0.7 * cov + 1.3 * acf).backward()
    assert abs(cov.item() - cov_ref.item()) <= TOL * abs(cov_ref.item())
    assert abs(acf.item() - acf_ref.item()) <= TOL * abs(acf_ref.item())
    assert relerr(xo.grad, xg.grad) < 5e-4   # sign(|.|) kinks + fp32 z-scoring: slightly looser than 1e-4


@pytest.mark.parametrize("clip", [0.5, 0.0, 1e6])
def test_clip_adam_matches_torch(clip):
    """clip_grad_norm_ + Adam(betas=(0.5,0.9)) over several steps (tt:141-142, 331)."""
    from timegan_b200 import FusedAdam
    g = torch.Generator().manual_seed(4)
    shapes = [(72, 14), (72, 24), (72,), (72,), (14, 24), (14,), (5000,)]
    ps_ref = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    ps = [p.detach().clone().to(DEV).requires_grad_(True) for p in ps_ref]
    o_ref = torch.optim.Adam(ps_ref, lr=1e-3, betas=(0.5, 0.9))
    o = FusedAdam(ps, lr=1e-3, betas=(0.5, 0.9))
    for it in range(5):
        for a, b in zip(ps_ref, ps):
            gr = torch.randn(a.shape, generator=g) * (10.0 if it % 2 == 0 else 0.01)
            a.grad = gr.clone()
            b.grad = gr.clone().to(DEV)
        if clip > 0:
            torch.nn.utils.clip_grad_norm_(ps_ref, clip)
        o_ref.step()
        o.clip_and_step(clip)
    for a, b in zip(ps_ref, ps):
        assert relerr(b, a) < 1e-6
    sd = o.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert relerr(sd["state"][0]["exp_avg"], o_ref.state_dict()["state"][0]["exp_avg"]) < 1e-5


def test_noise_kernels_statistics_and_reproducibility():
    from timegan_b200 import noise
    n = 1 << 20
    u = noise.uniform((n,), DEV, seed=1234, offset=0)
    u2 = noise.uniform((n,), DEV, seed=1234, offset=0)
    u3 = noise.uniform((n,), DEV, seed=1234, offset=n)
    assert torch.equal(u, u2) and not torch.equal(u, u3)
    assert 0.0 <= u.min().item() and u.max().item() < 1.0
    assert abs(u.mean().item() - 0.5) < 2e-3 and abs(u.var().item() - 1 / 12) < 1e-3
    base = torch.full((n,), 2.0, device=DEV)
    z = noise.add_normal(base, 0.3, seed=99, offset=0)
    assert abs(z.mean().item() - 2.0) < 2e-3 and abs(z.std().item() - 0.3) < 2e-3
    k = ((z - 2.0) / 0.3)
    assert abs((k ** 4).mean().item() - 3.0) < 0.05          # Gaussian kurtosis
    assert noise.add_normal(base, 0.0, seed=1, offset=0) is base  # tt:46-47: std <= 0 returns the input itself

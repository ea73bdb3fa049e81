"""Parity AT THE BENCHMARKED SHAPES (BASELINE.json configs c2 and c3), not at scaled-down stand-ins.

bench.py times z = h = 64, L = 3, B = 256, T = 768, C = 14 (c2) and reports c3 (h = 128, reduced-precision
projections).  These tests run the same kernels' instantiations the bench runs -- one sequence per CTA at B = 256, the
148-CTA split-M fused weight gradient at M = 196 608, the 2B-batch discriminator pass with dy_last -- against
torch.nn.GRU + autograd on the CPU (the reference's arithmetic path, timegan_model.py:32-34), fp64 contractions, and
the CPU port of the reference's step functions (oracle/timegan_ref.py; tt:166-276).

Tolerance: fp32 forward and gradients within 1e-4 normwise relative; reduced-precision projection mode within 2e-2
(BASELINE.json north_star)."""
import pytest
import torch

from parity_util import relerr, make_gru, flat_weights

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"
T_LEN = 768


def _stack_parity(B, I, H, L, dy_last, seed):
    from timegan_b200 import ops
    m = make_gru(I, H, L, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, T_LEN, I, generator=g, requires_grad=True)
    y_ref, _ = m(x)
    if dy_last:
        dy = torch.randn(B, H, generator=g)
        obj = (y_ref[:, -1] * dy).sum()
    else:
        dy = torch.randn(B, T_LEN, H, generator=g)
        obj = (y_ref * dy).sum()
    ref = torch.autograd.grad(obj, [x] + list(m.parameters()))
    w = flat_weights(m, DEV)
    y, saves = ops.stack_forward(x.detach().to(DEV), w, save=True)
    dx, grads = ops.stack_backward(dy.to(DEV), saves, w, need_dx=True, need_dw=True, dy_last=dy_last)
    assert relerr(y, y_ref) < TOL, ("y", relerr(y, y_ref))
    assert relerr(dx, ref[0]) < TOL, ("dx", relerr(dx, ref[0]))
    for k, gk in enumerate(grads):
        assert relerr(gk, ref[1 + k]) < TOL, (f"param {k}", relerr(gk, ref[1 + k]))


@pytest.mark.parametrize("name,B,I,H,dy_last", [
    ("c2 embedder", 256, 14, 64, False),          # E: 14 -> 64, full-sequence gradient
    ("c2 latent stack", 256, 64, 64, False),      # G / S / R: 64 -> 64
    ("c2 discriminator", 512, 64, 64, True),      # D over [real ; fake] (2B), gradient seeded at t = T-1 only
    ("c3 latent stack", 256, 128, 128, False),    # c3: h = 128, B = 256 per GPU
    ("c3 discriminator", 296, 128, 128, True),    # B = 2 x 148
])
def test_stack_forward_and_bptt_at_benchmark_shape(name, B, I, H, dy_last):
    """nn.GRU forward + autograd vs the persistent kernels at exactly (B, 768, I -> H, L = 3)."""
    _stack_parity(B, I, H, 3, dy_last, seed=B + H)


@pytest.mark.parametrize("H,I,B", [(64, 64, 256), (64, 14, 256), (128, 128, 256), (256, 256, 40), (256, 16, 40)])
def test_fused_weight_gradient_at_m_196608(H, I, B):
    """tg_wgrad_gru (one launch per layer, split-M over every SM) at M = B*T = 196 608 against fp64 contractions:
    dW_ih = dGI^T x, dW_hh = [dGI_r, dGI_z, dq]^T h_{t-1}, db_ih = colsum dGI, db_hh = colsum [dGI_r, dGI_z, dq]."""
    from timegan_b200 import ops
    T = T_LEN                      # B = 256 -> M = 196 608; H = 256 (8 launches: 4 M-tile ranges x 2 column ranges) at M = 30 720
    g = torch.Generator().manual_seed(H + I)
    dgi = torch.randn(B, T, 3 * H, generator=g)
    dq = torch.randn(B, T, H, generator=g)
    x = torch.rand(B, T, I, generator=g)
    y = torch.rand(B, T, H, generator=g) * 2 - 1
    dgh = torch.cat([dgi[..., :2 * H], dq], -1).double()
    hprev = torch.cat([torch.zeros(B, 1, H), y[:, :-1]], 1).double()
    ref_wih = dgi.double().reshape(-1, 3 * H).t() @ x.double().reshape(-1, I)
    ref_whh = dgh.reshape(-1, 3 * H).t() @ hprev.reshape(-1, H)
    ref_bih = dgi.double().sum((0, 1))
    ref_bhh = dgh.sum((0, 1))
    out = [torch.empty(3 * H, I, device=DEV), torch.empty(3 * H, H, device=DEV), torch.empty(3 * H, device=DEV),
           torch.empty(3 * H, device=DEV)]
    ops.wgrad_gru(dgi.to(DEV), dq.to(DEV), x.to(DEV).view(B * T, I), y.to(DEV), *out, accumulate=False)
    for got, ref, nm in zip(out, (ref_wih, ref_whh, ref_bih, ref_bhh), ("dW_ih", "dW_hh", "db_ih", "db_hh")):
        assert relerr(got, ref) < 2e-5, (nm, relerr(got, ref))
    # accumulate=True adds a second contribution on top (the tangent path of R1 uses it)
    ops.wgrad_gru(dgi.to(DEV), dq.to(DEV), x.to(DEV).view(B * T, I), y.to(DEV), *out, accumulate=True)
    assert relerr(out[0], 2 * ref_wih) < 2e-5 and relerr(out[1], 2 * ref_whh) < 2e-5


class _PreClipGrads:
    """Records every parameter's gradient as the port hands it to clip_grad_norm_ (i.e. BEFORE clipping rescales
    .grad in place) -- the product leaves .grad unclipped and folds the clip coefficient into its Adam kernel."""

    def __init__(self, R):
        self.R, self.calls = R, []

    def __enter__(self):
        self.orig = self.R.clip_grad_norm_

        def rec(params, max_norm, *a, **kw):
            params = list(params)
            self.calls.append({id(p): p.grad.detach().clone() for p in params if p.grad is not None})
            return self.orig(params, max_norm, *a, **kw)
        self.R.clip_grad_norm_ = rec
        return self

    def __exit__(self, *a):
        self.R.clip_grad_norm_ = self.orig


def _joint_step_vs_port(z, h, B, proj_mode, tol):
    import timegan_b200 as tg
    from timegan_b200 import ops, train_timegan as tt
    from oracle import timegan_ref as R
    torch.manual_seed(11)
    port = R.build_model(14, z, h, 3, 0.0)
    ours = tg.TimeGAN(14, z, h, 3, 0.0)
    ours.load_state_dict(port.state_dict())
    ours = ours.to(DEV)
    x = torch.rand(B, T_LEN, 14)
    xd = x.to(DEV)
    op = R.make_optimizers(port)
    P = tt._params
    oD = tg.FusedAdam(ours.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9))
    oG = tg.FusedAdam(P(ours.generator, ours.supervisor, ours.embedder, ours.recovery), lr=1e-3, betas=(0.5, 0.9))
    old = ops.get_proj_mode()
    try:
        ops.set_proj_mode(proj_mode)
        torch.manual_seed(2024)
        with _PreClipGrads(R) as rec_d:
            d_p = R.d_step(port, x, op["D"], R.TorchNoise(), 0.2, 0.3, 0.5, 1.0, 0.525, 0.15)
        gd_p = {n: rec_d.calls[0][id(p)] for n, p in port["discriminator"].named_parameters()}
        torch.manual_seed(2024)
        nz = tt.HostReplayNoise(DEV)
        d_o = tt.disc_step(ours, xd, DEV, oD, 0.2, 0.3, 0.5, None, 1.0, target_acc=0.525, band=0.15, noise=nz)
        gd_o = {n: p.grad.detach().clone() for n, p in ours.discriminator.named_parameters()}
        state = torch.get_rng_state()
        with _PreClipGrads(R) as rec_g:
            g_p = R.g_step(port, x, op["G"], R.TorchNoise(), 5.0, 0.2, 0.3, 0.5, 0.05, 0.05, 64)
        torch.set_rng_state(state)
        g_o = tt.gen_step(ours, xd, DEV, oG, 5.0, 0.2, 0.3, 0.5, None, 0.05, 0.05, 64, noise=nz)
    finally:
        ops.set_proj_mode(old)
    close = lambda a, b: abs(a - b) <= tol * max(abs(b), 1e-3)
    assert close(d_o[0], d_p[0]) and abs(d_o[1] - d_p[1]) <= 1.0 / B + 1e-6, (d_o, d_p)
    for a, b, nm in zip(g_o, g_p, ("total", "adv", "sup", "rec", "cov", "acf")):
        assert close(a, b), (nm, g_o, g_p)
    for n, ref in gd_p.items():
        assert relerr(gd_o[n], ref) < tol, ("D grad", n, relerr(gd_o[n], ref))
    for mod in ("generator", "supervisor", "embedder", "recovery"):
        for (n, p), (_, q) in zip(getattr(ours, mod).named_parameters(), port[mod].named_parameters()):
            ref = rec_g.calls[0].get(id(q))
            if ref is None:
                continue
            assert relerr(p.grad, ref) < tol, (mod, n, relerr(p.grad, ref))


def test_joint_step_matches_port_at_c2():
    """One full disc_step + gen_step (R1, throttle, cov, ACF on) at EXACTLY the bench workload: z = h = 64, L = 3,
    B = 256, T = 768, C = 14 -- losses within 1e-4, every pre-clip gradient normwise (tt:166-276)."""
    _joint_step_vs_port(64, 64, 256, "fp32", TOL)


def test_joint_step_matches_port_at_c3_fp32():
    """c3 dims (h = 128), one sequence per SM (B = 148), fp32-parity projections."""
    _joint_step_vs_port(128, 128, 148, "fp32", TOL)


def test_joint_step_matches_port_at_c3_bf16_input_projections():
    """c3 as BASELINE.json states it: h = 128 with bf16 input projections (bf16 operands on the tensor pipe, bf16 gi
    read by the recurrence; dX and weight gradients one TF32 pass), losses and every pre-clip gradient within 2e-2."""
    _joint_step_vs_port(128, 128, 148, "bf16", 2e-2)


def test_joint_step_matches_port_at_c3_tf32_projections():
    """The other reduced-precision mode: one TF32 tensor-core pass over the fp32 operands everywhere, fp32 gi."""
    _joint_step_vs_port(128, 128, 148, "tf32", 2e-2)

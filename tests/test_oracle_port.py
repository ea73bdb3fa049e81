"""Pins oracle/timegan_ref.py (the CPU restatement that stands in for the reference on the GPU box):
  * against tests/golden/steps_*.npz, i.e. outputs of the UNMODIFIED reference (always);
  * against the unmodified reference imported live from /root/reference/timeGAN (build container only).
Also checks that the product's model constructor consumes the RNG like the reference's (same initial weights)."""
import sys
from pathlib import Path

import pytest
import torch

from golden_util import CASES, HP, StepFixture
from oracle import timegan_ref as R
from parity_util import relerr

REF = Path("/root/reference/timeGAN")


def _run_port(fx, model):
    opts = R.make_optimizers(model, HP["lr_g"], HP["lr_d"], HP["betas"])
    grads = {}
    snap = lambda names: {n: p.grad.detach().clone() for n, p in model.named_parameters()
                          if p.grad is not None and n.split(".")[0] in names}
    # capture pre-clip gradients by clipping with an infinite threshold first is not possible after the fact,
    # so wrap the clip call like the fixture generator does
    import torch.nn.utils as U
    cap = []
    orig = R.clip_grad_norm_
    def rec(params, max_norm, *a, **k):
        params = list(params)
        cap.append({id(p): p.grad.detach().clone() for p in params if p.grad is not None})
        return orig(params, max_norm, *a, **k)
    R.clip_grad_norm_ = rec
    try:
        ae = R.ae_step(model, fx.x, opts["ER"], HP["clip"])
        sup = R.sup_step(model, fx.x, opts["S"], HP["clip"])
        torch.manual_seed(fx.seed + 1)
        nz = R.TorchNoise()
        d = R.d_step(model, fx.x, opts["D"], nz, HP["label_smooth"], HP["inst_noise"], HP["clip"], HP["r1_gamma"],
                     HP["target"], HP["band"])
        g = R.g_step(model, fx.x, opts["G"], nz, HP["alpha_sup"], HP["beta_rec"], HP["inst_noise"], HP["clip"],
                     HP["gamma_cov"], HP["gamma_acf"], HP["acf_max_lag"])
    finally:
        R.clip_grad_norm_ = orig
    names = {id(p): n for n, p in model.named_parameters()}
    for step, c in zip(("ae", "sup", "d", "g"), cap):
        grads[step] = {names[i]: t for i, t in c.items()}
    return float(ae), float(sup), d, g, grads


@pytest.mark.parametrize("case", CASES)
def test_port_reproduces_reference_golden_steps(case):
    fx = StepFixture(case)
    torch.set_num_threads(1)
    model = R.build_model(fx.x_dim, fx.z_dim, fx.h_dim, fx.layers, 0.0)
    model.load_state_dict(fx.init)
    ae, sup, d, g, grads = _run_port(fx, model)
    assert abs(ae - fx.ae_loss) < 1e-4 and abs(sup - fx.sup_loss) < 1e-4   # fixtures hold the 5-decimal log print
    assert d == pytest.approx(fx.d_out, rel=1e-5, abs=1e-7)
    assert list(g) == pytest.approx(fx.g_out, rel=1e-5, abs=1e-7)
    for step in ("ae", "sup", "d", "g"):
        assert set(grads[step]) == set(fx.grads[step]), step
        for k, ref in fx.grads[step].items():
            assert relerr(grads[step][k], ref) < 1e-5, (step, k)
    sd = model.state_dict()
    for k, ref in fx.final.items():
        assert relerr(sd[k], ref) < 1e-5 or (sd[k] - ref).abs().max() < 1e-6, k


@pytest.mark.parametrize("dims", [(14, 24, 24, 3), (14, 28, 56, 1), (5, 6, 8, 2)])
def test_product_model_initialises_like_the_port(dims):
    """timegan_b200.TimeGAN(seed) == reference TimeGAN(seed): same keys, same RNG consumption, same values."""
    tg = pytest.importorskip("timegan_b200")
    x_dim, z, h, L = dims
    torch.manual_seed(123)
    a = R.build_model(x_dim, z, h, L, 0.0)
    ra = torch.rand(3)
    torch.manual_seed(123)
    b = tg.TimeGAN(x_dim, z, h, L, 0.0)
    rb = torch.rand(3)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert torch.equal(ra, rb)


@pytest.mark.skipif(not REF.exists(), reason="the unmodified reference is only present in the build container")
def test_port_matches_live_reference_bitwise():
    sys.path.insert(0, str(REF))
    try:
        import timegan_model as tm
        import train_timegan as tt
    finally:
        sys.path.remove(str(REF))
    torch.set_num_threads(1)
    x_dim, z, h, L, B, T = 7, 8, 12, 2, 3, 20
    torch.manual_seed(5)
    ref = tm.TimeGAN(x_dim, z, h, L, 0.0)
    torch.manual_seed(5)
    port = R.build_model(x_dim, z, h, L, 0.0)
    assert all(torch.equal(a, b) for a, b in zip(ref.state_dict().values(), port.state_dict().values()))
    x = torch.rand(B, T, x_dim)
    cpu = torch.device("cpu")
    pl = lambda *m: [p for mm in m for p in mm.parameters()]
    A = torch.optim.Adam
    oD_r = A(ref.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9))
    oG_r = A(pl(ref.generator, ref.supervisor, ref.embedder, ref.recovery), lr=1e-3, betas=(0.5, 0.9))
    op = R.make_optimizers(port)
    for it in range(2):
        torch.manual_seed(100 + it)
        d_r = tt.disc_step(ref, x, cpu, oD_r, 0.2, 0.3, 0.5, None, 1.0, target_acc=0.525, band=0.15)
        g_r = tt.gen_step(ref, x, cpu, oG_r, 5.0, 0.2, 0.3, 0.5, None, 0.05, 0.05, 64)
        torch.manual_seed(100 + it)
        nz = R.TorchNoise()
        d_p = R.d_step(port, x, op["D"], nz, 0.2, 0.3, 0.5, 1.0, 0.525, 0.15)
        g_p = R.g_step(port, x, op["G"], nz, 5.0, 0.2, 0.3, 0.5, 0.05, 0.05, 64)
        assert d_r == d_p and g_r == g_p
    for a, b in zip(ref.state_dict().values(), port.state_dict().values()):
        assert torch.equal(a, b)
    # generation chain (generate_long_synth.py:117-121)
    zt = torch.rand(2, 9, z)
    ref.eval(); port.eval()
    with torch.no_grad():
        assert torch.equal(ref.decode(ref.refine_latent(ref.gen_latent(zt))), R.generate(port, zt))

"""GPU parity of the four optimiser steps (phase_autoencoder, phase_supervisor, disc_step incl. R1 + throttle,
gen_step incl. cov/ACF) through the product API, against
  * tests/golden/steps_*.npz -- outputs of the UNMODIFIED reference (oracle/make_golden_steps.py), and
  * oracle/timegan_ref.py run live on the CPU at other shapes / seeds / hyper-parameters.
Tolerance: losses and pre-clip gradients within 1e-4 normwise relative (north_star, fp32 mode)."""
import pytest
import torch

from golden_util import CASES, HP, StepFixture
from parity_util import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"


def _grads(model, prefixes):
    return {n: p.grad.detach().clone() for n, p in model.named_parameters()
            if p.grad is not None and n.split(".")[0] in prefixes}


def _close(a, b, tol=TOL):
    return abs(a - b) <= tol * max(abs(b), 1e-3)


@pytest.mark.parametrize("case", CASES)
def test_steps_match_reference_golden(case):
    import timegan_b200 as tg
    from timegan_b200 import train_timegan as tt
    fx = StepFixture(case)
    model = tg.TimeGAN(fx.x_dim, fx.z_dim, fx.h_dim, fx.layers, 0.0)
    model.load_state_dict(fx.init)
    model = model.to(DEV)
    x = fx.x.to(DEV)
    P = tt._params
    optER = tg.FusedAdam(P(model.embedder, model.recovery), lr=HP["lr_g"], betas=HP["betas"])
    optS = tg.FusedAdam(model.supervisor.parameters(), lr=HP["lr_g"], betas=HP["betas"])
    optD = tg.FusedAdam(model.discriminator.parameters(), lr=HP["lr_d"], betas=HP["betas"])
    optG = tg.FusedAdam(P(model.generator, model.supervisor, model.embedder, model.recovery), lr=HP["lr_g"],
                        betas=HP["betas"])
    logs = []
    tt.phase_autoencoder(model, [(fx.x,)], DEV, optER, HP["clip"], 1, logs.append)
    g_ae = _grads(model, ("embedder", "recovery"))
    tt.phase_supervisor(model, [(fx.x,)], DEV, optS, HP["clip"], 1, logs.append)
    g_sup = _grads(model, ("supervisor",))
    assert _close(float(logs[0].split("recon=")[1]), fx.ae_loss, 2e-5 / max(fx.ae_loss, 1e-3) + TOL)
    assert abs(float(logs[1].split("sup=")[1]) - fx.sup_loss) < 2e-5 + TOL * fx.sup_loss
    torch.manual_seed(fx.seed + 1)
    nz = tt.HostReplayNoise(DEV)
    d_loss, d_acc = tt.disc_step(model, x, DEV, optD, HP["label_smooth"], HP["inst_noise"], HP["clip"], None,
                                 HP["r1_gamma"], target_acc=HP["target"], band=HP["band"], noise=nz)
    g_d = _grads(model, ("discriminator",))
    g_vals = tt.gen_step(model, x, DEV, optG, HP["alpha_sup"], HP["beta_rec"], HP["inst_noise"], HP["clip"], None,
                         HP["gamma_cov"], HP["gamma_acf"], HP["acf_max_lag"], noise=nz)
    g_g = _grads(model, ("generator", "supervisor", "embedder", "recovery"))
    assert _close(d_loss, fx.d_out[0]) and abs(d_acc - fx.d_out[1]) < 1e-6
    for got, ref, nm in zip(g_vals, fx.g_out, ("total", "adv", "sup", "rec", "cov", "acf")):
        assert _close(got, ref), (nm, got, ref)
    for step, got in (("ae", g_ae), ("sup", g_sup), ("d", g_d), ("g", g_g)):
        assert set(got) == set(fx.grads[step]), step
        for k, ref in fx.grads[step].items():
            assert relerr(got[k], ref) < TOL, (step, k, relerr(got[k], ref))
    sd = model.state_dict()
    for k, ref in fx.final.items():
        # Adam's first steps move every weight by ~lr regardless of |g|: compare absolutely against lr
        assert (sd[k].cpu() - ref).abs().max().item() < 0.05 * HP["lr_g"] + 1e-6, k


@pytest.mark.parametrize("cfg", [
    # x_dim, z, h, L, B, T, r1, band, gcov, gacf, lag, noise_std
    (14, 24, 24, 3, 8, 128, 1.0, 0.15, 0.05, 0.05, 64, 0.3),
    (14, 28, 56, 1, 5, 200, 1.0, 0.23, 0.03, 0.02, 48, 0.25),   # timegan_config.json values
    (14, 16, 16, 2, 6, 64, 0.0, 0.0, 0.0, 0.0, 8, 0.0),          # every optional term off
    (14, 64, 64, 3, 3, 64, 2.5, 0.15, 0.1, 0.1, 16, 0.1),        # config c2 dims
    (14, 256, 256, 2, 5, 24, 1.0, 0.15, 0.05, 0.05, 8, 0.2),     # config c4's H = 256 point (capacity fallback)
])
def test_joint_steps_match_oracle_port_live(cfg):
    """Three consecutive D+G updates from identical weights and identical (replayed) noise."""
    import timegan_b200 as tg
    from timegan_b200 import train_timegan as tt
    from oracle import timegan_ref as R
    x_dim, z, h, L, B, T, r1, band, gcov, gacf, lag, std = cfg
    torch.set_num_threads(max(1, torch.get_num_threads()))
    torch.manual_seed(77)
    port = R.build_model(x_dim, z, h, L, 0.0)
    ours = tg.TimeGAN(x_dim, z, h, L, 0.0)
    ours.load_state_dict(port.state_dict())
    ours = ours.to(DEV)
    x = torch.rand(B, T, x_dim)
    xd = x.to(DEV)
    op = R.make_optimizers(port)
    P = tt._params
    oD = tg.FusedAdam(ours.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9))
    oG = tg.FusedAdam(P(ours.generator, ours.supervisor, ours.embedder, ours.recovery), lr=1e-3, betas=(0.5, 0.9))
    for it in range(3):
        torch.manual_seed(1000 + it)
        d_p = R.d_step(port, x, op["D"], R.TorchNoise(), 0.2, std, 0.5, r1, 0.525, band)
        torch.manual_seed(1000 + it)
        nz = tt.HostReplayNoise(DEV)
        d_o = tt.disc_step(ours, xd, DEV, oD, 0.2, std, 0.5, None, r1, target_acc=0.525, band=band, noise=nz)
        # both sides continue on the same stream position: the port consumed exactly what HostReplayNoise did
        state = torch.get_rng_state()
        g_p = R.g_step(port, x, op["G"], R.TorchNoise(), 5.0, 0.2, std, 0.5, gcov, gacf, lag)
        torch.set_rng_state(state)
        g_o = tt.gen_step(ours, xd, DEV, oG, 5.0, 0.2, std, 0.5, None, gcov, gacf, lag, noise=nz)
        tol = TOL * (1 + 4 * it)   # later iterations start from weights that already differ by rounding
        assert _close(d_o[0], d_p[0], tol) and abs(d_o[1] - d_p[1]) < 1e-6, (it, d_o, d_p)
        for a, b in zip(g_o, g_p):
            assert _close(a, b, tol), (it, g_o, g_p)
    for (n, p), (_, q) in zip(ours.named_parameters(), port.named_parameters()):
        assert (p.detach().cpu() - q.detach()).abs().max().item() < 0.1 * 1e-3, n


def test_joint_step_bf16_projection_mode_within_2e_2():
    """north_star: "bf16 projection mode within 2e-2".  The reduced-precision projection mode (one TF32 tensor-core
    pass over the fp32 operands, ops.set_proj_mode("bf16")) against the fp32 CPU port on one D + G update at the
    config-c2/c3 layer shape (z = h = 64, 3 layers; T*B >= 128 rows so the tcgen05 tile is really taken)."""
    import timegan_b200 as tg
    from timegan_b200 import ops, train_timegan as tt
    from oracle import timegan_ref as R
    torch.manual_seed(5)
    port = R.build_model(14, 64, 64, 3, 0.0)
    ours = tg.TimeGAN(14, 64, 64, 3, 0.0)
    ours.load_state_dict(port.state_dict())
    ours = ours.to(DEV)
    x = torch.rand(4, 96, 14)
    op = R.make_optimizers(port)
    oD = tg.FusedAdam(ours.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9))
    oG = tg.FusedAdam(tt._params(ours.generator, ours.supervisor, ours.embedder, ours.recovery), lr=1e-3, betas=(0.5, 0.9))
    old = ops.get_proj_mode()
    try:
        ops.set_proj_mode("bf16")
        torch.manual_seed(21)
        d_p = R.d_step(port, x, op["D"], R.TorchNoise(), 0.2, 0.3, 0.5, 1.0, 0.525, 0.15)
        torch.manual_seed(21)
        nz = tt.HostReplayNoise(DEV)
        d_o = tt.disc_step(ours, x.to(DEV), DEV, oD, 0.2, 0.3, 0.5, None, 1.0, target_acc=0.525, band=0.15, noise=nz)
        state = torch.get_rng_state()
        g_p = R.g_step(port, x, op["G"], R.TorchNoise(), 5.0, 0.2, 0.3, 0.5, 0.05, 0.05, 32)
        torch.set_rng_state(state)
        g_o = tt.gen_step(ours, x.to(DEV), DEV, oG, 5.0, 0.2, 0.3, 0.5, None, 0.05, 0.05, 32, noise=nz)
    finally:
        ops.set_proj_mode(old)
    assert _close(d_o[0], d_p[0], 2e-2), (d_o, d_p)
    for a, b in zip(g_o, g_p):
        assert _close(a, b, 2e-2), (g_o, g_p)
    # Adam's first step moves every weight by +-lr whatever |g| is, so a 5e-4 relative gradient error may flip the
    # sign of a near-zero entry: bound the FRACTION of weights that moved differently, not the maximum
    bad = tot = 0
    for (n, p), (_, q) in zip(ours.named_parameters(), port.named_parameters()):
        d = (p.detach().cpu() - q.detach()).abs()
        bad += int((d > 0.5e-3).sum())
        tot += d.numel()
    assert bad / tot < 0.02, (bad, tot)


def test_generation_chain_matches_port():
    """decode(refine_latent(gen_latent(Z))) in eval mode, chunked (generate_long_synth.py:117-121)."""
    import timegan_b200 as tg
    from timegan_b200.generate_long_synth import generate_windows
    from timegan_b200 import train_timegan as tt
    from oracle import timegan_ref as R
    torch.manual_seed(3)
    port = R.build_model(14, 28, 56, 1, 0.2)
    ours = tg.TimeGAN(14, 28, 56, 1, 0.2)
    ours.load_state_dict(port.state_dict())
    ours = ours.to(DEV).eval()
    port.eval()
    torch.manual_seed(9)
    ref = R.generate(port, torch.rand(10, 100, 28))
    torch.manual_seed(9)
    got = generate_windows(ours, 10, 100, 28, DEV, chunk=10, noise=tt.HostReplayNoise(DEV))
    assert relerr(torch.from_numpy(got), ref) < TOL
    # chunking does not change the windows' values for a fixed Z stream (different chunk sizes, device noise)
    a = generate_windows(ours, 7, 50, 28, DEV, chunk=3)
    assert a.shape == (7, 50, 14) and a.dtype.name == "float32" and bool((a == a).all())


def test_cpu_tensor_is_refused():
    import timegan_b200 as tg
    m = tg.TimeGAN(14, 8, 8, 1, 0.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.encode(torch.rand(2, 5, 14))


def test_cuda_graph_replay_equals_eager_steps():
    """GraphedJointStep (captured disc_step + gen_step, device-side Adam clock / lr / Philox counter) reproduces
    the eagerly issued steps: same weights, same noise stream, 8 steps incl. an LR-schedule change."""
    import copy
    import timegan_b200 as tg
    from timegan_b200 import train_timegan as tt
    torch.manual_seed(5)
    base = tg.TimeGAN(14, 24, 24, 2, 0.0).to(DEV)
    xs = [torch.rand(6, 48, 14, device=DEV) for _ in range(8)]
    hp = dict(label_smooth=0.2, clip=0.5, r1_gamma=1.0, target_acc=0.525, band=0.15, alpha_sup=5.0, beta_rec=0.2,
              gamma_cov=0.05, gamma_acf=0.05, acf_max_lag=16)
    P = tt._params
    outs = {}
    finals = {}
    for mode in ("eager", "graph"):
        m = copy.deepcopy(base)
        cap = mode == "graph"
        oD = tg.FusedAdam(m.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9), capturable=cap)
        oG = tg.FusedAdam(P(m.generator, m.supervisor, m.embedder, m.recovery), lr=1e-3, betas=(0.5, 0.9), capturable=cap)
        sD = torch.optim.lr_scheduler.MultiStepLR(oD, milestones=[4, 6], gamma=0.5)
        sG = torch.optim.lr_scheduler.MultiStepLR(oG, milestones=[4, 6], gamma=0.5)
        nz = tt.device_noise(1234, torch.device(DEV))
        rows = []
        if cap:
            step = tt.GraphedJointStep(m, oD, oG, DEV, schedulerD=sD, schedulerG=sG, warmup=2, noise=nz, **hp)
            for i, x in enumerate(xs):
                rows.append(step(x, 0.3 - 0.01 * i).clone())
            assert step.graph is not None
        else:
            for i, x in enumerate(xs):
                d = tt.disc_step(m, x, DEV, oD, hp["label_smooth"], 0.3 - 0.01 * i, hp["clip"], sD, hp["r1_gamma"],
                                 target_acc=hp["target_acc"], band=hp["band"], noise=nz, sync=False)
                g = tt.gen_step(m, x, DEV, oG, hp["alpha_sup"], hp["beta_rec"], 0.3 - 0.01 * i, hp["clip"], sG,
                                hp["gamma_cov"], hp["gamma_acf"], hp["acf_max_lag"], noise=nz, sync=False)
                rows.append(torch.stack([v.float().reshape(()) for v in d + g]))
        outs[mode] = torch.stack(rows).cpu()
        finals[mode] = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        if cap:
            sd = oG.state_dict()
            assert float(sd["state"][0]["step"]) == len(xs)
    assert torch.isfinite(outs["graph"]).all()
    assert torch.allclose(outs["graph"], outs["eager"], rtol=2e-4, atol=1e-6), (outs["graph"] - outs["eager"]).abs().max()
    for k in finals["eager"]:
        assert (finals["graph"][k] - finals["eager"][k]).abs().max().item() < 2e-5, k

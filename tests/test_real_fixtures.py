"""Parity on the reference's OWN artefacts: a shipped checkpoint (posture1_no_exo/ckpt_best.pt, z = 28, h = 56, one
layer) and the NPZ it was trained on (26 x 768 x 14 real EEG windows -> N = 26 < batch 64, the short-batch case of
tt:33-37).  tests/golden/real/expected.npz holds what the UNMODIFIED reference computes from them
(oracle/make_golden_real.py).

CPU tests pin the oracle port and the state_dict / NPZ / loader contract; `gpu` tests run the product."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from parity_util import relerr

REAL = Path(__file__).resolve().parent / "golden" / "real"
TOL = 1e-4


def _fx():
    z = np.load(REAL / "expected.npz")
    hp = json.loads(str(z["hp"]))
    state = torch.load(REAL / "ckpt_best.pt", map_location="cpu")
    X = np.load(REAL / "posture1_no_exo.npz")["X"].astype(np.float32)
    return z, hp, state, X


def _step_args(hp):
    target = 0.5 * (hp["d_min_acc"] + hp["d_max_acc"])
    band = max(0.0, hp["d_max_acc"] - hp["d_min_acc"])
    return target, band, (hp["beta1"], hp["beta2"])


# ------------------------------------------------------------------------------------------------
# CPU: oracle port + contracts
# ------------------------------------------------------------------------------------------------
def test_port_reproduces_reference_on_real_checkpoint_bitwise():
    """oracle/timegan_ref.py == the unmodified reference on real weights and real data: forwards, one D + G update
    (losses and every pre-clip gradient), bit for bit."""
    from oracle import timegan_ref as R
    z, hp, state, X = _fx()
    C, zd, hd, L, N, T = [int(v) for v in z["dims"]]
    torch.set_num_threads(1)
    port = R.build_model(C, zd, hd, L, hp["dropout"])
    port.load_state_dict(state["model"])            # same module tree => same keys, strict
    x = torch.from_numpy(X)
    port.eval()
    with torch.no_grad():
        h = port["embedder"](x)
        assert torch.equal(h[:8], torch.from_numpy(z["fwd/h"]))
        assert torch.equal(port["recovery"](h)[:8], torch.from_numpy(z["fwd/x_tilde"]))
        assert torch.equal(port["discriminator"](h), torch.from_numpy(z["fwd/d_real"]))
        torch.manual_seed(123)
        assert torch.equal(R.generate(port, torch.rand(N, T, zd)), torch.from_numpy(z["gen/x_hat"]))
    port.train()
    first = x[torch.from_numpy(z["loader/order"])]
    target, band, betas = _step_args(hp)
    op = R.make_optimizers(port, hp["lr_g"], hp["lr_d"], betas)
    rec = []
    orig = R.clip_grad_norm_

    def recording_clip(ps, mx, *a, **k):
        ps = list(ps)
        rec.append({id(p): p.grad.clone() for p in ps if p.grad is not None})
        return orig(ps, mx, *a, **k)
    R.clip_grad_norm_ = recording_clip
    try:
        torch.manual_seed(7)
        d = R.d_step(port, first, op["D"], R.TorchNoise(), hp["label_smooth"], hp["inst_noise_start"], hp["grad_clip"],
                     hp["r1_gamma"], target, band)
        g = R.g_step(port, first, op["G"], R.TorchNoise(), hp["alpha_sup"], hp["beta_rec"], hp["inst_noise_start"],
                     hp["grad_clip"], hp["gamma_cov"], hp["gamma_acf"], hp["acf_max_lag"])
    finally:
        R.clip_grad_norm_ = orig
    assert list(d) == z["step/d_out"].tolist()
    assert list(g) == z["step/g_out"].tolist()
    names = {id(p): n for n, p in port.named_parameters()}
    for step, grads in zip(("d", "g"), rec):
        for pid, gr in grads.items():
            assert torch.equal(gr, torch.from_numpy(z[f"grad_{step}/{names[pid]}"])), (step, names[pid])


def test_product_model_accepts_the_shipped_state_dict():
    """The drop-in's module tree has exactly the reference's state_dict keys and shapes (strict load)."""
    import timegan_b200 as tg
    from timegan_b200.generate_long_synth import layers_in_state_dict
    z, hp, state, X = _fx()
    C, zd, hd, L, N, T = [int(v) for v in z["dims"]]
    assert layers_in_state_dict(state["model"]) == L
    m = tg.TimeGAN(C, zd, hd, L, hp["dropout"])
    missing = m.load_state_dict(state["model"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    sd = m.state_dict()
    assert list(sd.keys()) == list(state["model"].keys())
    for k, v in state["model"].items():
        assert torch.equal(sd[k], v), k


def test_loader_order_on_the_real_npz():
    """make_loader (tt:33-37) on the real NPZ: one ragged batch of all 26 windows, in the reference's order."""
    from timegan_b200 import train_timegan as tt
    z, hp, state, X = _fx()
    tt.set_seeds(hp["seed"])
    (first,) = next(iter(tt.make_loader(X, hp["batch_size"])))
    assert first.shape == (26, 768, 14)
    assert torch.equal(first, torch.from_numpy(X[z["loader/order"]]))


# ------------------------------------------------------------------------------------------------
# GPU: the product on the real checkpoint / data
# ------------------------------------------------------------------------------------------------
def _product(state, dims, dropout):
    import timegan_b200 as tg
    C, zd, hd, L, N, T = dims
    m = tg.TimeGAN(C, zd, hd, L, dropout)
    m.load_state_dict(state["model"])
    return m.to("cuda:0")


@pytest.mark.gpu
def test_forward_parity_on_real_data():
    from timegan_b200 import train_timegan as tt
    from timegan_b200.generate_long_synth import generate_windows
    z, hp, state, X = _fx()
    dims = [int(v) for v in z["dims"]]
    m = _product(state, dims, hp["dropout"]).eval()
    x = torch.from_numpy(X).cuda()
    with torch.no_grad():
        h = m.encode(x)
        assert relerr(h[:8], torch.from_numpy(z["fwd/h"])) < TOL
        assert relerr(m.reconstruct(x)[:8], torch.from_numpy(z["fwd/x_tilde"])) < TOL
        assert relerr(m.disc(h), torch.from_numpy(z["fwd/d_real"])) < TOL
    torch.manual_seed(123)
    got = generate_windows(m, dims[4], dims[5], dims[1], torch.device("cuda:0"), chunk=64,
                           noise=tt.HostReplayNoise(torch.device("cuda:0")))
    assert relerr(torch.from_numpy(got), torch.from_numpy(z["gen/x_hat"])) < TOL


@pytest.mark.gpu
def test_joint_step_from_shipped_checkpoint_on_short_batch():
    """disc_step + gen_step from the shipped weights on the real N = 26 < B = 64 batch with timegan_config.json's
    hyper-parameters: losses and every pre-clip gradient within 1e-4 of the unmodified reference."""
    import timegan_b200 as tg
    from timegan_b200 import train_timegan as tt
    z, hp, state, X = _fx()
    dims = [int(v) for v in z["dims"]]
    m = _product(state, dims, hp["dropout"]).train()
    dev = torch.device("cuda:0")
    first = torch.from_numpy(X[z["loader/order"]]).to(dev)
    target, band, betas = _step_args(hp)
    oD = tg.FusedAdam(m.discriminator.parameters(), lr=hp["lr_d"], betas=betas)
    oG = tg.FusedAdam(tt._params(m.generator, m.supervisor, m.embedder, m.recovery), lr=hp["lr_g"], betas=betas)
    torch.manual_seed(7)
    nz = tt.HostReplayNoise(dev)
    d = tt.disc_step(m, first, dev, oD, hp["label_smooth"], hp["inst_noise_start"], hp["grad_clip"], None,
                     hp["r1_gamma"], target_acc=target, band=band, noise=nz)
    gd = {n: p.grad.detach().clone() for n, p in m.named_parameters() if n.startswith("discriminator") and p.grad is not None}
    g = tt.gen_step(m, first, dev, oG, hp["alpha_sup"], hp["beta_rec"], hp["inst_noise_start"], hp["grad_clip"], None,
                    hp["gamma_cov"], hp["gamma_acf"], hp["acf_max_lag"], noise=nz)
    gg = {n: p.grad.detach().clone() for n, p in m.named_parameters() if not n.startswith("discriminator") and p.grad is not None}
    close = lambda a, b: abs(a - b) <= TOL * max(abs(b), 1e-3)
    assert close(d[0], z["step/d_out"][0]) and abs(d[1] - z["step/d_out"][1]) < 1e-6, (d, z["step/d_out"])
    for a, b, nm in zip(g, z["step/g_out"], ("total", "adv", "sup", "rec", "cov", "acf")):
        assert close(a, b), (nm, g, z["step/g_out"])
    for step, got in (("d", gd), ("g", gg)):
        keys = [k.split("/", 1)[1] for k in z.files if k.startswith(f"grad_{step}/")]
        assert set(got) == set(keys), step
        for k in keys:
            assert relerr(got[k], torch.from_numpy(z[f"grad_{step}/{k}"])) < TOL, (step, k)


@pytest.mark.gpu
def test_generate_long_synth_cli_on_the_shipped_run(tmp_path):
    """generate_long_synth.main (gl:43-131) over a runs dir holding the shipped checkpoint: same files, same shapes,
    finite values in the scaled space, --denorm inverts the NPZ's scaling, --gen_len changes T."""
    import shutil
    from timegan_b200 import generate_long_synth as gl
    runs, real = tmp_path / "runs", tmp_path / "real"
    (runs / "posture1_no_exo").mkdir(parents=True)
    real.mkdir()
    shutil.copyfile(REAL / "ckpt_best.pt", runs / "posture1_no_exo" / "ckpt_best.pt")
    shutil.copyfile(REAL / "posture1_no_exo.npz", real / "posture1_no_exo.npz")
    gl.main(["--runs_dir", str(runs), "--real_dir", str(real)])
    a = np.load(runs / "posture1_no_exo" / "synthetic_long.npz")["X"]
    assert a.shape == (26, 768, 14) and a.dtype == np.float32 and np.isfinite(a).all()
    ref = np.load(REAL / "expected.npz")["gen/x_hat"]
    # a different Z stream (device Philox vs the reference's CPU draw): the generated windows share the statistics
    assert abs(float(a.mean()) - float(ref.mean())) < 0.05 and abs(float(a.std()) - float(ref.std())) < 0.05
    gl.main(["--runs_dir", str(runs), "--real_dir", str(real), "--gen_len", "1000", "--n", "5", "--denorm",
             "--out_suffix", "synthetic_T{T}.npz"])
    b = np.load(runs / "posture1_no_exo" / "synthetic_T1000.npz")["X"]
    assert b.shape == (5, 1000, 14) and np.isfinite(b).all()
    sc = np.load(real / "posture1_no_exo.npz")
    # --denorm maps the scaled space back to microvolts with the NPZ's per-channel affine map (gl:123-126): the spread
    # of every channel is a sizeable fraction of its recorded range, which no [0,1]-scaled output would have
    sd = b.std((0, 1))
    assert (sd > 0.02 * sc["scale_range"]).all() and (sd < 2.0 * sc["scale_range"]).all(), sd

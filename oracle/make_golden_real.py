"""Fixtures from the reference's OWN artefacts (SURVEY.md section 4 "hard pins" of the state_dict / NPZ contract).

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/make_golden_real.py

Copies, byte for byte, one shipped checkpoint and the NPZ it was trained on
    timeGAN/timegan_runs/posture1_no_exo/ckpt_best.pt   (z = 28, h = 56, 1 layer; torch.save dict of tt:58-61)
    timeGAN/preprocessed/posture1_no_exo.npz            (X: 26 x 768 x 14 -> N = 26 < batch 64: the short-batch case)
into tests/golden/real/, then runs the UNMODIFIED reference (imported from /root/reference/timeGAN) on them and
stores what it computes in tests/golden/real/expected.npz:
    eval-mode forwards    encode(X)[:8], reconstruct(X)[:8], disc(encode(X)), the generation chain of gl:117-121 on a seeded Z
    one joint step        disc_step + gen_step (tt:166-276) from the checkpoint's weights on the real short batch with
                          the hyper-parameters of timeGAN/timegan_config.json: losses and pre-clip gradients
    the first loader batch of make_loader(X, 64) after set_seeds(42) (tt:33-37): its row order
"""
import json
import shutil
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/timeGAN")
OUT = ROOT / "tests" / "golden" / "real"
RUN, NPZ = "posture1_no_exo", "posture1_no_exo.npz"


def main():
    sys.path.insert(0, str(REF))
    import timegan_model as tm    # noqa: E402  the unmodified reference
    import train_timegan as tt    # noqa: E402
    torch.set_num_threads(1)
    OUT.mkdir(parents=True, exist_ok=True)
    shutil.copyfile(REF / "timegan_runs" / RUN / "ckpt_best.pt", OUT / "ckpt_best.pt")
    shutil.copyfile(REF / "preprocessed" / NPZ, OUT / NPZ)
    cfg = json.loads((REF / "timegan_config.json").read_text())

    state = torch.load(OUT / "ckpt_best.pt", map_location="cpu", weights_only=False)
    X = np.load(OUT / NPZ)["X"].astype(np.float32)
    N, T, C = X.shape
    z_dim, h_dim = int(state["meta"]["z_dim"]), int(state["meta"]["h_dim"])
    assert (z_dim, h_dim) == tt.adaptive_dims(C, T)
    model = tm.TimeGAN(x_dim=C, z_dim=z_dim, hidden_dim=h_dim, num_layers=cfg["layers"], dropout=cfg["dropout"])
    model.load_state_dict(state["model"])
    x = torch.from_numpy(X)
    out = {"dims": np.array([C, z_dim, h_dim, cfg["layers"], N, T])}

    # ---- eval-mode forwards -------------------------------------------------------------------------------
    model.eval()
    with torch.no_grad():
        h = model.encode(x)
        out["fwd/h"] = h[:8].numpy().copy()                               # first 8 windows (fixture size)
        out["fwd/x_tilde"] = model.reconstruct(x)[:8].numpy().copy()
        out["fwd/d_real"] = model.disc(h).numpy().copy()
        torch.manual_seed(123)
        Z = tt.sample_noise(N, T, z_dim, torch.device("cpu"))
        out["gen/x_hat"] = model.decode(model.refine_latent(model.gen_latent(Z))).numpy().copy()

    # ---- loader order (tt:33-37) ---------------------------------------------------------------------------
    tt.set_seeds(cfg["seed"])
    (first,) = next(iter(tt.make_loader(X, cfg["batch_size"])))
    assert first.shape[0] == N            # N = 26 < 64: ONE ragged batch per epoch
    order = [int(np.argmin(np.abs(X - first[i].numpy()[None]).reshape(N, -1).sum(1))) for i in range(N)]
    out["loader/order"] = np.array(order)

    # ---- one joint step from the checkpoint's weights (train mode; 1-layer GRUs have no dropout) -----------
    model.train()
    captured = []
    orig_clip = torch.nn.utils.clip_grad_norm_

    def recording_clip(params, max_norm, *a, **kw):
        params = list(params)
        captured.append([None if p.grad is None else p.grad.detach().clone() for p in params])
        return orig_clip(params, max_norm, *a, **kw)

    names = {id(p): n for n, p in model.named_parameters()}
    plist = lambda *mods: [p for m in mods for p in m.parameters()]
    betas = (cfg["beta1"], cfg["beta2"])
    optD = torch.optim.Adam(model.discriminator.parameters(), lr=cfg["lr_d"], betas=betas)
    optG = torch.optim.Adam(plist(model.generator, model.supervisor, model.embedder, model.recovery), lr=cfg["lr_g"],
                            betas=betas)
    target = 0.5 * (cfg["d_min_acc"] + cfg["d_max_acc"])
    band = max(0.0, cfg["d_max_acc"] - cfg["d_min_acc"])
    dev = torch.device("cpu")
    torch.nn.utils.clip_grad_norm_ = recording_clip
    try:
        torch.manual_seed(7)
        d_loss, d_acc = tt.disc_step(model, first, dev, optD, cfg["label_smooth"], cfg["inst_noise_start"],
                                     cfg["grad_clip"], None, cfg["r1_gamma"], target_acc=target, band=band)
        g_vals = tt.gen_step(model, first, dev, optG, cfg["alpha_sup"], cfg["beta_rec"], cfg["inst_noise_start"],
                             cfg["grad_clip"], None, cfg["gamma_cov"], cfg["gamma_acf"], cfg["acf_max_lag"])
    finally:
        torch.nn.utils.clip_grad_norm_ = orig_clip
    out["step/d_out"] = np.array([d_loss, d_acc], dtype=np.float64)
    out["step/g_out"] = np.array(g_vals, dtype=np.float64)
    groups = [list(model.discriminator.parameters()),
              plist(model.generator, model.supervisor, model.embedder, model.recovery)]
    for step, params, grads in zip(("d", "g"), groups, captured):
        for p, g in zip(params, grads):
            if g is not None:
                out[f"grad_{step}/{names[id(p)]}"] = g.numpy().copy()
    out["hp"] = np.array(json.dumps({k: cfg[k] for k in (
        "batch_size", "lr_g", "lr_d", "beta1", "beta2", "alpha_sup", "beta_rec", "label_smooth", "inst_noise_start",
        "grad_clip", "layers", "dropout", "seed", "r1_gamma", "d_min_acc", "d_max_acc", "gamma_cov", "gamma_acf",
        "acf_max_lag")}))
    np.savez_compressed(OUT / "expected.npz", **out)
    print("d", out["step/d_out"], "g", out["step/g_out"], "order", order[:6], "...")


if __name__ == "__main__":
    main()

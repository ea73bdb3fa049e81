#!/usr/bin/env python3
"""Golden vectors for the four optimiser steps, produced by the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference); writes
tests/golden/steps_<case>.npz, which the parity tests read on the GPU box (where /root/reference is absent).

Per case: seed -> reference TimeGAN (timeGAN/timegan_model.py:101-111) -> one batch through
phase_autoencoder (train_timegan.py:131-144), phase_supervisor (tt:147-163), disc_step (tt:166-225) and
gen_step (tt:228-276), in that order on the same evolving model with the four Adam instances of
train_single_npz (tt:331-345).  Recorded: the initial state_dict, the batch, every returned/printed loss, the
gradients of each step BEFORE clipping (captured by wrapping torch.nn.utils.clip_grad_norm_, the function the
reference calls -- the reference's files are not touched), and the state_dict after the four steps.
Random draws come from torch's global CPU generator seeded with `seed + 1` right before disc_step
(SURVEY.md Appendix B order), so a consumer can replay them.
"""
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/timeGAN")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

# hyper-parameters: CLI defaults of train_timegan.py:429-456 (target/band as derived at tt:377-378)
HP = dict(lr_g=1e-3, lr_d=2e-4, betas=(0.5, 0.9), alpha_sup=5.0, beta_rec=0.2, label_smooth=0.2, inst_noise=0.3,
          clip=0.5, r1_gamma=1.0, target=0.5 * (0.45 + 0.60), band=0.60 - 0.45, gamma_cov=0.05, gamma_acf=0.05,
          acf_max_lag=64)

CASES = {
    # name: (x_dim, z_dim, hidden_dim, layers, B, T, seed)
    "tiny": (5, 6, 8, 2, 3, 16, 11),          # h != z -> Linear projections in G/S; z % 4 != 0
    "c1": (14, 24, 24, 3, 4, 96, 42),         # BASELINE config 1 dims (short T)
    "refdefault": (14, 28, 56, 1, 4, 64, 7),  # adaptive_dims(14, 768) with layers=1 (all shipped checkpoints)
}


def run_case(name, tm, tt):
    x_dim, z_dim, h_dim, layers, B, T, seed = CASES[name]
    torch.manual_seed(seed)
    model = tm.TimeGAN(x_dim=x_dim, z_dim=z_dim, hidden_dim=h_dim, num_layers=layers, dropout=0.0)
    x = torch.rand(B, T, x_dim)
    out = {"x": x.numpy().copy(), "dims": np.array([x_dim, z_dim, h_dim, layers, B, T, seed])}
    for k, v in model.state_dict().items():
        out[f"init/{k}"] = v.detach().numpy().copy()

    captured = []
    orig_clip = torch.nn.utils.clip_grad_norm_

    def recording_clip(params, max_norm, *a, **kw):
        params = list(params)
        captured.append([None if p.grad is None else p.grad.detach().clone() for p in params])
        return orig_clip(params, max_norm, *a, **kw)

    names = {id(p): n for n, p in model.named_parameters()}
    plist = lambda *mods: [p for m in mods for p in m.parameters()]
    optER = torch.optim.Adam(plist(model.embedder, model.recovery), lr=HP["lr_g"], betas=HP["betas"])
    optS = torch.optim.Adam(model.supervisor.parameters(), lr=HP["lr_g"], betas=HP["betas"])
    optD = torch.optim.Adam(model.discriminator.parameters(), lr=HP["lr_d"], betas=HP["betas"])
    optG = torch.optim.Adam(plist(model.generator, model.supervisor, model.embedder, model.recovery), lr=HP["lr_g"],
                            betas=HP["betas"])
    logs = []
    dev = torch.device("cpu")
    torch.nn.utils.clip_grad_norm_ = recording_clip
    try:
        tt.phase_autoencoder(model, [(x,)], dev, optER, HP["clip"], 1, logs.append)
        tt.phase_supervisor(model, [(x,)], dev, optS, HP["clip"], 1, logs.append)
        torch.manual_seed(seed + 1)
        d_loss, d_acc = tt.disc_step(model, x, dev, optD, HP["label_smooth"], HP["inst_noise"], HP["clip"], None,
                                     HP["r1_gamma"], target_acc=HP["target"], band=HP["band"])
        g_vals = tt.gen_step(model, x, dev, optG, HP["alpha_sup"], HP["beta_rec"], HP["inst_noise"], HP["clip"], None,
                             HP["gamma_cov"], HP["gamma_acf"], HP["acf_max_lag"])
    finally:
        torch.nn.utils.clip_grad_norm_ = orig_clip
    out["ae_loss"] = np.float64(logs[0].split("recon=")[1])
    out["sup_loss"] = np.float64(logs[1].split("sup=")[1])
    out["d_out"] = np.array([d_loss, d_acc], dtype=np.float64)
    out["g_out"] = np.array(g_vals, dtype=np.float64)
    groups = [plist(model.embedder, model.recovery), list(model.supervisor.parameters()),
              list(model.discriminator.parameters()),
              plist(model.generator, model.supervisor, model.embedder, model.recovery)]
    for step, params, grads in zip(("ae", "sup", "d", "g"), groups, captured):
        for p, g in zip(params, grads):
            if g is not None:
                out[f"grad_{step}/{names[id(p)]}"] = g.numpy().copy()
    for k, v in model.state_dict().items():
        out[f"final/{k}"] = v.detach().numpy().copy()
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f"steps_{name}.npz", **out)
    print(name, "ae", out["ae_loss"], "sup", out["sup_loss"], "d", out["d_out"], "g", out["g_out"])


def main():
    sys.path.insert(0, str(REF))
    import timegan_model as tm   # noqa: E402  (the unmodified reference)
    import train_timegan as tt   # noqa: E402
    torch.set_num_threads(1)
    for name in (sys.argv[1:] or CASES):
        run_case(name, tm, tt)


if __name__ == "__main__":
    main()

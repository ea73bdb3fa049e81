"""CPU restatement of the reference's TimeGAN training step in plain PyTorch.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (eeg-gan-timegan-cgan_b200/, imported as
timegan_b200) imports this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg do.  The reference is pure Python on top of PyTorch and cannot travel to the GPU box,
so this port stands in for it there: it issues the SAME library calls (torch.nn.GRU, nn.Linear, legacy
spectral_norm, BCELoss, autograd double backward for R1, clip_grad_norm_, optim.Adam) in the same order, so
both its results and its CPU cost are the reference's.

Parity pin: tests/test_oracle_port.py runs this port next to the UNMODIFIED reference imported from
/root/reference/timeGAN (when present, i.e. in the build container) on identical seeds and requires
bit-identical losses and gradients; tests/golden/steps_*.npz hold outputs of the unmodified reference
(generator: oracle/make_golden_steps.py) and are checked against both this port and the CUDA path.

Reference lines restated (timeGAN/timegan_model.py = tm, timeGAN/train_timegan.py = tt):
  build_model     tm:11-21 (init), tm:24-34 (GRUStack), tm:37-98 (five networks), tm:101-118 (bundle)
  ae_step         tt:131-144      sup_step   tt:147-163
  d_step          tt:166-225      g_step     tt:228-276
  losses          tt:70-126       noise      tt:40-47, 64-65
The only structural difference: random draws go through a `noise` object (`rand(*shape)`, `randn_like(t)`)
so a test can feed both sides the same numbers; `TorchNoise` draws from torch's global generator with the
reference's calls in the reference's order.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
from torch.nn.utils import clip_grad_norm_, spectral_norm


class TorchNoise:
    """torch.rand / torch.randn_like on the global CPU generator -- what the reference does on device=cpu."""

    def rand(self, *shape):
        return torch.rand(*shape)

    def randn_like(self, t):
        return torch.randn_like(t)


# ------------------------------------------------------------------------------------------------
# model (same module tree => same state_dict keys and the same RNG consumption at construction)
# ------------------------------------------------------------------------------------------------
class _Stack(nn.Module):          # tm:24-34
    def __init__(self, n_in, n_hidden, n_layers, p_drop):
        super().__init__()
        self.rnn = nn.GRU(n_in, n_hidden, num_layers=n_layers, dropout=p_drop if n_layers > 1 else 0.0,
                          batch_first=True)

    def forward(self, seq):
        return self.rnn(seq)[0]


class _Net(nn.Module):
    """One of the five networks: GRU stack + optional head, attribute names as in tm:37-98."""

    def __init__(self, n_in, n_hidden, n_layers, p_drop, head=None, head_out=None):
        super().__init__()
        self.rnn = _Stack(n_in, n_hidden, n_layers, p_drop)
        self.kind = head
        if head == "out":                       # Recovery, tm:53
            self.out = nn.Linear(n_hidden, head_out)
        elif head == "proj":                    # Generator / Supervisor, tm:66,79
            self.proj = nn.Linear(n_hidden, head_out) if n_hidden != head_out else nn.Identity()
        elif head == "fc":                      # Discriminator, tm:92-93
            self.fc = spectral_norm(nn.Linear(n_hidden, 1))
            self.sigmoid = nn.Sigmoid()

    def forward(self, seq):
        y = self.rnn(seq)
        if self.kind == "out":
            return self.out(y)
        if self.kind == "proj":
            return self.proj(y)
        if self.kind == "fc":
            return self.sigmoid(self.fc(y[:, -1, :]))       # tm:96-98
        return y


def _reference_init(module):                   # tm:11-21
    if isinstance(module, nn.Linear):
        nn.init.xavier_uniform_(module.weight)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    if isinstance(module, (nn.GRU, nn.LSTM)):
        for key, value in module.named_parameters():
            if "weight" in key:
                nn.init.xavier_uniform_(value)
            elif "bias" in key:
                nn.init.zeros_(value)


def build_model(x_dim, z_dim, hidden_dim, num_layers=2, dropout=0.1):
    """tm:101-111: embedder, recovery, generator, supervisor, discriminator (in this order), then init."""
    nets = OrderedDict(
        embedder=_Net(x_dim, z_dim, num_layers, dropout),
        recovery=_Net(z_dim, hidden_dim, num_layers, dropout, "out", x_dim),
        generator=_Net(z_dim, hidden_dim, num_layers, dropout, "proj", z_dim),
        supervisor=_Net(z_dim, hidden_dim, num_layers, dropout, "proj", z_dim),
        discriminator=_Net(z_dim, hidden_dim, num_layers, dropout, "fc"),
    )
    model = nn.ModuleDict(nets)
    model.apply(_reference_init)
    return model


def latent_size(model):
    return model["embedder"].rnn.rnn.hidden_size            # read at tt:179,235


# ------------------------------------------------------------------------------------------------
# losses (tt:70-126)
# ------------------------------------------------------------------------------------------------
_bce = nn.BCELoss()


def recon(x, x_rec, eps=1e-8):                               # tt:72-74
    return 10.0 * torch.sqrt(torch.mean((x - x_rec) ** 2) + eps)


def first_difference(h):                                     # tt:79-80
    return torch.mean((h[:, 1:, :] - h[:, :-1, :]) ** 2)


def channel_cov(x):                                          # tt:82-101
    flat = x.reshape(x.shape[0] * x.shape[1], x.shape[2])
    flat = flat - flat.mean(dim=0, keepdim=True)
    return (flat.t() @ flat) / (flat.size(0) - 1)


def acf_table(x, lags):                                      # tt:110-122
    mu = x.mean(dim=(0, 1), keepdim=True)
    sd = x.std(dim=(0, 1), keepdim=True) + 1e-8
    xz = (x - mu) / sd
    return torch.stack([(xz[:, :-k, :] * xz[:, k:, :]).mean(dim=(0, 1)) for k in range(1, lags + 1)], dim=0)


def acf_l1(x_gen, x_real, max_lag):                          # tt:103-126
    lags = max(1, min(max_lag, x_gen.shape[1] - 1))
    with torch.no_grad():
        ref = acf_table(x_real, lags)
    return torch.mean(torch.abs(acf_table(x_gen, lags) - ref))


def _jitter(h, std, noise):                                  # tt:46-47
    return h if std <= 0 else h + std * noise.randn_like(h)


# ------------------------------------------------------------------------------------------------
# the four optimiser steps
# ------------------------------------------------------------------------------------------------
def _plist(model, *names):
    out = []
    for n in names:
        out += list(model[n].parameters())
    return out


def ae_step(model, x, opt, clip):
    """One batch of phase_autoencoder (tt:136-143).  Returns the loss tensor."""
    loss = recon(x, model["recovery"](model["embedder"](x)))
    opt.zero_grad(set_to_none=True)
    loss.backward()
    clip_grad_norm_(_plist(model, "embedder", "recovery"), clip)
    opt.step()
    return loss.detach()


def sup_step(model, x, opt, clip):
    """One batch of phase_supervisor (tt:152-161)."""
    with torch.no_grad():
        h = model["embedder"](x)
    pred = model["supervisor"](h[:, :-1, :])
    loss = torch.mean((pred - h[:, 1:, :]) ** 2)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    clip_grad_norm_(model["supervisor"].parameters(), clip)
    opt.step()
    return loss.detach()


def d_step(model, x, opt, noise, label_smooth, inst_noise_std, clip, r1_gamma=1.0, target_acc=0.55, band=0.10,
           scheduler=None):
    """disc_step (tt:166-225).  Returns (loss, acc) as Python floats like the reference."""
    D = model["discriminator"]
    D.train()
    n, t = x.size(0), x.size(1)
    with torch.no_grad():
        h_real = model["embedder"](x)
    z = noise.rand(n, t, latent_size(model))
    h_fake = model["supervisor"](model["generator"](z))
    h_real_n = _jitter(h_real, inst_noise_std, noise).requires_grad_(True)
    h_fake_n = _jitter(h_fake.detach(), inst_noise_std, noise)
    y_real = (1.0 - label_smooth) + label_smooth * noise.rand(n, 1)     # tt:41
    y_fake = label_smooth * noise.rand(n, 1)                            # tt:42
    with torch.backends.cudnn.flags(enabled=False):
        p_real = D(h_real_n)
    p_fake = D(h_fake_n)
    loss = 0.5 * (_bce(p_real, y_real) + _bce(p_fake, y_fake))
    if r1_gamma > 0.0:
        g = torch.autograd.grad(p_real.sum(), h_real_n, create_graph=True, retain_graph=True)[0]
        loss = loss + 0.5 * r1_gamma * g.reshape(g.size(0), -1).pow(2).sum(1).mean()
    with torch.no_grad():
        acc = 0.5 * ((p_real > 0.5).float().mean().item() + (p_fake < 0.5).float().mean().item())
    if band > 0:
        loss = loss * max(0.2, 1.0 - max(0.0, acc - target_acc) / band)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    clip_grad_norm_(D.parameters(), clip)
    opt.step()
    if scheduler is not None:
        scheduler.step()
    return loss.item(), acc


def g_step(model, x, opt, noise, alpha_sup, beta_rec, inst_noise_std, clip, gamma_cov=0.0, gamma_acf=0.0,
           acf_max_lag=32, scheduler=None):
    """gen_step (tt:228-276).  Returns the six logged floats."""
    for name in ("generator", "supervisor", "embedder", "recovery"):
        model[name].train()
    n, t = x.size(0), x.size(1)
    z = noise.rand(n, t, latent_size(model))
    h_hat = model["supervisor"](model["generator"](z))
    p_fake = model["discriminator"](_jitter(h_hat, inst_noise_std, noise))
    adv = _bce(p_fake, torch.ones_like(p_fake))
    sup = first_difference(h_hat)
    rec = recon(x, model["recovery"](model["embedder"](x)))
    x_hat = model["recovery"](h_hat)
    cov = torch.tensor(0.0)
    if gamma_cov > 0:
        with torch.no_grad():
            c_real = channel_cov(x.detach())
        cov = torch.norm(channel_cov(x_hat) - c_real, p="fro") / (c_real.numel() ** 0.5)
    acf = torch.tensor(0.0)
    if gamma_acf > 0:
        acf = acf_l1(x_hat, x.detach(), acf_max_lag)
    total = adv + alpha_sup * sup + beta_rec * rec + gamma_cov * cov + gamma_acf * acf
    opt.zero_grad(set_to_none=True)
    total.backward()
    clip_grad_norm_(_plist(model, "generator", "supervisor", "embedder", "recovery"), clip)
    opt.step()
    if scheduler is not None:
        scheduler.step()
    return tuple(v.item() for v in (total, adv, sup, rec, cov, acf))


def generate(model, z):
    """decode(refine_latent(gen_latent(z))) -- tt:417-419 / generate_long_synth.py:117-121."""
    with torch.no_grad():
        return model["recovery"](model["supervisor"](model["generator"](z)))


def make_optimizers(model, lr_g=1e-3, lr_d=2e-4, betas=(0.5, 0.9)):
    """The four Adam instances of train_single_npz (tt:331,336,340-345)."""
    A = torch.optim.Adam
    return dict(
        ER=A(_plist(model, "embedder", "recovery"), lr=lr_g, betas=betas),
        S=A(model["supervisor"].parameters(), lr=lr_g, betas=betas),
        D=A(model["discriminator"].parameters(), lr=lr_d, betas=betas),
        G=A(_plist(model, "generator", "supervisor", "embedder", "recovery"), lr=lr_g, betas=betas),
    )

"""Host side of csrc/head.cu: the discriminator head of timegan_model.py:92-98 (spectral-norm Linear(h, 1) + sigmoid)
fused with what train_timegan.py does to its output -- BCE (tt:70,196,241), balanced accuracy and the soft throttle
(tt:205-215), the R1 seed and the reverse pass of all of it -- in four one-CTA kernels instead of ~190 ATen launches
per joint step.

disc_step (tt:166-225)              gen_step (tt:240-241, D frozen)
  forward(D, y_last[2B,H], labels)    adv_loss(D, y_last[B,H])  -> autograd.Function returning g_adv
  seed(...)   -> scale, R1 seed, dL/dy_last(fake)
  backward(...) after the tangent forward -> dL/dy_last(real), dL/d tangent, fc gradients, logged loss
Under data parallelism the four batch sums pass through dist.allreduce_stats between forward() and seed(), so the
accuracy, the throttle scale and the BCE mean are the global batch's on every rank.
"""
from typing import Optional

import torch
from torch.autograd.function import once_differentiable

from . import dist as _dist
from ._lib import lib, check, ptr, stream_ptr, require_cuda


class HeadState:
    """Outputs of tg_head_fwd that the later kernels read."""
    __slots__ = ("B", "H", "n_half", "wbar", "uv", "sigma", "p", "stats", "labels")


def _row_stride(t: torch.Tensor) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("head: y_last must be a (rows, H) view with unit stride along H")
    return t.stride(0)


@torch.no_grad()
def forward(D, y_last: torch.Tensor, labels: Optional[torch.Tensor], n_half: int) -> HeadState:
    """Power iteration(s) + probabilities + local batch sums.  y_last: (n_half*B, H), real rows first."""
    require_cuda(y_last, "discriminator head input")
    rows, H = y_last.shape
    B = rows // n_half
    dev = y_last.device
    fc = D.fc
    st = HeadState()
    st.B, st.H, st.n_half = B, H, n_half
    st.wbar = torch.empty(n_half, H, dtype=torch.float32, device=dev)
    st.uv = torch.empty(n_half, 1 + H, dtype=torch.float32, device=dev)
    st.sigma = torch.empty(n_half, dtype=torch.float32, device=dev)
    st.p = torch.empty(rows, dtype=torch.float32, device=dev)
    st.stats = torch.zeros(4, dtype=torch.float32, device=dev)
    st.labels = None if labels is None else labels.reshape(-1).contiguous()
    check(lib.tg_head_fwd(stream_ptr(), ptr(y_last), _row_stride(y_last), B, H, n_half, ptr(fc.weight_orig.detach()),
                          ptr(fc.bias.detach()), ptr(fc.weight_u), ptr(fc.weight_v), int(D.training), ptr(st.labels),
                          ptr(st.wbar), ptr(st.uv), ptr(st.sigma), ptr(st.p), ptr(st.stats)), "tg_head_fwd")
    return st


@torch.no_grad()
def seed(st: HeadState, stats: torch.Tensor, Bg: float, target_acc: float, band: float, need_seed: bool):
    """(scal = [loss_bce, acc, scale, 0], R1 seed (B,H) or None, dL/dy_last(fake) (B,H)) from the GLOBAL batch sums."""
    dev = st.p.device
    scal = torch.empty(4, dtype=torch.float32, device=dev)
    sd = torch.empty(st.B, st.H, dtype=torch.float32, device=dev) if need_seed else None
    gyf = torch.empty(st.B, st.H, dtype=torch.float32, device=dev)
    check(lib.tg_head_seed(stream_ptr(), ptr(st.p), ptr(st.labels), ptr(st.wbar), ptr(stats.contiguous()), ptr(scal),
                           ptr(sd), ptr(gyf), st.B, st.H, float(Bg), float(target_acc), float(band)), "tg_head_seed")
    return scal, sd, gyf


@torch.no_grad()
def backward(D, st: HeadState, y_last: torch.Tensor, hd_last: Optional[torch.Tensor], scal: torch.Tensor,
             r1: Optional[torch.Tensor], Bg: float, gamma: float):
    """Returns (gyr, ghd or None, g_weight_orig (1,H), g_bias (1,), loss_val (1,))."""
    dev = st.p.device
    B, H = st.B, st.H
    gyr = torch.empty(B, H, dtype=torch.float32, device=dev)
    ghd = torch.empty(B, H, dtype=torch.float32, device=dev) if hd_last is not None else None
    gw = torch.empty(1, H, dtype=torch.float32, device=dev)
    gb = torch.empty(1, dtype=torch.float32, device=dev)
    loss_val = torch.empty(1, dtype=torch.float32, device=dev)
    r1c = None if r1 is None else r1.reshape(1).contiguous()
    check(lib.tg_head_bwd(stream_ptr(), ptr(y_last), _row_stride(y_last), ptr(hd_last),
                          _row_stride(hd_last) if hd_last is not None else 0, ptr(st.p), ptr(st.labels),
                          ptr(D.fc.weight_orig.detach()), ptr(st.wbar), ptr(st.uv), ptr(st.sigma), ptr(scal), ptr(r1c),
                          ptr(gyr), ptr(ghd), ptr(gw), ptr(gb), ptr(loss_val), B, H, float(Bg), float(gamma)),
          "tg_head_bwd")
    return gyr, ghd, gw, gb, loss_val


class _AdvLoss(torch.autograd.Function):
    """g_adv = bce(D_head(y_last), ones) with D's weights as constants (gen_step, tt:240-241)."""

    @staticmethod
    def forward(ctx, y_last, D):
        y_last = y_last.contiguous()
        st = forward(D, y_last, None, 1)
        s, n = _dist.allreduce_stats(st.stats[0:1], st.B)
        ctx.st, ctx.n = st, n
        return (s / n).reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        st = ctx.st
        gy = torch.empty(st.B, st.H, dtype=torch.float32, device=st.p.device)
        check(lib.tg_head_adv_bwd(stream_ptr(), ptr(st.p), ptr(st.wbar), ptr(g.reshape(1).float().contiguous()), ptr(gy),
                                  st.B, st.H, float(ctx.n)), "tg_head_adv_bwd")
        return gy, None


def adv_loss(D, y_last: torch.Tensor) -> torch.Tensor:
    return _AdvLoss.apply(y_last, D)

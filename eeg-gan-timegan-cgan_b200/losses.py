"""Fused TimeGAN losses (host side of csrc/losses.cu) as torch.autograd.Functions.

Mirrors timeGAN/train_timegan.py:70-126:
    recon_loss(x, x_tilde)            tt:72-74    10*sqrt(mean((x-x_tilde)^2)+eps)
    mse_loss(pred, target)            tt:158      mean((pred-target)^2)
    sup_loss_fake(h)                  tt:79-80    mean((h[:,1:]-h[:,:-1])^2)
    bce(p, y)                         tt:70       nn.BCELoss on (B,1) probabilities (tiny; torch ops)
    cov_acf_losses(x_gen, x_real, L)  tt:82-126   channel-covariance Frobenius term and ACF L1 term

Scalars stay on the device; nothing here synchronises with the host.  Under data parallelism the small
sufficient statistics (sums, Gram matrices, lagged products) are all-reduced through `dist.allreduce_stats`
between the kernel stages so every rank evaluates the GLOBAL-batch loss.
"""
import ctypes as _C

import torch
from torch.autograd.function import once_differentiable

from . import dist as _dist
from ._lib import lib, check, ptr, stream_ptr, require_cuda
from . import ops


def _red_ws(device):
    n = lib.tg_reduce_workspace_bytes()
    return torch.empty(n, dtype=torch.uint8, device=device), n


def _sqdiff_sum(a, b):
    out = torch.empty(1, dtype=torch.float32, device=a.device)
    ws, n = _red_ws(a.device)
    check(lib.tg_sqdiff_sum(stream_ptr(), ptr(a), ptr(b), a.numel(), ptr(out), ptr(ws), n), "tg_sqdiff_sum")
    return out


def sumsq(t):
    """sum(t^2) as a device scalar (fixed-order fp64 partials; the R1 norm of train_timegan.py:201)."""
    t = t.contiguous()
    out = torch.empty(1, dtype=torch.float32, device=t.device)
    sizes = (_C.c_longlong * 1)(t.numel())
    ptrs = (_C.c_void_p * 1)(t.data_ptr())
    nbytes = lib.tg_sumsq_workspace_bytes(1, sizes)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=t.device)
    check(lib.tg_sumsq(stream_ptr(), 1, ptrs, sizes, ptr(out), ptr(ws), nbytes), "tg_sumsq")
    return out.reshape(())


class _ReconLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, x_tilde, eps):
        require_cuda(x_tilde, "x_tilde")
        x, x_tilde = x.contiguous(), x_tilde.contiguous()
        sse, cnt = _dist.allreduce_stats(_sqdiff_sum(x, x_tilde), x.numel())
        root = torch.sqrt(sse / cnt + eps)
        ctx.save_for_backward(x, x_tilde, root)
        ctx.cnt = cnt
        return (10.0 * root).reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, x_tilde, root = ctx.saved_tensors
        # d/dx_tilde 10*sqrt(sse/n+eps) = 10 (x_tilde-x) / (n root);  d/dx is the negative
        coef = (g.reshape(1) * 10.0 / (ctx.cnt * root)).contiguous()
        gx = gt = None
        if ctx.needs_input_grad[1]:
            gt = torch.empty_like(x_tilde)
            check(lib.tg_scaled_diff(stream_ptr(), ptr(x_tilde), ptr(x), ptr(coef), ptr(gt), gt.numel(), 0),
                  "tg_scaled_diff")
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            check(lib.tg_scaled_diff(stream_ptr(), ptr(x), ptr(x_tilde), ptr(coef), ptr(gx), gx.numel(), 0),
                  "tg_scaled_diff")
        return gx, gt, None


def recon_loss(x, x_tilde, eps=1e-8):
    return _ReconLoss.apply(x, x_tilde, eps)


class _MseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        require_cuda(pred, "pred")
        pred, target = pred.contiguous(), target.contiguous()
        sse, cnt = _dist.allreduce_stats(_sqdiff_sum(pred, target), pred.numel())
        ctx.save_for_backward(pred, target)
        ctx.cnt = cnt
        return (sse / cnt).reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        coef = (g.reshape(1) * (2.0 / ctx.cnt)).contiguous()
        gp = gt = None
        if ctx.needs_input_grad[0]:
            gp = torch.empty_like(pred)
            check(lib.tg_scaled_diff(stream_ptr(), ptr(pred), ptr(target), ptr(coef), ptr(gp), gp.numel(), 0),
                  "tg_scaled_diff")
        if ctx.needs_input_grad[1]:
            gt = torch.empty_like(target)
            check(lib.tg_scaled_diff(stream_ptr(), ptr(target), ptr(pred), ptr(coef), ptr(gt), gt.numel(), 0),
                  "tg_scaled_diff")
        return gp, gt


def mse_loss(pred, target):
    return _MseLoss.apply(pred, target)


class _Diff1Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h):
        require_cuda(h, "h")
        h = h.contiguous()
        B, T, H = h.shape
        out = torch.empty(1, dtype=torch.float32, device=h.device)
        ws, n = _red_ws(h.device)
        check(lib.tg_diff1_sum(stream_ptr(), ptr(h), B, T, H, ptr(out), ptr(ws), n), "tg_diff1_sum")
        sse, cnt = _dist.allreduce_stats(out, B * (T - 1) * H)
        ctx.save_for_backward(h)
        ctx.cnt = cnt
        return (sse / cnt).reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (h,) = ctx.saved_tensors
        B, T, H = h.shape
        coef = (g.reshape(1) * (2.0 / ctx.cnt)).contiguous()
        gh = torch.empty_like(h)
        check(lib.tg_diff1_grad(stream_ptr(), ptr(h), ptr(coef), ptr(gh), B, T, H, 0), "tg_diff1_grad")
        return gh


def sup_loss_fake(h_fake):
    return _Diff1Loss.apply(h_fake)


def sup_loss(h_real):  # dead code in the reference (tt:76-77), kept for API parity
    return _Diff1Loss.apply(h_real)


def bce(p, y):
    """nn.BCELoss() on (B,1) probabilities: log terms clamped at -100 like ATen. Tiny -> plain torch ops.
    Under DP the mean is over the global batch."""
    lp = torch.clamp(torch.log(p), min=-100.0)
    l1p = torch.clamp(torch.log1p(-p), min=-100.0)
    s = -(y * lp + (1.0 - y) * l1p).sum()
    return _dist.global_mean(s, p.numel())


# ------------------------------------------------------------------------------------------------
# covariance + autocorrelation terms of gen_step
# ------------------------------------------------------------------------------------------------
def _moments_pair(xa, xb):
    """Centred copies and centred Grams (C,C) of TWO (rows,C) matrices (generated and real windows), statistics over
    the global batch.  The two column sums travel in ONE statistics all-reduce, the two Grams in another: under data
    parallelism every call site is a kernel launch plus an NVLink flag round trip on the forward's critical path."""
    rows, Cc = xa.shape
    sums = torch.cat([ops.colsum(xa), ops.colsum(xb)])
    sums, n = _dist.allreduce_stats(sums, rows)
    out = []
    grams = torch.empty(2, Cc, Cc, dtype=torch.float32, device=xa.device)
    for i, x2d in enumerate((xa, xb)):
        mean = (sums[i * Cc:(i + 1) * Cc] / n).contiguous()
        xc = torch.empty_like(x2d)
        check(lib.tg_center_scale(stream_ptr(), ptr(x2d), ptr(mean), None, ptr(xc), rows, Cc), "tg_center_scale")
        ops.wgrad(xc, xc, grams[i], None, Cc)
        out.append(xc)
    grams, _ = _dist.allreduce_stats(grams, rows)
    return out[0], grams[0], out[1], grams[1], n


def _acf_pair(xc_a, gram_a, xc_b, gram_b, n, B, T, Cc, L):
    """Lag-1..L autocorrelation tables of both (centred) tensors; the lagged-product sums share one all-reduce."""
    dev = xc_a.device
    zero = torch.zeros(Cc, dtype=torch.float32, device=dev)
    parts, keep = [], []
    for xc2d, gram in ((xc_a, gram_a), (xc_b, gram_b)):
        std = torch.sqrt(torch.diagonal(gram) / (n - 1))
        inv_s = (1.0 / (std + 1e-8)).contiguous()
        xz = torch.empty_like(xc2d)
        check(lib.tg_center_scale(stream_ptr(), ptr(xc2d), ptr(zero), ptr(inv_s), ptr(xz), B * T, Cc), "tg_center_scale")
        part = torch.empty(B, L * Cc, dtype=torch.float32, device=dev)
        check(lib.tg_acf_fwd(stream_ptr(), ptr(xz), B, T, Cc, L, ptr(part)), "tg_acf_fwd")
        parts.append(ops.colsum(part))
        keep.append((xz, std, inv_s))
    sums, Bg = _dist.allreduce_stats(torch.cat(parts), B)
    lags = torch.arange(1, L + 1, device=dev, dtype=torch.float32)
    denom = (Bg * (T - lags)).unsqueeze(1)                      # (L,1): B*(T-lag) samples per lag
    acf_a = sums[:L * Cc].view(L, Cc) / denom
    acf_b = sums[L * Cc:].view(L, Cc) / denom
    return acf_a, acf_b, keep[0], denom


class _CovAcf(torch.autograd.Function):
    """Returns (cov_term, acf_term) of train_timegan.py:254-263 for generated x_gen against real x_real."""

    @staticmethod
    def forward(ctx, x_gen, x_real, max_lag, need_cov, need_acf):
        require_cuda(x_gen, "x_gen")
        x_gen, x_real = x_gen.contiguous(), x_real.contiguous()
        B, T, Cc = x_gen.shape
        L = max(1, min(int(max_lag), T - 1))
        dev = x_gen.device
        xc_g, gram_g, xc_r, gram_r, n = _moments_pair(x_gen.view(B * T, Cc), x_real.view(B * T, Cc))
        cov_term = torch.zeros((), device=dev)
        acf_term = torch.zeros((), device=dev)
        ctx.need_cov, ctx.need_acf = need_cov, need_acf
        ctx.dims = (B, T, Cc, L)
        ctx.n = n
        saved = [xc_g]
        if need_cov:
            diff = (gram_g - gram_r) / (n - 1)
            fro = torch.sqrt((diff * diff).sum())
            cov_term = fro / (float(Cc * Cc) ** 0.5)
            saved += [diff, fro]
        if need_acf:
            acf_g, acf_r, (xz_g, std_g, inv_s_g), denom = _acf_pair(xc_g, gram_g, xc_r, gram_r, n, B, T, Cc, L)
            d = acf_g - acf_r
            acf_term = d.abs().mean()
            saved += [xz_g, std_g, inv_s_g, torch.sign(d) / denom / float(L * Cc)]
        ctx.save_for_backward(*saved)
        return cov_term, acf_term

    @staticmethod
    @once_differentiable
    def backward(ctx, g_cov, g_acf):
        B, T, Cc, L = ctx.dims
        n = ctx.n
        saved = list(ctx.saved_tensors)
        xc_g = saved.pop(0)
        dx = None
        if ctx.need_cov:
            diff, fro = saved.pop(0), saved.pop(0)
            # d fro/Cc  / d cov = diff / (Cc fro);  d cov / d x = 2 xc G / (n-1)   (G symmetric)
            G = (g_cov * 2.0 / (n - 1)) * diff / (float(Cc) * fro)
            dx = ops.dgrad(xc_g, G.contiguous())
        if ctx.need_acf:
            xz_g, std_g, inv_s_g, S = saved
            S = (S * g_acf).contiguous()
            gz = torch.empty_like(xz_g)
            stat = torch.empty(B, 2 * Cc, dtype=torch.float32, device=xz_g.device)
            check(lib.tg_acf_bwd(stream_ptr(), ptr(xz_g), ptr(S), B, T, Cc, L, ptr(gz), ptr(stat)), "tg_acf_bwd")
            tot, _ = _dist.allreduce_stats(ops.colsum(stat), B)
            mean_gz = (tot[:Cc] / n).contiguous()
            kc = (tot[Cc:] / ((n - 1) * std_g)).contiguous()
            acc = dx is not None
            if dx is None:
                dx = torch.empty_like(xz_g)
            check(lib.tg_acf_bwd_final(stream_ptr(), ptr(gz), ptr(xz_g), ptr(mean_gz), ptr(kc), ptr(inv_s_g), ptr(dx),
                                       B * T, Cc, int(acc)), "tg_acf_bwd_final")
        if dx is None:
            dx = torch.zeros_like(xc_g)
        return dx.view(B, T, Cc), None, None, None, None


def cov_acf_losses(x_gen, x_real, max_lag, need_cov=True, need_acf=True):
    return _CovAcf.apply(x_gen, x_real.detach(), max_lag, need_cov, need_acf)

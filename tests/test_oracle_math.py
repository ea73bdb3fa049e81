"""Pins oracle/gru_math.py (the formulas the CUDA kernels implement) against torch.nn.GRU + autograd,
including the R1 double backward of train_timegan.py:198-202 restated as JVP + reverse-over-tangent."""
import torch
from oracle import gru_math as gm

import pytest


@pytest.fixture(autouse=True)
def _fp64_default():
    """fp64 for these tests only (a module-level set_default_dtype would leak into every other test)."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _mk(I, H, L, seed=0):
    torch.manual_seed(seed)
    g = torch.nn.GRU(I, H, num_layers=L, batch_first=True).double()
    for p in g.parameters():
        torch.nn.init.uniform_(p, -0.6, 0.6)
    return g


def _layer_w(g, l):
    return [getattr(g, f"{n}_l{l}").detach() for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]


def test_forward_and_bptt_match_autograd():
    B, T, I, H, L = 3, 9, 5, 4, 2
    g = _mk(I, H, L)
    x = torch.randn(B, T, I, requires_grad=True)
    y_ref, _ = g(x)
    dy = torch.randn_like(y_ref)
    grads = torch.autograd.grad((y_ref * dy).sum(), [x] + list(g.parameters()))
    # explicit
    acts, saves, inp = [], [], x.detach()
    for l in range(L):
        w_ih, w_hh, b_ih, b_hh = _layer_w(g, l)
        y, sv = gm.gru_layer_fwd(inp, w_ih, w_hh, b_ih, b_hh)
        acts.append((inp, y)); saves.append(sv); inp = y
    assert torch.allclose(inp, y_ref, atol=1e-12)
    d = dy
    out = {}
    for l in reversed(range(L)):
        w_ih, w_hh, _, _ = _layer_w(g, l)
        d, dWi, dWh, dbi, dbh = gm.gru_layer_bwd(d, acts[l][0], acts[l][1], saves[l], w_ih, w_hh)
        out[l] = (dWi, dWh, dbi, dbh)
    assert torch.allclose(d, grads[0], atol=1e-11)
    k = 1
    for l in range(L):
        for a in out[l]:
            assert torch.allclose(a, grads[k], atol=1e-11), (l, k)
            k += 1


def test_tangent_matches_forward_ad():
    import torch.autograd.forward_ad as fwAD
    B, T, I, H, L = 2, 7, 3, 4, 2
    g = _mk(I, H, L, seed=1)
    x = torch.randn(B, T, I); v = torch.randn(B, T, I)
    with fwAD.dual_level():
        yd = fwAD.unpack_dual(g(fwAD.make_dual(x, v))[0]).tangent
    inp, tin = x, v
    for l in range(L):
        w_ih, w_hh, b_ih, b_hh = _layer_w(g, l)
        y, sv = gm.gru_layer_fwd(inp, w_ih, w_hh, b_ih, b_hh)
        tin, _ = gm.gru_layer_jvp(tin, y, sv, w_ih, w_hh)
        inp = y
    assert torch.allclose(tin, yd, atol=1e-12)


def test_r1_gradient_equals_double_backward():
    """d r1/d theta via JVP + reverse-over-tangent == autograd double backward (tt:198-202)."""
    B, T, I, H, L = 3, 8, 4, 5, 2
    g = _mk(I, H, L, seed=2)
    w = torch.randn(1, H, requires_grad=True); b = torch.randn(1, requires_grad=True)
    x = torch.randn(B, T, I, requires_grad=True)

    def disc(xx):
        y, _ = g(xx)
        wn = w / w.norm()
        return torch.sigmoid(y[:, -1] @ wn.T + b)

    d = disc(x)
    grad_real = torch.autograd.grad(d.sum(), x, create_graph=True)[0]
    r1 = grad_real.reshape(B, -1).pow(2).sum(1).mean()
    params = list(g.parameters()) + [w, b]
    ref = torch.autograd.grad(r1, params)

    # explicit path
    with torch.no_grad():
        wn = (w / w.norm())
        acts, saves, inp = [], [], x.detach()
        for l in range(L):
            y, sv = gm.gru_layer_fwd(inp, *_layer_w(g, l))
            acts.append((inp, y)); saves.append(sv); inp = y
        p = torch.sigmoid(inp[:, -1] @ wn.T + b)               # (B,1)
        # v = d sum(p) / dx  (dX-only BPTT)
        dy = torch.zeros_like(inp); dy[:, -1] = (p * (1 - p)) * wn
        dd = dy
        for l in reversed(range(L)):
            w_ih, w_hh, _, _ = _layer_w(g, l)
            dd = gm.gru_layer_bwd(dd, acts[l][0], acts[l][1], saves[l], w_ih, w_hh)[0]
        v = dd
        assert torch.allclose(v, grad_real.detach(), atol=1e-12)
        # tangent forward
        tacts, tsaves, tin = [], [], v
        for l in range(L):
            w_ih, w_hh, _, _ = _layer_w(g, l)
            td, ts = gm.gru_layer_jvp(tin, acts[l][1], saves[l], w_ih, w_hh)
            tacts.append((tin, td)); tsaves.append(ts); tin = td
    # head of sdot with autograd (tiny):  sdot = sum_b p(1-p) * (hdot_last . wn)
    hl = inp[:, -1].clone().requires_grad_(True)
    hdl = tin[:, -1].clone().requires_grad_(True)
    wn_ = w / w.norm()
    pp = torch.sigmoid(hl @ wn_.T + b)
    sdot = (pp * (1 - pp) * (hdl @ wn_.T)).sum()
    assert torch.allclose(sdot / B, r1.detach(), atol=1e-12)
    g_hl, g_hdl, g_w, g_b = torch.autograd.grad(sdot, [hl, hdl, w, b])
    with torch.no_grad():
        hb = torch.zeros_like(inp); hb[:, -1] = g_hl
        hdb = torch.zeros_like(inp); hdb[:, -1] = g_hdl
        got = {}
        for l in reversed(range(L)):
            w_ih, w_hh, _, _ = _layer_w(g, l)
            hb, hdb, dWi, dWh, dbi, dbh = gm.gru_layer_jvp_bwd(
                hb, hdb, acts[l][0], tacts[l][0], acts[l][1], tacts[l][1], saves[l], tsaves[l], w_ih, w_hh)
            got[l] = (dWi, dWh, dbi, dbh)
    k = 0
    for l in range(L):
        for a in got[l]:
            assert torch.allclose(a * (2.0 / B), ref[k], atol=1e-10), (l, k, (a * 2 / B - ref[k]).abs().max())
            k += 1
    assert torch.allclose(g_w * (2.0 / B), ref[k], atol=1e-10)
    assert torch.allclose(g_b * (2.0 / B), ref[k + 1], atol=1e-10)

"""Explicit per-timestep GRU math (forward, BPTT, tangent forward, reverse-over-tangent).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Parity: these formulas
are pinned against torch.nn.GRU + autograd (incl. the R1 double backward of
train_timegan.py:198-202) by tests/test_oracle_math.py; the CUDA kernels in
eeg-gan-timegan-cgan_b200/csrc/ implement exactly these recurrences.

Reference call sites restated here:
  * timeGAN/timegan_model.py:24-34  GRUStack -> torch.nn.GRU(batch_first=True), h0 = 0
  * torch.nn.GRU cell (gate order r,z,n):   n = tanh(W_in x + b_in + r*(W_hn h + b_hn)),
    h' = (1-z)*n + z*h      
  * train_timegan.py:198-202  R1 = mean_b ||d sum(D(x)) / dx||^2 with create_graph=True;
    restated as a JVP so no generic double backward is needed.

All functions take/return torch tensors (any float dtype, CPU) in (B,T,*) batch-first layout.
"""
import torch


def _split3(a, H):
    return a[..., :H], a[..., H:2 * H], a[..., 2 * H:]


def gru_layer_fwd(x, w_ih, w_hh, b_ih, b_hh):
    """One GRU layer.  Returns y (B,T,H) and saved = (r,z,n,q) each (B,T,H), q = h_{t-1} W_hn^T + b_hn."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gi = x @ w_ih.T + b_ih                       # time-batched projection
    h = x.new_zeros(B, H)
    ys, rs, zs, ns, qs = [], [], [], [], []
    for t in range(T):
        gh = h @ w_hh.T + b_hh
        gir, giz, gin = _split3(gi[:, t], H)
        ghr, ghz, q = _split3(gh, H)
        r = torch.sigmoid(gir + ghr)
        z = torch.sigmoid(giz + ghz)
        n = torch.tanh(gin + r * q)
        h = n + z * (h - n)
        ys.append(h); rs.append(r); zs.append(z); ns.append(n); qs.append(q)
    st = lambda l: torch.stack(l, dim=1)
    return st(ys), (st(rs), st(zs), st(ns), st(qs))


def gru_layer_bwd(dy, x, y, saved, w_ih, w_hh):
    """BPTT of one layer.  Returns dx, dW_ih, dW_hh, db_ih, db_hh."""
    B, T, H = y.shape
    r, z, n, q = saved
    dgi = torch.zeros(B, T, 3 * H, dtype=y.dtype)
    dgh = torch.zeros(B, T, 3 * H, dtype=y.dtype)
    carry = y.new_zeros(B, H)
    for t in range(T - 1, -1, -1):
        hp = y[:, t - 1] if t > 0 else y.new_zeros(B, H)
        dh = dy[:, t] + carry
        dn = dh * (1 - z[:, t])
        dz = dh * (hp - n[:, t])
        dan = dn * (1 - n[:, t] ** 2)
        daz = dz * z[:, t] * (1 - z[:, t])
        dr = dan * q[:, t]
        dar = dr * r[:, t] * (1 - r[:, t])
        dgi[:, t] = torch.cat([dar, daz, dan], -1)
        dgh[:, t] = torch.cat([dar, daz, dan * r[:, t]], -1)
        carry = dh * z[:, t] + dgh[:, t] @ w_hh
    hprev = torch.cat([y.new_zeros(B, 1, H), y[:, :-1]], 1)
    dW_hh = dgh.reshape(-1, 3 * H).T @ hprev.reshape(-1, H)
    db_hh = dgh.sum((0, 1))
    dW_ih = dgi.reshape(-1, 3 * H).T @ x.reshape(-1, x.shape[-1])
    db_ih = dgi.sum((0, 1))
    dx = dgi @ w_ih
    return dx, dW_ih, dW_hh, db_ih, db_hh


def gru_layer_jvp(xdot, y, saved, w_ih, w_hh):
    """Tangent forward with fixed weights (A.4).  Returns ydot and tsaved=(ar_dot,az_dot,an_dot,qdot)."""
    B, T, H = y.shape
    r, z, n, q = saved
    gid = xdot @ w_ih.T                          # no bias in the tangent
    hd = y.new_zeros(B, H)
    yd, ard, azd, andd, qd = [], [], [], [], []
    for t in range(T):
        hp = y[:, t - 1] if t > 0 else y.new_zeros(B, H)
        ghd = hd @ w_hh.T
        gidr, gidz, gidn = _split3(gid[:, t], H)
        ghdr, ghdz, qdot = _split3(ghd, H)
        a_r = gidr + ghdr
        a_z = gidz + ghdz
        rdot = r[:, t] * (1 - r[:, t]) * a_r
        zdot = z[:, t] * (1 - z[:, t]) * a_z
        a_n = gidn + rdot * q[:, t] + r[:, t] * qdot
        ndot = (1 - n[:, t] ** 2) * a_n
        hd = (1 - z[:, t]) * ndot + z[:, t] * hd + zdot * (hp - n[:, t])
        yd.append(hd); ard.append(a_r); azd.append(a_z); andd.append(a_n); qd.append(qdot)
    st = lambda l: torch.stack(l, dim=1)
    return st(yd), (st(ard), st(azd), st(andd), st(qd))


def gru_layer_jvp_bwd(hbar, hdbar, x, xdot, y, ydot, saved, tsaved, w_ih, w_hh):
    """Reverse of (primal forward + tangent forward) w.r.t. x, xdot and the weights.

    hbar / hdbar: adjoints flowing into y / ydot, (B,T,H).
    Returns xbar, xdbar, dW_ih, dW_hh, db_ih, db_hh.
    """
    B, T, H = y.shape
    r, z, n, q = saved
    a_r, a_z, a_n, qdot = tsaved
    gib = torch.zeros(B, T, 3 * H, dtype=y.dtype)    # adjoint of primal gi
    ghb = torch.zeros(B, T, 3 * H, dtype=y.dtype)    # adjoint of primal gh
    gidb = torch.zeros(B, T, 3 * H, dtype=y.dtype)   # adjoint of tangent gi
    ghdb = torch.zeros(B, T, 3 * H, dtype=y.dtype)   # adjoint of tangent gh
    ch = y.new_zeros(B, H)
    chd = y.new_zeros(B, H)
    for t in range(T - 1, -1, -1):
        hp = y[:, t - 1] if t > 0 else y.new_zeros(B, H)
        hdp = ydot[:, t - 1] if t > 0 else y.new_zeros(B, H)
        rt, zt, nt, qt = r[:, t], z[:, t], n[:, t], q[:, t]
        art, azt, ant, qdt = a_r[:, t], a_z[:, t], a_n[:, t], qdot[:, t]
        sr = rt * (1 - rt); sz = zt * (1 - zt); sn = 1 - nt * nt
        rdot = sr * art; zdot = sz * azt; ndot = sn * ant
        hb = hbar[:, t] + ch
        hdb = hdbar[:, t] + chd
        # hdot_t = (1-z) ndot + z hdot_{t-1} + zdot (h_{t-1} - n)
        ndb = (1 - zt) * hdb
        zb = hdb * (hdp - ndot)
        zdb = hdb * (hp - nt)
        nb = -zdot * hdb
        ch_next = zdot * hdb
        chd_next = zt * hdb
        # h_t = n + z (h_{t-1} - n)
        nb = nb + (1 - zt) * hb
        zb = zb + hb * (hp - nt)
        ch_next = ch_next + zt * hb
        # ndot = (1-n^2) a_n
        anb_d = sn * ndb
        nb = nb - 2 * nt * ant * ndb
        # a_n(tangent) = gid_n + rdot q + r qdot
        rdb = qt * anb_d
        qb = rdot * anb_d
        rb = qdt * anb_d
        qdb = rt * anb_d
        # n = tanh(gi_n + r q)
        anb = sn * nb
        rb = rb + qt * anb
        qb = qb + rt * anb
        # zdot = sz a_z ; rdot = sr a_r
        azb_d = sz * zdb
        zb = zb + (1 - 2 * zt) * azt * zdb
        arb_d = sr * rdb
        rb = rb + (1 - 2 * rt) * art * rdb
        azb = sz * zb
        arb = sr * rb
        gib[:, t] = torch.cat([arb, azb, anb], -1)
        ghb[:, t] = torch.cat([arb, azb, qb], -1)
        gidb[:, t] = torch.cat([arb_d, azb_d, anb_d], -1)
        ghdb[:, t] = torch.cat([arb_d, azb_d, qdb], -1)
        ch = ch_next + ghb[:, t] @ w_hh
        chd = chd_next + ghdb[:, t] @ w_hh
    zero = y.new_zeros(B, 1, H)
    hprev = torch.cat([zero, y[:, :-1]], 1).reshape(-1, H)
    hdprev = torch.cat([zero, ydot[:, :-1]], 1).reshape(-1, H)
    f = lambda a: a.reshape(-1, a.shape[-1])
    dW_hh = f(ghb).T @ hprev + f(ghdb).T @ hdprev
    db_hh = ghb.sum((0, 1))
    dW_ih = f(gib).T @ f(x) + f(gidb).T @ f(xdot)
    db_ih = gib.sum((0, 1))
    xbar = gib @ w_ih
    xdbar = gidb @ w_ih
    return xbar, xdbar, dW_ih, dW_hh, db_ih, db_hh

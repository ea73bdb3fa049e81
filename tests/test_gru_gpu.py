"""GPU parity of the persistent GRU kernels (forward, BPTT, tangent forward, reverse-over-tangent) through the
C ABI, against torch.nn.GRU + autograd on the CPU in fp32 -- the arithmetic path the reference takes from
timegan_model.py:32-34 -- and against the explicit recurrences of oracle/gru_math.py in fp64.

Tolerance: fp32 forward and gradients within 1e-4 normwise relative (BASELINE.json north_star)."""
import pytest
import torch

from parity_util import relerr, make_gru, flat_weights

pytestmark = pytest.mark.gpu
TOL = 1e-4

SHAPES = [
    # B, T, I, H, L
    (3, 17, 5, 8, 2),
    (1, 1, 14, 24, 1),       # single step
    (5, 33, 14, 6, 2),       # H % 4 != 0 -> generic (non-bulk) streaming path
    (4, 100, 28, 56, 1),     # reference default dims (z=28, h=56, L=1)
    (32, 768, 14, 24, 3),    # config c1 embedder
    (7, 768, 24, 24, 3),     # ragged batch (N % B != 0 tail)
    (9, 200, 14, 64, 3),     # config c2 dims
    (3, 96, 14, 128, 2),     # config c3 dims
    (5, 40, 14, 160, 2),     # H > 128: capacity fallback (W_hh streamed from L2)
    (6, 24, 14, 256, 2),     # config c4's H = 256 point
    (301, 24, 14, 128, 1),   # H = 128 cluster kernels, two sequence groups per cluster (> 148 CTAs otherwise), ragged
    (150, 20, 14, 256, 1),   # H = 256 cluster kernels (8 CTAs), two groups, ragged last cluster
    (5, 33, 128, 128, 2),    # H = 128, fewer sequences than one group + one extra
]


def _ops():
    from timegan_b200 import ops
    return ops


@pytest.mark.parametrize("B,T,I,H,L", SHAPES)
def test_forward_matches_nn_gru(B, T, I, H, L):
    ops = _ops()
    m = make_gru(I, H, L, seed=B + T)
    x = torch.rand(B, T, I, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y_ref, _ = m(x)
    dev = torch.device("cuda:0")
    y, _ = ops.stack_forward(x.to(dev), flat_weights(m, dev), save=False)
    assert relerr(y, y_ref) < TOL
    y2, saves = ops.stack_forward(x.to(dev), flat_weights(m, dev), save=True)
    assert relerr(y2, y_ref) < TOL
    assert len(saves) == L


@pytest.mark.parametrize("B,T,I,H,L", SHAPES)
def test_bptt_matches_autograd(B, T, I, H, L):
    ops = _ops()
    m = make_gru(I, H, L, seed=3 * B + T)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(B, T, I, generator=g, requires_grad=True)
    y_ref, _ = m(x)
    dy = torch.randn(B, T, H, generator=g)
    ref = torch.autograd.grad((y_ref * dy).sum(), [x] + list(m.parameters()))
    dev = torch.device("cuda:0")
    w = flat_weights(m, dev)
    _, saves = ops.stack_forward(x.detach().to(dev), w, save=True)
    dx, grads = ops.stack_backward(dy.to(dev), saves, w, need_dx=True, need_dw=True)
    assert relerr(dx, ref[0]) < TOL
    for k, gk in enumerate(grads):
        assert relerr(gk, ref[1 + k]) < TOL, f"param {k}"


@pytest.mark.parametrize("bt", [1, 2, 4])
@pytest.mark.parametrize("H", [24, 64, 128])
def test_sequences_per_cta_variants(bt, H):
    """Every BT instantiation of the forward/backward kernels gives the same answer (incl. a ragged last CTA)."""
    ops = _ops()
    B, T, I, L = 7, 40, 14, 2
    m = make_gru(I, H, L, seed=H + bt)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, T, I, generator=g, requires_grad=True)
    y_ref, _ = m(x)
    dy = torch.randn(B, T, H, generator=g)
    ref = torch.autograd.grad((y_ref * dy).sum(), [x] + list(m.parameters()))
    dev = torch.device("cuda:0")
    w = flat_weights(m, dev)
    ops.set_bt_override(bt)
    try:
        y, saves = ops.stack_forward(x.detach().to(dev), w, save=True)
        dx, grads = ops.stack_backward(dy.to(dev), saves, w, need_dx=True, need_dw=True)
    finally:
        ops.set_bt_override(0)
    assert relerr(y, y_ref) < TOL
    assert relerr(dx, ref[0]) < TOL
    for k, gk in enumerate(grads):
        assert relerr(gk, ref[1 + k]) < TOL, f"param {k}"


@pytest.mark.parametrize("B,T,H,dy_last", [(7, 40, 128, False), (301, 24, 128, True), (9, 33, 256, False),
                                           (150, 20, 256, True), (4, 1, 128, False)])
def test_cluster_kernels_forced(B, T, H, dy_last):
    """csrc/gru_cluster.cu (W_hh split over the SMs of a thread-block cluster, state all-gathered through distributed
    shared memory) for EVERY H = 128 / 256 launch, whatever the dispatch heuristic would pick: forward and BPTT against
    nn.GRU + autograd, one and two sequence groups per cluster, ragged last cluster, dy_last, T = 1."""
    from timegan_b200._lib import lib
    ops = _ops()
    m = make_gru(H, H, 1, seed=H + B)
    g = torch.Generator().manual_seed(B)
    x = torch.rand(B, T, H, generator=g, requires_grad=True)
    y_ref, _ = m(x)
    if dy_last:
        dy = torch.randn(B, H, generator=g)
        obj = (y_ref[:, -1] * dy).sum()
    else:
        dy = torch.randn(B, T, H, generator=g)
        obj = (y_ref * dy).sum()
    ref = torch.autograd.grad(obj, [x] + list(m.parameters()))
    dev = torch.device("cuda:0")
    w = flat_weights(m, dev)
    assert lib.tg_set_option(b"cluster", 2) == 0
    try:
        y, saves = ops.stack_forward(x.detach().to(dev), w, save=True)
        dx, grads = ops.stack_backward(dy.to(dev), saves, w, need_dx=True, need_dw=True, dy_last=dy_last)
    finally:
        lib.tg_set_option(b"cluster", 1)
    assert relerr(y, y_ref) < TOL
    assert relerr(dx, ref[0]) < TOL
    for k, gk in enumerate(grads):
        assert relerr(gk, ref[1 + k]) < TOL, f"param {k}"


@pytest.mark.parametrize("B,T,I,H,L,dy_last", [(9, 200, 14, 64, 3, False), (6, 64, 24, 24, 3, True), (4, 100, 28, 56, 1, False)])
def test_two_columns_per_thread_bptt_variant(B, T, I, H, L, dy_last):
    """gru_bwd_pair_kernel (one sequence per CTA, a lane group owns two adjacent output columns): same gradients as
    the default kernel's reference.  Off by default (measured slower); kept selectable."""
    from timegan_b200._lib import lib
    ops = _ops()
    m = make_gru(I, H, L, seed=7 * B + H)
    g = torch.Generator().manual_seed(B + 1)
    x = torch.rand(B, T, I, generator=g, requires_grad=True)
    y_ref, _ = m(x)
    dy = torch.randn(B, H, generator=g) if dy_last else torch.randn(B, T, H, generator=g)
    obj = (y_ref[:, -1] * dy).sum() if dy_last else (y_ref * dy).sum()
    ref = torch.autograd.grad(obj, [x] + list(m.parameters()))
    dev = torch.device("cuda:0")
    w = flat_weights(m, dev)
    assert lib.tg_set_option(b"bwd_pair", 1) == 0
    try:
        _, saves = ops.stack_forward(x.detach().to(dev), w, save=True)
        dx, grads = ops.stack_backward(dy.to(dev), saves, w, need_dx=True, need_dw=True, dy_last=dy_last)
    finally:
        lib.tg_set_option(b"bwd_pair", 0)
    assert relerr(dx, ref[0]) < TOL
    for k, gk in enumerate(grads):
        assert relerr(gk, ref[1 + k]) < TOL, f"param {k}"


def test_cluster_capacity_is_reported():
    """tg_cluster_capacity: resident-cluster capacity of the H = 128 / 256 kernels (cudaOccupancyMaxActiveClusters);
    the launchers size the sequence groups per cluster from it so that one wave covers the batch."""
    from timegan_b200._lib import lib
    torch.zeros(1, device="cuda:0")
    sms = lib.tg_device_sm_count()
    for back in (0, 1, 2):
        c = lib.tg_cluster_capacity(128, back, 1)
        assert 1 <= c <= sms // 2, (back, c)
    c8 = lib.tg_cluster_capacity(256, 0, 1)
    assert 1 <= c8 <= sms // 8
    assert lib.tg_cluster_capacity(256, 1, 3) >= 1           # the three-group BPTT instantiation fits (218 KB of smem)
    assert lib.tg_cluster_capacity(256, 2, 1) >= 1           # reverse-over-tangent at H = 256: one group per cluster
    assert lib.tg_cluster_capacity(64, 0, 1) == 0 and lib.tg_cluster_capacity(256, 2, 2) == 0


def test_last_step_only_gradient():
    """Discriminator pattern (timegan_model.py:97): only y[:, -1] feeds the loss."""
    ops = _ops()
    B, T, I, H, L = 6, 64, 24, 24, 3
    m = make_gru(I, H, L, seed=11)
    g = torch.Generator().manual_seed(6)
    x = torch.rand(B, T, I, generator=g, requires_grad=True)
    y_ref, _ = m(x)
    dl = torch.randn(B, H, generator=g)
    ref = torch.autograd.grad((y_ref[:, -1] * dl).sum(), [x] + list(m.parameters()))
    dev = torch.device("cuda:0")
    w = flat_weights(m, dev)
    _, saves = ops.stack_forward(x.detach().to(dev), w, save=True)
    dx, grads = ops.stack_backward(dl.to(dev), saves, w, need_dx=True, need_dw=True, dy_last=True)
    assert relerr(dx, ref[0]) < TOL
    for k, gk in enumerate(grads):
        assert relerr(gk, ref[1 + k]) < TOL, f"param {k}"


def test_autograd_function_module_api():
    """timegan_b200.GRUStack == reference GRUStack (nn.GRU) through torch autograd."""
    import timegan_b200 as tg
    B, T, I, H, L = 4, 50, 14, 24, 3
    m = make_gru(I, H, L, seed=21)
    ours = tg.GRUStack(I, H, L, dropout=0.0)
    ours.rnn.load_state_dict(m.state_dict())
    ours = ours.cuda()
    g = torch.Generator().manual_seed(7)
    x = torch.rand(B, T, I, generator=g)
    xr = x.clone().requires_grad_(True)
    y_ref, _ = m(xr)
    (y_ref ** 2).sum().backward()
    xo = x.clone().cuda().requires_grad_(True)
    y = ours(xo)
    (y ** 2).sum().backward()
    assert relerr(y, y_ref) < TOL
    assert relerr(xo.grad, xr.grad) < TOL
    for (n, p), (_, q) in zip(ours.rnn.named_parameters(), m.named_parameters()):
        assert relerr(p.grad, q.grad) < TOL, n


@pytest.mark.parametrize("B,T,I,H,L", [(3, 12, 5, 8, 2), (4, 64, 24, 24, 3), (2, 30, 14, 64, 1), (2, 20, 14, 128, 2),
                                       (5, 16, 14, 256, 2),
                                       (301, 10, 14, 128, 2)])   # cluster reverse-over-tangent: two groups, ragged
def test_tangent_forward_and_reverse_match_oracle(B, T, I, H, L):
    """R1 building blocks (train_timegan.py:198-202 restated per SURVEY.md A.4) vs oracle/gru_math.py in fp64."""
    from oracle import gru_math as gm
    ops = _ops()
    m = make_gru(I, H, L, seed=31, scale=0.5)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(B, T, I, generator=g)
    v = torch.randn(B, T, I, generator=g)
    hb_last = torch.randn(B, H, generator=g)
    hdb_last = torch.randn(B, H, generator=g)
    # ---- oracle, fp64 ----
    md = m.double()
    lw = lambda l: [getattr(md, f"{n}_l{l}").detach() for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    acts, saves, tacts, tsaves = [], [], [], []
    inp, tin = x.double(), v.double()
    for l in range(L):
        w_ih, w_hh, b_ih, b_hh = lw(l)
        y, sv = gm.gru_layer_fwd(inp, w_ih, w_hh, b_ih, b_hh)
        yd, ts = gm.gru_layer_jvp(tin, y, sv, w_ih, w_hh)
        acts.append((inp, y)); saves.append(sv); tacts.append((tin, yd)); tsaves.append(ts)
        inp, tin = y, yd
    ydot_ref = tin
    hb = torch.zeros(B, T, H, dtype=torch.float64); hb[:, -1] = hb_last.double()
    hdb = torch.zeros(B, T, H, dtype=torch.float64); hdb[:, -1] = hdb_last.double()
    ref = {}
    for l in reversed(range(L)):
        w_ih, w_hh, _, _ = lw(l)
        hb, hdb, dWi, dWh, dbi, dbh = gm.gru_layer_jvp_bwd(hb, hdb, acts[l][0], tacts[l][0], acts[l][1], tacts[l][1],
                                                           saves[l], tsaves[l], w_ih, w_hh)
        ref[l] = (dWi, dWh, dbi, dbh)
    # ---- CUDA ----
    dev = torch.device("cuda:0")
    w = flat_weights(m.float(), dev)
    _, sv_c = ops.stack_forward(x.to(dev), w, save=True)
    ydot, ts_c = ops.stack_jvp_forward(v.to(dev), sv_c, w)
    assert relerr(ydot, ydot_ref) < TOL
    grads = [torch.zeros_like(t) for t in w]
    ops.stack_jvp_backward(hb_last.to(dev), hdb_last.to(dev), sv_c, ts_c, w, grads, accumulate=True)
    for l in range(L):
        for k in range(4):
            assert relerr(grads[4 * l + k], ref[l][k]) < TOL, (l, k)


def test_side_stream_weight_gradients_equal_inline():
    """ops.set_wgrad_overlap(n): the per-layer weight-gradient kernels run on a side stream on at most n SMs.
    Same kernels, same data, fixed-order partial reduction per split count -> results agree to fp32 rounding of the
    different split, and dX (main stream) is bit-identical."""
    ops = _ops()
    B, T, I, H, L = 6, 96, 64, 64, 3
    m = make_gru(I, H, L, seed=11)
    g = torch.Generator().manual_seed(3)
    dev = torch.device("cuda:0")
    x = torch.rand(B, T, I, generator=g).to(dev)
    dy = torch.randn(B, T, H, generator=g).to(dev)
    w = flat_weights(m, dev)
    _, saves = ops.stack_forward(x, w, save=True)
    dx0, g0 = ops.stack_backward(dy, saves, w, need_dx=True, need_dw=True)
    try:
        ops.set_wgrad_overlap(40)
        _, saves = ops.stack_forward(x, w, save=True)
        dx1, g1 = ops.stack_backward(dy, saves, w, need_dx=True, need_dw=True)
        torch.cuda.synchronize()
    finally:
        ops.set_wgrad_overlap(0)
    assert torch.equal(dx0, dx1)
    for a, b in zip(g0, g1):
        assert relerr(b, a) < 1e-5


def test_inter_layer_dropout_masks_forward_and_bptt():
    """nn.GRU(dropout=p) feeds layer l+1 with y_l * mask_l in train mode (timegan_model.py:27-30).  With the masks
    given, forward and BPTT must equal a stack of single-layer torch GRUs with the same masks under autograd."""
    ops = _ops()
    B, T, I, H, L = 5, 48, 14, 24, 3
    m = make_gru(I, H, L, seed=21)
    g = torch.Generator().manual_seed(8)
    x = torch.rand(B, T, I, generator=g, requires_grad=True)
    dy = torch.randn(B, T, H, generator=g)
    masks = [(torch.rand(B, T, H, generator=g) > 0.3).float() / 0.7 for _ in range(L - 1)]
    layers = []
    for l in range(L):
        one = torch.nn.GRU(I if l == 0 else H, H, 1, batch_first=True)
        with torch.no_grad():
            for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                getattr(one, name + "_l0").copy_(getattr(m, f"{name}_l{l}"))
        layers.append(one)
    h = x
    for l, one in enumerate(layers):
        h, _ = one(h)
        if l < L - 1:
            h = h * masks[l]
    params = [p for one in layers for p in (one.weight_ih_l0, one.weight_hh_l0, one.bias_ih_l0, one.bias_hh_l0)]
    ref = torch.autograd.grad((h * dy).sum(), [x] + params)
    dev = torch.device("cuda:0")
    w = flat_weights(m, dev)
    md = [mk.to(dev) for mk in masks]
    y, saves = ops.stack_forward(x.detach().to(dev), w, save=True, masks=md)
    assert relerr(y, h.detach()) < TOL
    dx, grads = ops.stack_backward(dy.to(dev), saves, w, need_dx=True, need_dw=True, masks=md)
    assert relerr(dx, ref[0]) < TOL
    for k, gk in enumerate(grads):
        assert relerr(gk, ref[1 + k]) < TOL, f"param {k}"


@pytest.mark.parametrize("B,T,H,bt", [(5, 37, 64, 0), (3, 50, 64, 2), (9, 21, 64, 4), (3, 40, 128, 0), (5, 19, 128, 2),
                                      (6, 13, 128, 4), (256, 768, 64, 0), (256, 768, 128, 0), (512, 768, 128, 0)])
@pytest.mark.parametrize("save", [True, False])
def test_forward_from_bf16_gi_is_the_fp32_kernel_on_the_widened_input(B, T, H, bt, save):
    """tg_gru_fwd_bf16gi (reads the bf16 projection of tg_proj_bf16, saves r,z,n to their own fp32 tensor) does the same
    arithmetic as tg_gru_fwd on gi16.float(): y, r,z,n and q must come out BIT-identical."""
    from timegan_b200 import _lib
    from timegan_b200._lib import lib, check, ptr, stream_ptr
    DEV = "cuda:0"
    g = torch.Generator().manual_seed(B * 1000 + T + H)
    gi16 = (torch.randn(B, T, 3 * H, generator=g)).to(DEV).to(torch.bfloat16)
    whh = (torch.randn(3 * H, H, generator=g) / H ** 0.5).to(DEV)
    bhh = (torch.randn(3 * H, generator=g) * 0.1).to(DEV)
    flags = (_lib.GRU_SAVE if save else 0) | (bt << 8)
    y0 = torch.empty(B, T, H, device=DEV); q0 = torch.empty(B, T, H, device=DEV) if save else None
    gi = gi16.float().contiguous()
    check(lib.tg_gru_fwd(stream_ptr(), ptr(gi), ptr(whh), ptr(bhh), ptr(y0), ptr(q0), B, T, H, flags), "tg_gru_fwd")
    y1 = torch.full((B, T, H), float("nan"), device=DEV)
    q1 = torch.full((B, T, H), float("nan"), device=DEV) if save else None
    rzn = torch.full((B, T, 3 * H), float("nan"), device=DEV) if save else None
    check(lib.tg_gru_fwd_bf16gi(stream_ptr(), ptr(gi16), ptr(whh), ptr(bhh), ptr(y1), ptr(q1), ptr(rzn), B, T, H, flags),
          "tg_gru_fwd_bf16gi")
    torch.cuda.synchronize()
    assert torch.equal(y1, y0)
    if save:
        assert torch.equal(q1, q0) and torch.equal(rzn, gi)


def test_bf16_gi_unsupported_hidden_size_is_refused():
    from timegan_b200._lib import lib, ptr, stream_ptr
    assert lib.tg_bf16_gi_supported(8, 512, 64, 64) == 1 and lib.tg_bf16_gi_supported(8, 512, 128, 128) == 1
    assert lib.tg_bf16_gi_supported(8, 512, 24, 24) == 0 and lib.tg_bf16_gi_supported(2, 32, 64, 64) == 0
    assert lib.tg_bf16_gi_supported(512, 768, 128, 128) == 0     # the cluster forward kernel's batch: fp32 gi
    t = torch.zeros(2, 4, 72, device="cuda:0")
    rc = lib.tg_gru_fwd_bf16gi(stream_ptr(), ptr(t), ptr(t), ptr(t), ptr(t), None, None, 2, 4, 24, 0)
    assert rc == -4

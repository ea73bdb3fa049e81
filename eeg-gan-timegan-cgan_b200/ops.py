"""torch.autograd.Function wrappers around the C ABI (the thin host layer the north_star asks for).

Low-level "stack" helpers run an L-layer GRU stack layer by layer:
    projection GEMM (time-batched X W_ih^T + b_ih)  ->  persistent recurrent kernel
and the matching backward / tangent-forward / reverse-over-tangent passes.  They are used directly by the
fused training steps (train_timegan.disc_step needs the R1 passes) and wrapped in GRUStackFunction for the
module API of timegan_model.py.

Reference behaviour being reproduced: timeGAN/timegan_model.py:24-34 (GRUStack -> nn.GRU(batch_first=True),
h0 = 0, returns y only), SURVEY.md Appendix A.1/A.2/A.4 for the math.
"""
from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import lib, check, ptr, stream_ptr, require_cuda

# process-wide projection mode for the GRU input projections ("fp32" exact, "bf16" tensor-core mode)
_PROJ_MODE = _lib.PROJ_TF32X3
_BT_OVERRIDE = 0  # sequences per CTA override for the recurrent kernels (0 = heuristic)


_MODES = {"ffma": _lib.PROJ_FP32, "fp32": _lib.PROJ_TF32X3, "bf16": _lib.PROJ_TF32, "tf32": _lib.PROJ_TF32,
          "tf32x3": _lib.PROJ_TF32X3}
_MODE_NAME = "fp32"
# "bf16" mode: the GRU input projections X W_ih^T run with bf16 operands (tcgen05 kind::f16) and store a bf16 result that
# the forward recurrence reads directly (tg_proj_bf16 / tg_gru_fwd_bf16gi); every other contraction of the step (dX,
# weight gradients) takes one TF32 pass, the heads and the R1 tangent projections stay in fp32-parity precision.
_BF16_GI = False


def set_proj_mode(mode: str):
    """Precision of the time-batched contractions (input projections, dX, heads):
    "ffma"   exact fp32 on CUDA cores;
    "tf32x3" tcgen05 tensor cores, 3xTF32 split (fp32-parity mode, 1e-4);
    "fp32"   alias of the fp32-parity mode in use (see _MODES);
    "tf32"   tcgen05 tensor cores, one TF32 pass over the fp32 operands (reduced precision, 2e-2);
    "bf16"   BASELINE config c3's "bf16 input projections": bf16 operands and a bf16 gi tensor for the input
             projections (layers with H = 64 / 128; others fall back to "tf32"), one TF32 pass elsewhere (2e-2)."""
    global _PROJ_MODE, _MODE_NAME, _BF16_GI
    if mode not in _MODES:
        raise ValueError(f"proj mode must be one of {sorted(_MODES)}")
    _PROJ_MODE, _MODE_NAME, _BF16_GI = _MODES[mode], mode, mode == "bf16"


def get_proj_mode() -> str:
    return _MODE_NAME


def set_bt_override(bt: int):
    global _BT_OVERRIDE
    if bt not in (0, 1, 2, 4):
        raise ValueError("bt override must be 0, 1, 2 or 4")
    _BT_OVERRIDE = bt


# Sequences-per-CTA hint for passes that run BESIDE another recurrent pass (train_timegan._Fork regions): two concurrent
# B = 256 launches with one sequence per CTA need 512 CTA slots where the SMs have 444, so the second launch trails the
# first; with two sequences per CTA both fit (tools/probe_concurrency.py: 395 us for both against 462 us back to back).
# 0 = no hint.  Experimental: enabled by TIMEGAN_B200_FORK_BT=2.
_BT_HINT = 0
_FORK_BT = int(__import__("os").environ.get("TIMEGAN_B200_FORK_BT", "0"))


def set_fork_bt(bt: int):
    global _FORK_BT
    _FORK_BT = int(bt)


def fork_begin():
    global _BT_HINT
    _BT_HINT = _FORK_BT


def fork_end():
    global _BT_HINT
    _BT_HINT = 0


def _flags(base: int = 0, hint: int = None) -> int:
    bt = _BT_OVERRIDE or (_BT_HINT if hint is None else hint)
    return base | (bt << 8)


# Weight gradients off the critical path: nothing downstream of BPTT needs dW until clip + Adam, so the fused
# weight-gradient kernel of layer l can run while layer l-1's BPTT is already going -- on a side stream and on a
# CAPPED number of SMs (it is a one-CTA-per-SM persistent kernel that owns the whole shared memory of its SM, so
# an uncapped launch would simply push the recurrent kernel out).  0 = in line on the calling stream.
_WGRAD_CTAS = 0
_WGRAD_STREAMS = {}


def set_wgrad_overlap(n_ctas: int):
    """n_ctas > 0: run the per-layer weight-gradient kernels on a side stream using at most n_ctas SMs."""
    global _WGRAD_CTAS
    _WGRAD_CTAS = max(0, int(n_ctas))
    check(lib.tg_set_option(b"wgrad_ctas", _WGRAD_CTAS), "tg_set_option")


class _WgradSide:
    """Per stack-backward helper: issue(fn, *keep) runs fn on the side stream after everything issued so far on
    the calling stream; join() makes the calling stream wait for it.  Tensors the side stream reads are kept
    referenced until join(), so the caching allocator cannot hand their blocks to later main-stream allocations."""

    def __init__(self, device):
        self.on = _WGRAD_CTAS > 0
        self.keep = []
        if self.on:
            self.main = torch.cuda.current_stream(device)
            key = (str(device), self.main.cuda_stream)
            if key not in _WGRAD_STREAMS:
                _WGRAD_STREAMS[key] = torch.cuda.Stream(device=device)
            self.side = _WGRAD_STREAMS[key]
            self.used = False

    def issue(self, fn, *keep):
        if not self.on:
            fn()
            return
        self.keep.extend(keep)
        self.side.wait_stream(self.main)
        with torch.cuda.stream(self.side):
            fn()
        self.used = True

    def join(self):
        if self.on and self.used:
            self.main.wait_stream(self.side)
        self.keep.clear()


class LayerSave:
    """Activations one layer keeps for its backward pass (all (B,T,*) fp32, contiguous)."""
    __slots__ = ("inp", "rzn", "q", "y", "ipad")

    def __init__(self, inp, rzn, q, y, ipad: int = 0):
        # ipad > 0: `inp` carries that many zero columns on the right (see pad_cols)
        self.inp, self.rzn, self.q, self.y, self.ipad = inp, rzn, q, y, ipad

    def narrow(self, start: int, length: int) -> "LayerSave":
        """Batch slice [start, start+length) as views (contiguous: the batch is the leading dimension)."""
        f = lambda t: t.narrow(0, start, length)
        return LayerSave(f(self.inp), f(self.rzn), f(self.q), f(self.y), self.ipad)


def pad_cols(t: torch.Tensor, mult: int = 4) -> Tuple[torch.Tensor, int]:
    """Zero-pad the last dimension to a multiple of `mult` floats.  The tensor-core tiles are fed by TMA, whose rows
    must be 16-byte multiples: the 14 EEG channels (a 56-byte pitch) are carried as 16 -- two zero columns that add
    nothing to any contraction -- instead of sending every 14-wide GEMM to the FFMA fallback."""
    k = t.shape[-1]
    pad = (-k) % mult
    if pad == 0:
        return t, 0
    return torch.nn.functional.pad(t, (0, pad)), pad


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# raw GEMM helpers
# ------------------------------------------------------------------------------------------------
def proj(a2d: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out2d: Optional[torch.Tensor] = None,
         mode: Optional[int] = None, accumulate: bool = False) -> torch.Tensor:
    """out[M,N] (+)= a[M,K] @ w[N,K]^T + bias."""
    M, K = a2d.shape
    N = w.shape[0]
    if out2d is None:
        out2d = torch.empty(M, N, dtype=torch.float32, device=a2d.device)
    check(lib.tg_proj(stream_ptr(), ptr(a2d), a2d.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(out2d),
                      out2d.stride(0), M, N, K, int(accumulate), _PROJ_MODE if mode is None else mode), "tg_proj")
    return out2d


def dgrad(dg2d: torch.Tensor, w: torch.Tensor, out2d: Optional[torch.Tensor] = None,
          accumulate: bool = False) -> torch.Tensor:
    """out[M,N] (+)= dg[M,K] @ w[K,N]."""
    M, K = dg2d.shape
    N = w.shape[1]
    if out2d is None:
        out2d = torch.empty(M, N, dtype=torch.float32, device=dg2d.device)
    if _PROJ_MODE != _lib.PROJ_FP32:
        # tensor-core path: the same TN kernel as the projection, with the (small) weight transposed to K-major
        wt = w.t().contiguous()
        check(lib.tg_proj(stream_ptr(), ptr(dg2d), dg2d.stride(0), ptr(wt), wt.stride(0), None, ptr(out2d),
                          out2d.stride(0), M, N, K, int(accumulate), _PROJ_MODE), "tg_proj(dgrad)")
        return out2d
    check(lib.tg_dgrad(stream_ptr(), ptr(dg2d), dg2d.stride(0), ptr(w), w.stride(0), ptr(out2d), out2d.stride(0),
                       M, N, K, int(accumulate)), "tg_dgrad")
    return out2d


def wgrad(dg2d: torch.Tensor, a2d: torch.Tensor, dw: torch.Tensor, db: Optional[torch.Tensor], n_cols: int,
          shift_T: int = 0, accumulate: bool = False, mode: Optional[int] = None):
    """dw[N,K] (+)= dg[:, :N]^T @ a ; db[N] (+)= colsum(dg[:, :N]).  shift_T>0: a row m := a[m-1], 0 at m%T==0."""
    M = dg2d.shape[0]
    K = a2d.shape[1]
    nbytes = lib.tg_wgrad_workspace_bytes(M, n_cols, K)
    ws = _ws(nbytes, dg2d.device)
    check(lib.tg_wgrad(stream_ptr(), ptr(dg2d), dg2d.stride(0), ptr(a2d), a2d.stride(0), ptr(dw), dw.stride(0),
                       ptr(db), M, n_cols, K, shift_T, int(accumulate), ptr(ws), nbytes,
                       _PROJ_MODE if mode is None else mode), "tg_wgrad")


def wgrad_gru(dgi: torch.Tensor, dq: torch.Tensor, x2d: Optional[torch.Tensor], y: torch.Tensor, g_wih, g_whh,
              g_bih, g_bhh, accumulate: bool = False, mode: Optional[int] = None):
    """All weight gradients of one GRU layer from the BPTT outputs dgi (B,T,3H), dq (B,T,H), the layer input
    x2d (B*T,I) (None: skip dW_ih) and the layer output y (B,T,H).  Biases may be None (tangent path)."""
    B, T, H = y.shape
    I = x2d.shape[1] if x2d is not None else 0
    if x2d is not None and (x2d.stride(1) != 1):
        x2d = x2d.contiguous()
    nbytes = lib.tg_wgrad_gru_workspace_bytes(B, T, max(I, 1), H)
    ws = _ws(nbytes, y.device)
    check(lib.tg_wgrad_gru(stream_ptr(), ptr(dgi), ptr(dq), ptr(x2d), x2d.stride(0) if x2d is not None else 0, ptr(y),
                           ptr(g_wih) if x2d is not None else None, ptr(g_whh), ptr(g_bih), ptr(g_bhh), B, T, I, H,
                           int(accumulate), ptr(ws), nbytes, _PROJ_MODE if mode is None else mode), "tg_wgrad_gru")


def colsum(x2d: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    M, N = x2d.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=x2d.device)
    nbytes = lib.tg_colsum_workspace_bytes(N)
    ws = _ws(nbytes, x2d.device)
    check(lib.tg_colsum(stream_ptr(), ptr(x2d), x2d.stride(0), M, N, ptr(out), int(accumulate), ptr(ws), nbytes),
          "tg_colsum")
    return out


# ------------------------------------------------------------------------------------------------
# GRU stack passes
# ------------------------------------------------------------------------------------------------
def _layer_weights(weights: Sequence[torch.Tensor], l: int):
    return weights[4 * l], weights[4 * l + 1], weights[4 * l + 2], weights[4 * l + 3]


def stack_forward(x: torch.Tensor, weights: Sequence[torch.Tensor], save: bool,
                  masks: Optional[Sequence[torch.Tensor]] = None) -> Tuple[torch.Tensor, List[LayerSave]]:
    """Run the L-layer stack.  weights = [w_ih, w_hh, b_ih, b_hh] * L.  Returns (y_last_layer, saves).

    masks (optional, L-1 tensors (B,T,H) already scaled by 1/(1-p)): inter-layer dropout of
    nn.GRU(dropout=p) in train mode (timegan_model.py:27-30) -- layer l+1 reads y_l * masks[l]."""
    require_cuda(x, "GRU input")
    x = x.contiguous()
    B, T, _ = x.shape
    L = len(weights) // 4
    saves: List[LayerSave] = []
    inp = x
    for l in range(L):
        w_ih, w_hh, b_ih, b_hh = _layer_weights(weights, l)
        H = w_hh.shape[1]
        ipad = 0
        if inp.shape[-1] % 4 != 0 and _PROJ_MODE != _lib.PROJ_FP32 and B * T >= 128:
            inp, ipad = pad_cols(inp)               # (B,T,14) -> (B,T,16): layer 0 of the embedder
            w_ih, _ = pad_cols(w_ih)
        y = torch.empty(B, T, H, dtype=torch.float32, device=x.device)
        q = torch.empty(B, T, H, dtype=torch.float32, device=x.device) if save else None
        K = inp.shape[-1]
        if _BF16_GI and lib.tg_bf16_gi_supported(B, T, K, H):
            # bf16 input projection: bf16 operands, bf16 gi (half the bytes of the layer's largest tensor, out of the
            # projection and into the recurrence); r,z,n are saved in fp32 in their own tensor
            gi16 = torch.empty(B, T, 3 * H, dtype=torch.bfloat16, device=x.device)
            w16 = w_ih.to(torch.bfloat16)
            a2 = inp.view(B * T, K)
            check(lib.tg_proj_bf16(stream_ptr(), ptr(a2), a2.stride(0), ptr(w16), w16.stride(0), ptr(b_ih), ptr(gi16),
                                   3 * H, B * T, 3 * H, K), "tg_proj_bf16")
            gi = torch.empty(B, T, 3 * H, dtype=torch.float32, device=x.device) if save else None
            check(lib.tg_gru_fwd_bf16gi(stream_ptr(), ptr(gi16), ptr(w_hh), ptr(b_hh), ptr(y), ptr(q), ptr(gi), B, T, H,
                                        _flags(_lib.GRU_SAVE if save else 0)), "tg_gru_fwd_bf16gi")
        else:
            gi = torch.empty(B, T, 3 * H, dtype=torch.float32, device=x.device)
            proj(inp.view(B * T, -1), w_ih, b_ih, gi.view(B * T, 3 * H))
            check(lib.tg_gru_fwd(stream_ptr(), ptr(gi), ptr(w_hh), ptr(b_hh), ptr(y), ptr(q), B, T, H,
                                 _flags(_lib.GRU_SAVE if save else 0)), "tg_gru_fwd")
        if save:
            saves.append(LayerSave(inp, gi, q, y, ipad))
        inp = y
        if masks is not None and l < L - 1:
            inp = y * masks[l]
    return inp, saves


def stack_backward(dy: torch.Tensor, saves: List[LayerSave], weights: Sequence[torch.Tensor], need_dx: bool,
                   need_dw: bool, dy_last: bool = False, grads: Optional[List[torch.Tensor]] = None,
                   accumulate: bool = False, masks: Optional[Sequence[torch.Tensor]] = None):
    """BPTT through the stack.  dy: (B,T,H) or (B,H) if dy_last.  Returns (dx or None, grads list like weights).

    grads (optional) are pre-allocated tensors shaped like `weights` to write (or accumulate) into.
    """
    L = len(saves)
    B, T, _ = saves[0].y.shape
    dev = saves[0].y.device
    if need_dw and grads is None:
        grads = alloc_like_flat(weights)
        accumulate = False
    d = dy.contiguous()
    last_only = dy_last
    dx = None
    side = _WgradSide(dev)
    for l in reversed(range(L)):
        sv = saves[l]
        w_ih, w_hh, _, _ = _layer_weights(weights, l)
        H = w_hh.shape[1]
        I = w_ih.shape[1]
        Ip = I + sv.ipad              # width of the saved (zero-padded) layer input
        dgi = torch.empty(B, T, 3 * H, dtype=torch.float32, device=dev)
        dq = torch.empty(B, T, H, dtype=torch.float32, device=dev)
        w_hh_t = w_hh.t().contiguous() if H > 128 else None      # the H > 128 fallback walks W_hh^T rows
        check(lib.tg_gru_bwd(stream_ptr(), ptr(d), ptr(sv.rzn), ptr(sv.q), ptr(sv.y), ptr(w_hh), ptr(dgi), ptr(dq),
                             B, T, H, _flags(_lib.GRU_DY_LAST if last_only else 0), ptr(w_hh_t)), "tg_gru_bwd")
        last_only = False
        dgi2 = dgi.view(B * T, 3 * H)
        if l > 0 or need_dx:          # the critical path first: dX feeds the next layer's BPTT
            if sv.ipad:
                dxp = torch.empty(B, T, Ip, dtype=torch.float32, device=dev)
                dgrad(dgi2, pad_cols(w_ih)[0], dxp.view(B * T, Ip))
                dx = dxp[..., :I].contiguous()
            else:
                dx = torch.empty(B, T, I, dtype=torch.float32, device=dev)
                dgrad(dgi2, w_ih, dx.view(B * T, I))
            d = dx
            if masks is not None and l > 0:
                d = dx * masks[l - 1]
        if need_dw:
            g_wih, g_whh, g_bih, g_bhh = _layer_weights(grads, l)
            x2 = sv.inp.reshape(B * T, Ip)
            if sv.ipad:
                def _wg_padded(dgi=dgi, dq=dq, x2=x2, sv=sv, g_wih=g_wih, g_whh=g_whh, g_bih=g_bih, g_bhh=g_bhh):
                    # dW_ih against the padded input (the two extra columns come out as exact zeros and are dropped);
                    # with accumulate the kernel adds onto its output, so the scratch starts at zero
                    alloc = torch.zeros if accumulate else torch.empty
                    tmp = alloc(g_wih.shape[0], x2.shape[1], dtype=torch.float32, device=dev)
                    wgrad_gru(dgi, dq, x2, sv.y, tmp, g_whh, g_bih, g_bhh, accumulate)
                    if accumulate:
                        g_wih.add_(tmp[:, :g_wih.shape[1]])
                    else:
                        g_wih.copy_(tmp[:, :g_wih.shape[1]])
                side.issue(_wg_padded, dgi, dq, x2)
            else:
                side.issue(lambda: wgrad_gru(dgi, dq, x2, sv.y, g_wih, g_whh, g_bih, g_bhh, accumulate), dgi, dq, x2)
    side.join()
    return (dx if need_dx else None), grads


class TangentSave:
    __slots__ = ("xdot", "ta", "qdot", "ydot")

    def __init__(self, xdot, ta, qdot, ydot):
        self.xdot, self.ta, self.qdot, self.ydot = xdot, ta, qdot, ydot


def stack_jvp_forward(xdot: torch.Tensor, saves: List[LayerSave], weights: Sequence[torch.Tensor],
                      masks: Optional[Sequence[torch.Tensor]] = None):
    """Tangent forward with fixed weights.  Returns (ydot_last_layer, tangent saves)."""
    L = len(saves)
    B, T, _ = saves[0].y.shape
    dev = xdot.device
    tin = xdot.contiguous()
    tsaves: List[TangentSave] = []
    for l in range(L):
        sv = saves[l]
        w_ih, w_hh, _, _ = _layer_weights(weights, l)
        H = w_hh.shape[1]
        gid = torch.empty(B, T, 3 * H, dtype=torch.float32, device=dev)
        # the R1 tangent keeps fp32-parity precision in the reduced-precision modes too (3xTF32 on the tensor pipe;
        # it used to fall back to the FFMA kernel there: 424 vs 140 us per call at the c3 shape)
        proj(tin.view(B * T, -1), w_ih, None, gid.view(B * T, 3 * H),
             mode=_lib.PROJ_TF32X3 if _PROJ_MODE == _lib.PROJ_TF32 else _PROJ_MODE)
        ydot = torch.empty(B, T, H, dtype=torch.float32, device=dev)
        qdot = torch.empty(B, T, H, dtype=torch.float32, device=dev)
        check(lib.tg_gru_jvp_fwd(stream_ptr(), ptr(gid), ptr(sv.rzn), ptr(sv.q), ptr(sv.y), ptr(w_hh), ptr(ydot),
                                 ptr(qdot), B, T, H, _flags()), "tg_gru_jvp_fwd")
        tsaves.append(TangentSave(tin, gid, qdot, ydot))
        tin = ydot
        if masks is not None and l < L - 1:
            tin = ydot * masks[l]
    return tin, tsaves


def stack_jvp_backward(hbar_last: torch.Tensor, hdbar_last: torch.Tensor, saves: List[LayerSave],
                       tsaves: List[TangentSave], weights: Sequence[torch.Tensor], grads: List[torch.Tensor],
                       accumulate: bool, masks: Optional[Sequence[torch.Tensor]] = None):
    """Reverse over (primal + tangent) forward.  hbar_last / hdbar_last: (B,H) adjoints of the last step of the
    top layer's y / ydot.  Accumulates weight gradients into `grads`; input adjoints are not needed (R1: the
    stack input and its tangent are constants)."""
    L = len(saves)
    B, T, _ = saves[0].y.shape
    dev = hbar_last.device
    hb, hdb = hbar_last.contiguous(), hdbar_last.contiguous()
    last_only = True
    side = _WgradSide(dev)
    for l in reversed(range(L)):
        sv, ts = saves[l], tsaves[l]
        w_ih, w_hh, _, _ = _layer_weights(weights, l)
        H = w_hh.shape[1]
        I = w_ih.shape[1]
        gib = torch.empty(B, T, 3 * H, dtype=torch.float32, device=dev)
        gidb = torch.empty(B, T, 3 * H, dtype=torch.float32, device=dev)
        qb = torch.empty(B, T, H, dtype=torch.float32, device=dev)
        qdb = torch.empty(B, T, H, dtype=torch.float32, device=dev)
        w_hh_t = w_hh.t().contiguous() if H > 128 else None
        check(lib.tg_gru_jvp_bwd(stream_ptr(), ptr(hb), ptr(hdb), ptr(sv.rzn), ptr(sv.q), ptr(ts.ta), ptr(ts.qdot),
                                 ptr(sv.y), ptr(ts.ydot), ptr(w_hh), ptr(gib), ptr(qb), ptr(gidb), ptr(qdb), B, T, H,
                                 _flags(_lib.GRU_DY_LAST if last_only else 0), ptr(w_hh_t)), "tg_gru_jvp_bwd")
        last_only = False
        g_wih, g_whh, g_bih, g_bhh = _layer_weights(grads, l)
        gib2, gidb2 = gib.view(B * T, 3 * H), gidb.view(B * T, 3 * H)
        y2, yd2 = sv.y.view(B * T, H), ts.ydot.view(B * T, H)
        if l > 0:                     # the critical path first
            hb = torch.empty(B, T, I, dtype=torch.float32, device=dev)
            hdb = torch.empty(B, T, I, dtype=torch.float32, device=dev)
            dgrad(gib2, w_ih, hb.view(B * T, I))
            dgrad(gidb2, w_ih, hdb.view(B * T, I))
            if masks is not None:
                hb, hdb = hb * masks[l - 1], hdb * masks[l - 1]

        def _wg(gib=gib, qb=qb, gidb=gidb, qdb=qdb, sv=sv, ts=ts, g_wih=g_wih, g_whh=g_whh, g_bih=g_bih,
                g_bhh=g_bhh, I=I):
            # primal path, then tangent path (no bias terms in the tangent)
            x2 = sv.inp.reshape(B * T, I + sv.ipad)
            if sv.ipad:                      # a layer input that stack_forward carried zero-padded (pad_cols)
                x2 = x2[:, :I].contiguous()
            wgrad_gru(gib, qb, x2, sv.y, g_wih, g_whh, g_bih, g_bhh, accumulate)
            wgrad_gru(gidb, qdb, ts.xdot.reshape(B * T, I), ts.ydot, g_wih, g_whh, None, None, True)
        side.issue(_wg, gib, qb, gidb, qdb)
    side.join()


class FlatList(list):
    """List of views carved out of ONE allocation (`.flat`): whole-set operations are a single launch."""
    flat: torch.Tensor = None


def alloc_like_flat(tensors: Sequence[torch.Tensor]) -> "FlatList":
    """One allocation carved into views shaped like `tensors` (weight-gradient buffers of a stack)."""
    total = sum(t.numel() for t in tensors)
    flat = torch.empty(total, dtype=torch.float32, device=tensors[0].device)
    out, o = FlatList(), 0
    for t in tensors:
        n = t.numel()
        out.append(flat[o:o + n].view(t.shape))
        o += n
    out.flat = flat
    return out


# ------------------------------------------------------------------------------------------------
# autograd Functions (module API)
# ------------------------------------------------------------------------------------------------
class GRUStackFunction(torch.autograd.Function):
    """y = GRUStack(x) for an L-layer stack (timegan_model.py:32-34).

    last_only: return only y[:, -1, :] (the discriminator head reads nothing else, timegan_model.py:97); the
    backward then seeds BPTT at t = T-1 instead of streaming a (B,T,H) gradient that is zero everywhere else.
    masks: None or a tuple of L-1 scaled inter-layer dropout masks (constants)."""

    @staticmethod
    def forward(ctx, x, last_only, masks, *weights):
        need = any(ctx.needs_input_grad)
        wd = [w.detach() for w in weights]
        y, saves = stack_forward(x, wd, save=need, masks=masks)
        ctx.saves, ctx.weights, ctx.masks, ctx.last_only = saves, wd, masks, bool(last_only)
        ctx.bt_hint = _BT_HINT
        return y[:, -1, :].contiguous() if last_only else y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        need_dx = ctx.needs_input_grad[0]
        need_dw = any(ctx.needs_input_grad[3:])
        global _BT_HINT
        old, _BT_HINT = _BT_HINT, ctx.bt_hint          # the BPTT of a forked forward runs beside its sibling's BPTT too
        try:
            dx, grads = stack_backward(dy, ctx.saves, ctx.weights, need_dx, need_dw, dy_last=ctx.last_only,
                                       masks=ctx.masks)
        finally:
            _BT_HINT = old
        ctx.saves = None
        if not need_dw:
            grads = [None] * len(ctx.weights)
        else:
            grads = [g if n else None for g, n in zip(grads, ctx.needs_input_grad[3:])]
        return (dx, None, None, *grads)


class LinearFunction(torch.autograd.Function):
    """y = x W^T + b over the last dim (timegan_model.py:53 Recovery.out; :66/:79 G/S proj when h != z).

    A head whose width is not a multiple of 4 floats (Recovery.out: 14 channels) is evaluated 16 wide -- W and b
    zero-padded, the two extra output columns dropped -- so that forward, dX and dW all run on the TMA-fed
    tensor-core tiles (3xTF32, fp32 parity) instead of the FFMA fallback."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        require_cuda(x, "Linear input")
        x = x.contiguous()
        K = x.shape[-1]
        N = weight.shape[0]
        x2 = x.view(-1, K)
        tc = _PROJ_MODE != _lib.PROJ_FP32 and x2.shape[0] >= 128 and K % 4 == 0
        mode = _lib.PROJ_TF32X3 if tc else _lib.PROJ_FP32     # heads always in fp32-parity precision
        w, b = weight.detach(), (None if bias is None else bias.detach())
        npad = (-N) % 4 if tc else 0
        if npad:
            w = torch.nn.functional.pad(w, (0, 0, 0, npad))
            b = None if b is None else torch.nn.functional.pad(b, (0, npad))
        out = proj(x2, w, b, mode=mode)
        ctx.save_for_backward(x2, weight)
        ctx.has_bias, ctx.npad, ctx.mode = bias is not None, npad, mode
        out = out.view(*x.shape[:-1], N + npad)
        return out[..., :N] if npad else out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x2, weight = ctx.saved_tensors
        N, K = weight.shape
        npad = ctx.npad
        d2 = dout.reshape(-1, N)
        if npad:
            d2 = torch.nn.functional.pad(d2, (0, npad))                  # (M, N) -> (M, N+npad), zeros on the right
            wp = torch.nn.functional.pad(weight.detach(), (0, 0, 0, npad))
        else:
            d2 = d2.contiguous()
            wp = weight.detach()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = dgrad(d2, wp).view(*dout.shape[:-1], K)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dwp = torch.empty(N + npad, K, dtype=torch.float32, device=weight.device)
            dbp = torch.empty(N + npad, dtype=torch.float32, device=weight.device) if ctx.has_bias else None
            wgrad(d2, x2, dwp, dbp, N + npad, mode=ctx.mode)
            dw = dwp[:N].contiguous() if npad else dwp
            db = None if dbp is None else (dbp[:N].contiguous() if npad else dbp)
        return dx, dw, db


def gru_stack(x: torch.Tensor, weights: Sequence[torch.Tensor], last_only: bool = False,
              masks: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
    return GRUStackFunction.apply(x, last_only, None if masks is None else tuple(masks), *weights)


def dropout_masks(B: int, T: int, H: int, n: int, p: float, device) -> List[torch.Tensor]:
    """n scaled Bernoulli masks (keep prob 1-p, scale 1/(1-p)) like nn.GRU's inter-layer dropout."""
    keep = 1.0 - p
    return [torch.bernoulli(torch.full((B, T, H), keep, dtype=torch.float32, device=device)) / keep for _ in range(n)]


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    return LinearFunction.apply(x, weight, bias)

"""BPTT kernel variants at the c2 / c1 layer shapes: two columns per thread (gru_bwd_pair_kernel) vs one."""
import sys, torch
sys.path.insert(0, '.')
import timegan_b200
from timegan_b200 import ops, _lib
from timegan_b200._lib import lib
dev = 'cuda'
T = 768
for H, B in [(64, 256), (64, 148), (24, 256), (56, 64)]:
    w = [torch.randn(3 * H, H, device=dev) / H ** 0.5, torch.randn(3 * H, H, device=dev) / H ** 0.5,
         torch.zeros(3 * H, device=dev), torch.zeros(3 * H, device=dev)]
    x = torch.rand(B, T, H, device=dev); dy = torch.randn(B, T, H, device=dev)
    for pair in (1, 0):
        lib.tg_set_option(b"bwd_pair", pair)
        def run():
            y, sv = ops.stack_forward(x, w, save=True)
            ops.stack_backward(dy, sv, w, need_dx=False, need_dw=False)
        for _ in range(2): run()
        _lib.prof_reset(); _lib.prof_enable(True)
        for _ in range(4): run()
        torch.cuda.synchronize(); _lib.prof_enable(False)
        p = _lib.prof_read()
        f, b = p['gru_fwd']['ms'] / p['gru_fwd']['calls'] * 1e3, p['gru_bwd']['ms'] / p['gru_bwd']['calls'] * 1e3
        print(f"H={H} B={B} pair={pair}: fwd {f:7.1f} us | bwd {b:7.1f} us  {b * 1965 / T:5.0f} clk/step", flush=True)
lib.tg_set_option(b"bwd_pair", 0)

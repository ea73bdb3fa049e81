"""Calibration for the wavefront / packing decisions: per-launch time of the recurrent kernels as a function of
(B, BT, T) -- one launch with B = 768, BT = 2 occupies the machine like three concurrent B = 256 passes."""
import sys, torch
sys.path.insert(0, '.')
import timegan_b200
from timegan_b200 import ops, _lib
dev = 'cuda'
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cases = [(256, 1, 768), (256, 2, 768), (512, 1, 768), (512, 2, 768), (768, 2, 768), (768, 4, 768), (1024, 2, 768),
         (1024, 4, 768), (256, 1, 128), (256, 2, 128), (768, 2, 128), (768, 2, 256), (296, 1, 768), (296, 2, 768)]
if H > 64:
    cases = [(256, 1, 768), (256, 2, 768), (256, 4, 768), (512, 2, 768), (512, 4, 768), (148, 1, 768), (296, 2, 768)]
for (B, bt, T) in cases:
    w = [torch.randn(3 * H, H, device=dev) / H ** 0.5, torch.randn(3 * H, H, device=dev) / H ** 0.5,
         torch.zeros(3 * H, device=dev), torch.zeros(3 * H, device=dev)]
    x = torch.rand(B, T, H, device=dev)
    dy = torch.randn(B, T, H, device=dev)
    ops.set_bt_override(bt)
    res = []
    for what in ['fwd_save', 'bwd']:
        def run():
            y, sv = ops.stack_forward(x, w, save=True)
            if what == 'bwd':
                ops.stack_backward(dy, sv, w, need_dx=False, need_dw=False)
        for _ in range(2):
            run()
        _lib.prof_reset(); _lib.prof_enable(True)
        for _ in range(4):
            run()
        torch.cuda.synchronize(); _lib.prof_enable(False)
        p = _lib.prof_read()
        k = 'gru_bwd' if what == 'bwd' else 'gru_fwd'
        us = p[k]['ms'] / p[k]['calls'] * 1e3
        res.append(f"{what} {us:7.1f} us  {us * 1965.0 / T:6.0f} clk/step  {us * 1965.0 / T / B * 148:6.1f} clk/seq-step/SM")
    print(f'H={H} B={B} BT={bt} T={T}: ' + ' | '.join(res), flush=True)
    del x, dy
ops.set_bt_override(0)

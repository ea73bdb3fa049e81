"""Reverse-over-tangent pass of one H = 256 layer at B = 256, T = 768: cluster kernel (W_hh^T in the registers of 8 SMs)
vs the L2-streaming kernel of gru_bigh.cu."""
import sys
import torch
sys.path.insert(0, '.')
import timegan_b200  # noqa
from timegan_b200 import ops
from timegan_b200._lib import lib
dev = 'cuda'
print("capacity jvp256:", lib.tg_cluster_capacity(256, 2, 1))
for B in (120, 256):
    T, H = 768, 256
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, T, H, generator=g).to(dev)
    v = torch.randn(B, T, H, generator=g).to(dev)
    w = [(torch.randn(3 * H, H, generator=g) / H ** 0.5).to(dev), (torch.randn(3 * H, H, generator=g) / H ** 0.5).to(dev),
         torch.zeros(3 * H, device=dev), torch.zeros(3 * H, device=dev)]
    _, sv = ops.stack_forward(x, w, save=True)
    _, ts = ops.stack_jvp_forward(v, sv, w)
    hb = torch.randn(B, H, generator=g).to(dev); hdb = torch.randn(B, H, generator=g).to(dev)
    res = {}
    for mode in (1, 0):
        lib.tg_set_option(b"cluster_jvp256", mode)
        grads = [torch.zeros_like(t) for t in w]
        for _ in range(2):
            ops.stack_jvp_backward(hb, hdb, sv, ts, w, grads, accumulate=False)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.stack_jvp_backward(hb, hdb, sv, ts, w, grads, accumulate=False)
        e1.record(); torch.cuda.synchronize()
        res[mode] = [g_.clone() for g_ in grads]
        print(f"B={B} cluster_jvp256={mode}: {e0.elapsed_time(e1) / 3:.2f} ms per reverse pass incl. weight gradients", flush=True)
    lib.tg_set_option(b"cluster_jvp256", 1)
    err = max(((a - b).norm() / b.norm()).item() for a, b in zip(res[1][:2], res[0][:2]))
    print(f"B={B}: cluster vs streaming weight gradients rel diff {err:.2e}")

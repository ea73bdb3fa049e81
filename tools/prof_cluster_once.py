import sys, torch
sys.path.insert(0, '.')
import timegan_b200
from timegan_b200 import ops
from timegan_b200._lib import lib
lib.tg_set_option(b"cluster", 2)
dev = 'cuda'; B, T, H = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 768, (int(sys.argv[2]) if len(sys.argv) > 2 else 128)
torch.manual_seed(0)
w = [torch.randn(3 * H, H, device=dev) / H ** 0.5, torch.randn(3 * H, H, device=dev) / H ** 0.5, torch.zeros(3 * H, device=dev), torch.zeros(3 * H, device=dev)]
x = torch.rand(B, T, H, device=dev); dy = torch.randn(B, T, H, device=dev)
for _ in range(2):
    y, sv = ops.stack_forward(x, w, save=True)
    ops.stack_backward(dy, sv, w, need_dx=False, need_dw=False)
torch.cuda.synchronize(); print("ok")

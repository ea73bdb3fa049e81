"""Stand-alone timing of the input projection in its three tensor-core forms at the c2 / c3 layer shapes:
3xTF32 and 1xTF32 (fp32 gi) vs bf16 operands + bf16 gi (tg_proj_bf16)."""
import sys
import torch
sys.path.insert(0, '.')
import timegan_b200  # noqa
from timegan_b200 import ops
from timegan_b200._lib import lib, check, ptr, stream_ptr
dev = 'cuda'


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (M, N, K) in [(196608, 192, 64), (196608, 384, 128), (393216, 384, 128), (196608, 192, 16), (196608, 384, 16)]:
    A = torch.rand(M, K, device=dev) * 2 - 0.7
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    for name in ['tf32x3', 'tf32']:
        mode = ops._MODES[name]
        ms = timeit(lambda: ops.proj(A, W, b, out2d=out, mode=mode))
        byts = 4 * (M * K + M * N)
        print(f'M={M} N={N} K={K} {name:7s} {ms*1e3:8.1f} us  {byts/ms/1e6:7.0f} GB/s', flush=True)
    W16 = W.to(torch.bfloat16)
    C16 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: check(lib.tg_proj_bf16(stream_ptr(), ptr(A), K, ptr(W16), K, ptr(b), ptr(C16), N, M, N, K), "p"))
    byts = 4 * M * K + 2 * M * N
    print(f'M={M} N={N} K={K} bf16    {ms*1e3:8.1f} us  {byts/ms/1e6:7.0f} GB/s', flush=True)

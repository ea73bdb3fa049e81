import torch, time, sys
sys.path.insert(0, '.')
import timegan_b200
from timegan_b200 import ops, _lib
dev='cuda'
def relerr(a,b): return ((a.double()-b.double()).norm()/b.double().norm()).item()
for (M,N,K) in [(196608,192,64),(196608,192,14),(196608,64,192),(196608,72,24),(393216,192,64)]:
    A=torch.rand(M,K,device=dev)*2-0.7; W=torch.randn(N,K,device=dev)/K**0.5; b=torch.randn(N,device=dev)
    ref=(A.double()@W.double().T+b.double())
    out=torch.empty(M,N,device=dev)
    for name in ['ffma','tf32','tf32x3']:
        mode=ops._MODES[name]
        for _ in range(3): ops.proj(A,W,b,out2d=out,mode=mode)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.proj(A,W,b,out2d=out,mode=mode)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        byts=4*(M*K+M*N+N*K)
        print(f'M={M} N={N} K={K} {name:7s} {ms*1e3:8.1f} us  {byts/ms/1e6:7.0f} GB/s  {2*M*N*K/ms/1e9:7.1f} TFLOP/s  err {relerr(out,ref):.2e}', flush=True)

print("---- wgrad ----")
for (M,N,K,T) in [(196608,192,64,0),(196608,128,64,768),(196608,64,64,768),(196608,72,24,768)]:
    dG=torch.randn(M,N,device=dev); A=torch.rand(M,K,device=dev)
    dW=torch.empty(N,K,device=dev); db=torch.empty(N,device=dev)
    for name in ['ffma','tf32','tf32x3']:
        mode=ops._MODES[name]
        for _ in range(3): ops.wgrad(dG,A,dW,db,N,shift_T=T,mode=mode)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.wgrad(dG,A,dW,db,N,shift_T=T,mode=mode)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        byts=4*(M*K+M*N)
        print(f'M={M} N={N} K={K} T={T} {name:7s} {ms*1e3:8.1f} us  {byts/ms/1e6:7.0f} GB/s  {2*M*N*K/ms/1e9:7.1f} TFLOP/s', flush=True)

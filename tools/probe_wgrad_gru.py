import sys, torch
sys.path.insert(0, '.')
import timegan_b200
from timegan_b200 import ops
dev='cuda'
B,T,I,H=256,768,64,64
dgi=torch.randn(B,T,3*H,device=dev); dq=torch.randn(B,T,H,device=dev); x=torch.rand(B*T,I,device=dev); y=torch.rand(B,T,H,device=dev)
gw=torch.empty(3*H,I,device=dev); gh=torch.empty(3*H,H,device=dev); bi=torch.empty(3*H,device=dev); bh=torch.empty(3*H,device=dev)
for name in ['tf32x3','tf32']:
    for _ in range(3): ops.wgrad_gru(dgi,dq,x,y,gw,gh,bi,bh,mode=ops._MODES[name])
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.wgrad_gru(dgi,dq,x,y,gw,gh,bi,bh,mode=ops._MODES[name])
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print(f'wgrad_gru {name}: {ms*1e3:.1f} us  {4*B*T*(4*H+I+H)/ms/1e6:.0f} GB/s', flush=True)
# accuracy of the fp32-parity mode against fp64 (dW_ih = dGI^T x, dW_hh[:2H] = dGI[:, :2H]^T h_prev)
ops.wgrad_gru(dgi,dq,x,y,gw,gh,bi,bh,mode=ops._MODES['tf32x3'])
ref_ih=(dgi.view(B*T,3*H).double().T@x.double())
yp=torch.zeros_like(y); yp[:,1:]=y[:,:-1]
ref_hh=(dgi.view(B*T,3*H)[:,:2*H].double().T@yp.view(B*T,H).double())
rel=lambda a,b: ((a.double()-b).norm()/b.norm()).item()
print(f'tf32x3 relerr dW_ih {rel(gw,ref_ih):.2e}  dW_hh[:2H] {rel(gh[:2*H],ref_hh):.2e}', flush=True)

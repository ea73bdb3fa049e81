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

import sys, torch
sys.path.insert(0, '.')
import timegan_b200
from timegan_b200 import ops
dev='cuda'
M,N,K,T=196608,192,64,768
dG=torch.randn(M,N,device=dev); A=torch.rand(M,K,device=dev); W=torch.randn(N,K,device=dev)
dW=torch.empty(N,K,device=dev); db=torch.empty(N,device=dev); out=torch.empty(M,N,device=dev)
for name in ['tf32','tf32x3']:
    for _ in range(3):
        ops.wgrad(dG,A,dW,db,N,shift_T=T,mode=ops._MODES[name])
        ops.proj(A,W,db,out2d=out,mode=ops._MODES[name])
torch.cuda.synchronize()
print('done')

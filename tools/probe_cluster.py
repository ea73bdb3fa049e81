import sys, torch
sys.path.insert(0, '.')
import timegan_b200
from timegan_b200 import ops, _lib
from timegan_b200._lib import lib
dev='cuda'
torch.zeros(1, device=dev)
print('resident-cluster capacity: H=128 fwd', [lib.tg_cluster_capacity(128,0,g) for g in (1,2)], 'bwd', [lib.tg_cluster_capacity(128,1,g) for g in (1,2)], 'jvp', [lib.tg_cluster_capacity(128,2,g) for g in (1,2)], '| H=256 fwd', [lib.tg_cluster_capacity(256,0,g) for g in (1,2,3,4)], 'bwd', [lib.tg_cluster_capacity(256,1,g) for g in (1,2,3)], flush=True)
for H,B in [(128,256),(128,296),(128,512),(256,256),(256,128)]:
    T=768
    w=[torch.randn(3*H,H,device=dev)/H**0.5, torch.randn(3*H,H,device=dev)/H**0.5, torch.zeros(3*H,device=dev), torch.zeros(3*H,device=dev)]
    x=torch.rand(B,T,H,device=dev); dy=torch.randn(B,T,H,device=dev)
    for cl in (2,0):
        lib.tg_set_option(b"cluster", cl)
        res=[]
        for what in ['fwd_save','bwd']:
            def run():
                y,sv=ops.stack_forward(x,w,save=True)
                if what=='bwd': ops.stack_backward(dy,sv,w,need_dx=False,need_dw=False)
            for _ in range(2): run()
            _lib.prof_reset(); _lib.prof_enable(True)
            for _ in range(3): run()
            torch.cuda.synchronize(); _lib.prof_enable(False)
            p=_lib.prof_read(); k='gru_bwd' if what=='bwd' else 'gru_fwd'
            us=p[k]['ms']/p[k]['calls']*1e3
            res.append(f"{what} {us:8.1f} us {us*1965/T:6.0f} clk/step")
        print(f"H={H} B={B} cluster={cl}: "+' | '.join(res), flush=True)
lib.tg_set_option(b"cluster", 1)

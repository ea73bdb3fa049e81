import sys, torch
sys.path.insert(0, '.')
import timegan_b200
from timegan_b200 import ops, _lib
dev='cuda'
T=768
for (B,H) in [(256,64),(512,64),(256,24),(32,24),(256,128)]:
    I=H
    w=[torch.randn(3*H,I,device=dev)/H**0.5, torch.randn(3*H,H,device=dev)/H**0.5, torch.zeros(3*H,device=dev), torch.zeros(3*H,device=dev)]
    x=torch.rand(B,T,I,device=dev); dy=torch.randn(B,T,H,device=dev)
    for bt in [1,2,4]:
        ops.set_bt_override(bt)
        res=[]
        for what in ['fwd','fwd_save','bwd']:
            def run():
                if what=='fwd': return ops.stack_forward(x,w,save=False)
                y,sv=ops.stack_forward(x,w,save=True)
                if what=='bwd': ops.stack_backward(dy,sv,w,need_dx=False,need_dw=False)
            for _ in range(2): run()
            _lib.prof_reset(); _lib.prof_enable(True)
            for _ in range(5): run()
            torch.cuda.synchronize(); _lib.prof_enable(False)
            p=_lib.prof_read()
            k='gru_bwd' if what=='bwd' else 'gru_fwd'
            res.append(f"{what} {p[k]['ms']/p[k]['calls']*1e3:7.1f} us")
        print(f'B={B} H={H} BT={bt}: '+'  '.join(res), flush=True)
    ops.set_bt_override(0)

import sys, torch
sys.path.insert(0, '.')
import timegan_b200
from timegan_b200 import ops
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
dev = 'cuda'
M, N, K = 512, 64, 32
for name in ['tf32', 'tf32x3']:
    mode = ops._MODES[name]
    for case in range(3):
        if case == 0:
            dG = torch.ones(M, N, device=dev); A = torch.ones(M, K, device=dev)
        elif case == 1:
            dG = (torch.arange(N, device=dev).float() + 1).repeat(M, 1); A = torch.ones(M, K, device=dev)
        else:
            dG = torch.ones(M, N, device=dev); A = (torch.arange(K, device=dev).float() + 1).repeat(M, 1)
        dW = torch.full((N, K), -1.0, device=dev); db = torch.full((N,), -1.0, device=dev)
        ops.wgrad(dG, A, dW, db, N, mode=mode)
        torch.cuda.synchronize()
        ref = dG.double().T @ A.double()
        print(name, 'case', case, 'max err', (dW.double() - ref).abs().max().item(), 'ref max', ref.abs().max().item())
        print(' dW[:4,:8]=', dW[:4, :8].tolist())
        print(' dW[32:34,:8]=', dW[32:34, :8].tolist())
        print(' db[:8]=', db[:8].tolist(), 'expect', dG.sum(0)[:8].tolist())

"""Shared helpers of the parity tests: normwise error, torch.nn.GRU reference stacks (CPU fp32 = the
reference's own arithmetic path, timegan_model.py:24-34)."""
import torch


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| (normwise; SURVEY.md 7.2 item 4: elementwise relative error is meaningless for
    gradients that vanish over 768 steps)."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    den = b.norm().item()
    num = (a - b).norm().item()
    return num / den if den > 0 else num


def make_gru(I, H, L, seed=0, scale=None):
    """CPU torch.nn.GRU with reproducible weights."""
    g = torch.Generator().manual_seed(seed)
    m = torch.nn.GRU(I, H, num_layers=L, batch_first=True)
    s = scale if scale is not None else 1.0 / (H ** 0.5)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * s)
    return m


def flat_weights(m, device):
    out = []
    for l in range(m.num_layers):
        for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            out.append(getattr(m, f"{n}_l{l}").detach().to(device).contiguous())
    return out

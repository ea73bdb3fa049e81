"""B200-native drop-in for timeGAN/timegan_model.py (reference file:line cited per class).

Same class names, constructor signatures, attribute paths (`model.embedder.rnn.rnn.hidden_size` is read at
train_timegan.py:179,235), helper passes and `state_dict` keys
(`embedder.rnn.rnn.weight_ih_l0` ... `discriminator.fc.{bias,weight_orig,weight_u,weight_v}`), so reference
checkpoints load unchanged -- but every GRU stack runs through the hand-written sm_100a kernels of
csrc/ (projection GEMM + persistent recurrent kernel per layer) instead of torch.nn.GRU.
There is no CPU path: calling a module on a CPU tensor raises.
"""
import math

import torch
import torch.nn as nn
import torch.nn.utils as U

from . import ops


class FusedGRU(nn.Module):
    """Parameter container + forward for an L-layer batch_first GRU with h0 = 0.

    Exposes the nn.GRU attributes the reference touches: weight_ih_l{k}, weight_hh_l{k}, bias_ih_l{k},
    bias_hh_l{k}, hidden_size, input_size, num_layers, dropout, batch_first.  Default initialisation draws
    from the global generator exactly like nn.RNNBase.reset_parameters (same order, same calls), so seeding
    and then building TimeGAN reproduces the reference's initial weights.
    """

    def __init__(self, input_size: int, hidden_size: int, num_layers: int = 1, dropout: float = 0.0,
                 batch_first: bool = True):
        super().__init__()
        if not batch_first:
            raise ValueError("FusedGRU only implements batch_first=True (the layout the reference uses)")
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.dropout, self.batch_first = float(dropout), True
        self.bias, self.bidirectional = True, False
        self._flat_names = []
        for l in range(num_layers):
            i = input_size if l == 0 else hidden_size
            shapes = [(3 * hidden_size, i), (3 * hidden_size, hidden_size), (3 * hidden_size,), (3 * hidden_size,)]
            for nm, shp in zip(("weight_ih", "weight_hh", "bias_ih", "bias_hh"), shapes):
                name = f"{nm}_l{l}"
                self.register_parameter(name, nn.Parameter(torch.empty(*shp)))
                self._flat_names.append(name)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.hidden_size) if self.hidden_size > 0 else 0
        for w in self.parameters():
            nn.init.uniform_(w, -stdv, stdv)

    def layer_weights(self, lo: int = 0, hi: int = None):
        hi = self.num_layers if hi is None else hi
        return [getattr(self, n) for n in self._flat_names[4 * lo:4 * hi]]

    def dropout_masks(self, x):
        """Inter-layer dropout masks for this call (None when inactive: eval mode, p = 0 or one layer)."""
        p = self.dropout if (self.training and self.num_layers > 1) else 0.0
        if p <= 0.0:
            return None
        return ops.dropout_masks(x.shape[0], x.shape[1], self.hidden_size, self.num_layers - 1, p, x.device)

    def forward(self, x, hx=None, last_only: bool = False, frozen: bool = False):
        """Returns (y, None) like nn.GRU (h_n is not produced: the reference discards it, timegan_model.py:33).

        last_only: y is (B,H) = the last timestep only.  frozen: treat the weights as constants (no weight
        gradients; the input gradient still flows) -- the discriminator inside gen_step."""
        if hx is not None:
            raise NotImplementedError("FusedGRU: the reference never passes an initial state (h0 = 0)")
        w = self.layer_weights()
        if frozen:
            w = [t.detach() for t in w]
        y = ops.gru_stack(x, w, last_only=last_only, masks=self.dropout_masks(x))
        return y, None

    def extra_repr(self):
        return f"{self.input_size}, {self.hidden_size}, num_layers={self.num_layers}, dropout={self.dropout}"


def init_weights_(m):
    """timegan_model.py:11-21: xavier_uniform on Linear weights and on the stacked GRU matrices, zero biases."""
    if isinstance(m, (nn.Linear,)):
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    if isinstance(m, (FusedGRU, nn.GRU, nn.LSTM)):
        for name, param in m.named_parameters():
            if "weight" in name:
                nn.init.xavier_uniform_(param)
            elif "bias" in name:
                nn.init.zeros_(param)


class GRUStack(nn.Module):
    """timegan_model.py:24-34."""

    def __init__(self, input_dim: int, hidden_dim: int, num_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        self.rnn = FusedGRU(input_dim, hidden_dim, num_layers=num_layers,
                            dropout=dropout if num_layers > 1 else 0.0, batch_first=True)

    def forward(self, x):
        y, _ = self.rnn(x)  # (B, T, H)
        return y


class Embedder(nn.Module):
    """X -> H (timegan_model.py:37-44)."""

    def __init__(self, x_dim: int, z_dim: int, num_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        self.rnn = GRUStack(x_dim, z_dim, num_layers, dropout)

    def forward(self, x):
        return self.rnn(x)


class Recovery(nn.Module):
    """H -> X~ (timegan_model.py:47-57)."""

    def __init__(self, z_dim: int, x_dim: int, hidden_dim: int = None, num_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        h = hidden_dim or z_dim
        self.rnn = GRUStack(z_dim, h, num_layers, dropout)
        self.out = nn.Linear(h, x_dim)

    def forward(self, h):
        return ops.linear(self.rnn(h), self.out.weight, self.out.bias)


class _LatentStack(nn.Module):
    def __init__(self, z_dim: int, hidden_dim: int = None, num_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        h = hidden_dim or z_dim
        self.rnn = GRUStack(z_dim, h, num_layers, dropout)
        self.proj = nn.Linear(h, z_dim) if h != z_dim else nn.Identity()

    def forward(self, z):
        y = self.rnn(z)
        if isinstance(self.proj, nn.Identity):
            return y
        return ops.linear(y, self.proj.weight, self.proj.bias)


class Generator(_LatentStack):
    """Z -> E_hat (timegan_model.py:60-70)."""


class Supervisor(_LatentStack):
    """E_hat -> H_hat (timegan_model.py:73-83)."""


class Discriminator(nn.Module):
    """H or H_hat -> prob(real) (timegan_model.py:86-98): last step -> spectral-norm Linear -> sigmoid.

    `fc` is registered through torch.nn.utils.spectral_norm exactly like the reference, so the parameters and
    buffers (`fc.bias`, `fc.weight_orig`, `fc.weight_u`, `fc.weight_v`) and their state_dict keys are the
    reference's.  The (B,H)x(H,1) head itself is evaluated by `head()` below, which restates the legacy
    spectral-norm hook (one power iteration per call in train mode, eps 1e-12) so the training steps can
    differentiate it separately from the GRU stack (R1)."""

    def __init__(self, z_dim: int, hidden_dim: int = 32, num_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        self.rnn = GRUStack(z_dim, hidden_dim, num_layers, dropout)
        self.fc = U.spectral_norm(nn.Linear(hidden_dim, 1))
        self.sigmoid = nn.Sigmoid()

    def sn_weight(self, frozen: bool = False):
        """w / sigma with sigma = u^T W v after one power iteration when training (legacy spectral_norm)."""
        w = self.fc.weight_orig.detach() if frozen else self.fc.weight_orig
        u, v = self.fc.weight_u, self.fc.weight_v
        if self.training:
            with torch.no_grad():
                wm = self.fc.weight_orig.detach()
                v.copy_(torch.nn.functional.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12))
                u.copy_(torch.nn.functional.normalize(torch.mv(wm, v), dim=0, eps=1e-12))
            u, v = u.clone(), v.clone()
        sigma = torch.dot(u, torch.mv(w, v))
        return w / sigma

    def head(self, last, frozen: bool = False):
        """(B,H) last hidden state -> (B,1) probability."""
        b = self.fc.bias.detach() if frozen else self.fc.bias
        return self.sigmoid(torch.nn.functional.linear(last, self.sn_weight(frozen), b))

    def forward(self, h, frozen: bool = False):
        last, _ = self.rnn.rnn(h, last_only=True, frozen=frozen)
        return self.head(last, frozen)

    def adv_loss(self, h):
        """bce(self(h), ones) with this network's weights as constants (gen_step, train_timegan.py:240-241): the GRU
        stack keeps only its last step, and head + sigmoid + BCE (+ their backward) are one fused kernel each."""
        from . import head as _head
        last, _ = self.rnn.rnn(h, last_only=True, frozen=True)
        return _head.adv_loss(self, last)


class TimeGAN(nn.Module):
    """Bundle of submodules + convenience calls (timegan_model.py:101-118)."""

    def __init__(self, x_dim: int, z_dim: int, hidden_dim: int, num_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        self.embedder = Embedder(x_dim, z_dim, num_layers, dropout)
        self.recovery = Recovery(z_dim, x_dim, hidden_dim, num_layers, dropout)
        self.generator = Generator(z_dim, hidden_dim, num_layers, dropout)
        self.supervisor = Supervisor(z_dim, hidden_dim, num_layers, dropout)
        self.discriminator = Discriminator(z_dim, hidden_dim, num_layers, dropout)
        self.apply(init_weights_)

    def encode(self, x):        return self.embedder(x)
    def reconstruct(self, x):   return self.recovery(self.embedder(x))
    def gen_latent(self, z):    return self.generator(z)
    def refine_latent(self, e): return self.supervisor(e)
    def decode(self, h):        return self.recovery(h)
    def disc(self, h):          return self.discriminator(h)

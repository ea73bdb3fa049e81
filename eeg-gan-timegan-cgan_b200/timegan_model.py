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

    def forward(self, x, hx=None):
        if hx is not None:
            raise NotImplementedError("FusedGRU: the reference never passes an initial state (h0 = 0)")
        p = self.dropout if (self.training and self.num_layers > 1) else 0.0
        if p == 0.0:
            y = ops.gru_stack(x, self.layer_weights())
        else:
            # inter-layer dropout (timegan_model.py:29): layer-by-layer with a mask in between
            y = x
            for l in range(self.num_layers):
                y = ops.gru_stack(y, self.layer_weights(l, l + 1))
                if l < self.num_layers - 1:
                    y = torch.nn.functional.dropout(y, p, True)
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

    The (B,H)x(H,1) head is host-side torch (legacy spectral_norm hook, same buffers/keys as the reference).
    """

    def __init__(self, z_dim: int, hidden_dim: int = 32, num_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        self.rnn = GRUStack(z_dim, hidden_dim, num_layers, dropout)
        self.fc = U.spectral_norm(nn.Linear(hidden_dim, 1))
        self.sigmoid = nn.Sigmoid()

    def head(self, last):
        return self.sigmoid(self.fc(last))

    def forward(self, h):
        y = self.rnn(h)
        return self.head(y[:, -1, :])


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

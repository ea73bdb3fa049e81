"""On-device noise for the training step (host side of csrc/rng.cu, Philox4x32-10).

Replaces sample_noise (train_timegan.py:64-65), add_instance_noise (tt:46-47) and smooth_labels (tt:40-43)
in production mode.  Parity runs inject the reference's own CPU draws instead (`HostReplayNoise` in
train_timegan.py of this package).
"""
import torch

from ._lib import lib, check, ptr, stream_ptr, require_cuda


def uniform(shape, device, seed: int, offset: int, lo: float = 0.0, hi: float = 1.0) -> torch.Tensor:
    out = torch.empty(shape, dtype=torch.float32, device=device)
    require_cuda(out, "noise output")
    check(lib.tg_rng_uniform(stream_ptr(), ptr(out), out.numel(), seed, offset, lo, hi), "tg_rng_uniform")
    return out


def add_normal(h: torch.Tensor, std: float, seed: int, offset: int) -> torch.Tensor:
    """h + std*N(0,1); returns h itself when std <= 0 (tt:46-47)."""
    if std <= 0:
        return h
    require_cuda(h, "instance-noise input")
    h = h.contiguous()
    out = torch.empty_like(h)
    check(lib.tg_rng_add_normal(stream_ptr(), ptr(h), ptr(out), h.numel(), float(std), seed, offset),
          "tg_rng_add_normal")
    return out


def draws_for(n: int) -> int:
    """Philox counter values one call of n elements consumes (offset bookkeeping)."""
    return (n + 3) // 4


class DeviceNoise:
    """Production noise source: Philox stream keyed by (seed, running offset); draw order of SURVEY.md App. B."""

    def __init__(self, seed: int, device):
        self.seed, self.device, self.offset = int(seed) & 0xFFFFFFFFFFFFFFFF, device, 0

    def _adv(self, n):
        o = self.offset
        self.offset += draws_for(n)
        return o

    def rand(self, *shape):
        n = 1
        for s in shape:
            n *= s
        return uniform(shape, self.device, self.seed, self._adv(n))

    def add_randn(self, h, std):
        if std <= 0:
            return h
        return add_normal(h, std, self.seed, self._adv(h.numel()))

    def state(self):
        return {"seed": self.seed, "offset": self.offset}

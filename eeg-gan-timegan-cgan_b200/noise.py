"""On-device noise for the training step (host side of csrc/rng.cu, Philox4x32-10).

Replaces sample_noise (train_timegan.py:64-65), add_instance_noise (tt:46-47) and smooth_labels (tt:40-43)
in production mode.  Parity runs inject the reference's own CPU draws instead (`HostReplayNoise` in
train_timegan.py of this package).
"""
import torch

from ._lib import lib, check, ptr, stream_ptr, require_cuda


def uniform(shape, device, seed: int, offset: int, lo: float = 0.0, hi: float = 1.0, ctr=None) -> torch.Tensor:
    """U[lo,hi).  `ctr`: optional int64 device scalar added to `offset` on the device (CUDA-graph replay)."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    require_cuda(out, "noise output")
    check(lib.tg_rng_uniform(stream_ptr(), ptr(out), out.numel(), seed, offset, lo, hi, ptr(ctr)), "tg_rng_uniform")
    return out


def add_normal(h: torch.Tensor, std, seed: int, offset: int, ctr=None, out=None) -> torch.Tensor:
    """h + std*N(0,1) in ONE kernel (tt:46-47).  `std` is a float (returns h itself when <= 0) or a 0-dim device
    tensor (CUDA-graph replay); `out` may be a contiguous view to write into (e.g. one half of a [real;fake] batch)."""
    dev_std = torch.is_tensor(std)
    if not dev_std and std <= 0:
        if out is None:
            return h
        out.copy_(h)
        return out
    require_cuda(h, "instance-noise input")
    h = h.contiguous()
    if out is None:
        out = torch.empty_like(h)
    elif not (out.is_contiguous() and out.shape == h.shape):
        raise ValueError("add_normal: `out` must be a contiguous tensor of the input's shape")
    if dev_std:
        check(lib.tg_rng_add_normal_dev(stream_ptr(), ptr(h), ptr(out), h.numel(), ptr(std), seed, offset, ptr(ctr)),
              "tg_rng_add_normal_dev")
    else:
        check(lib.tg_rng_add_normal(stream_ptr(), ptr(h), ptr(out), h.numel(), float(std), seed, offset, ptr(ctr)),
              "tg_rng_add_normal")
    return out


def normal(shape, device, seed: int, offset: int, ctr=None) -> torch.Tensor:
    """N(0,1) of the given shape."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    require_cuda(out, "noise output")
    check(lib.tg_rng_add_normal(stream_ptr(), None, ptr(out), out.numel(), 1.0, seed, offset, ptr(ctr)),
          "tg_rng_add_normal")
    return out


def draws_for(n: int) -> int:
    """Philox counter values one call of n elements consumes (offset bookkeeping)."""
    return (n + 3) // 4


class DeviceNoise:
    """Production noise source: one Philox stream keyed by `seed`; position = device counter + offset within the
    step.  The counter lives in device memory and is advanced by `end()` with a device-side add, so a captured
    CUDA graph (which bakes the per-call offsets in) draws fresh numbers at every replay, and eager and graphed
    runs consume the identical stream.  Draw order inside a step follows SURVEY.md Appendix B."""

    def __init__(self, seed: int, device):
        self.seed, self.device = int(seed) & 0x7FFFFFFFFFFFFFFF, device
        self.ctr = torch.zeros((), dtype=torch.int64, device=device)
        self.local = 0
        self.held = False

    def _adv(self, n):
        o = self.local
        self.local += draws_for(n)
        return o

    def begin(self):
        if not self.held:
            self.local = 0

    def end(self):
        if self.held:
            return
        if self.local:
            self.ctr.add_(self.local)
        self.local = 0

    def hold(self):
        """Treat the step functions that follow as ONE transaction: their begin()/end() neither reset the offset nor
        touch the device counter, so step functions running on different streams never race on it; release() advances
        the counter once.  Every draw lands on the Philox position it has when the functions run one after another."""
        self.local = 0
        self.held = True

    def release(self):
        self.held = False
        self.end()

    def rand(self, *shape):
        n = 1
        for s in shape:
            n *= s
        return uniform(shape, self.device, self.seed, self._adv(n), ctr=self.ctr)

    def randn(self, shape):
        n = 1
        for s in shape:
            n *= s
        return normal(tuple(shape), self.device, self.seed, self._adv(n), ctr=self.ctr)

    def add_randn(self, h, std, out=None):
        if not torch.is_tensor(std) and std <= 0:
            return add_normal(h, std, self.seed, 0, out=out)
        return add_normal(h, std, self.seed, self._adv(h.numel()), ctr=self.ctr, out=out)

    def state(self):
        return {"seed": self.seed, "ctr": int(self.ctr.item())}

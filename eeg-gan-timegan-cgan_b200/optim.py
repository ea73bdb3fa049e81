"""FusedAdam: clip_grad_norm_ + torch.optim.Adam.step in two multi-tensor kernel launches (csrc/optim.cu).

Replaces the pairs `nn.utils.clip_grad_norm_(params, clip); opt.step()` at
timeGAN/train_timegan.py:141-142, 160-161, 220-221, 268-272.  Constructor signature and `state_dict()` layout
follow torch.optim.Adam (state: step / exp_avg / exp_avg_sq per parameter; same param_group keys), so the
checkpoints written by `save_ckpt` (tt:58-61) keep the reference's schema and load into torch.optim.Adam.
"""
import ctypes as C
from typing import Iterable, List, Optional

import torch

from ._lib import lib, check, ptr, stream_ptr


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False,
                 capturable=False):
        """capturable=True keeps the step counter, the learning rate and the bias corrections in device memory
        (one float[4] per param group), so a captured CUDA graph replays a correct Adam step; `push_lr()` copies
        the param groups' current lr (e.g. after a scheduler step) to the device outside the graph."""
        if weight_decay != 0 or amsgrad:
            raise ValueError("FusedAdam implements the reference's configuration only (no weight decay, no amsgrad)")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                        foreach=None, capturable=bool(capturable), differentiable=False, fused=None,
                        decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.capturable = bool(capturable)
        self._dev_state = {}     # group index -> device float[4] {lr, step, step_size, bc2_sqrt}
        self._pushed_lr = {}

    def _group_state(self, gi, device):
        t = self._dev_state.get(gi)
        if t is None:
            steps = [float(self.state[p]["step"]) for p in self.param_groups[gi]["params"] if "step" in self.state[p]]
            t = torch.tensor([float(self.param_groups[gi]["lr"]), max(steps) if steps else 0.0, 0.0, 1.0],
                             dtype=torch.float32, device=device)
            self._dev_state[gi] = t
            self._pushed_lr[gi] = float(self.param_groups[gi]["lr"])
        return t

    def push_lr(self):
        """Copy each group's lr to its device state if it changed (call outside graph capture/replay)."""
        for gi, t in self._dev_state.items():
            lr = float(self.param_groups[gi]["lr"])
            if self._pushed_lr.get(gi) != lr:
                t[0:1].fill_(lr)
                self._pushed_lr[gi] = lr

    def state_dict(self):
        if self.capturable:   # materialise the device step counters in torch.optim.Adam's per-parameter layout
            for gi, t in self._dev_state.items():
                step = float(t[1].item())
                for p in self.param_groups[gi]["params"]:
                    if p in self.state and "step" in self.state[p]:
                        self.state[p]["step"] = torch.tensor(step, dtype=torch.float32)
        return super().state_dict()

    # -- helpers ---------------------------------------------------------------------------------
    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @staticmethod
    def _ptr_array(tensors: List[torch.Tensor]):
        arr = (C.c_void_p * len(tensors))()
        for i, t in enumerate(tensors):
            arr[i] = t.data_ptr()
        return arr

    def grad_sumsq(self, params: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """Device scalar holding sum g^2 over all parameters that have gradients."""
        if params is None:
            params = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in params]
        dev = grads[0].device
        sizes = (C.c_longlong * len(grads))(*[g.numel() for g in grads])
        out = torch.empty(1, dtype=torch.float32, device=dev)
        nbytes = lib.tg_sumsq_workspace_bytes(len(grads), sizes)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        check(lib.tg_sumsq(stream_ptr(), len(grads), self._ptr_array(grads), sizes, ptr(out), ptr(ws), nbytes),
              "tg_sumsq")
        return out

    @torch.no_grad()
    def clip_and_step(self, max_norm: float = 0.0, grad_scale: float = 1.0) -> Optional[torch.Tensor]:
        """g <- g*grad_scale; clip to global L2 norm `max_norm` (<=0: no clip); one Adam step.

        Returns the device scalar sum g^2 (before scaling) when clipping, else None.
        """
        self._opt_called = True  # lr_scheduler's "scheduler before optimizer" check looks at this flag
        all_params = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not all_params:
            return None
        if not all_params[0].is_cuda:
            raise RuntimeError("FusedAdam: parameters are on the CPU; the fused optimiser only exists as CUDA kernels")
        sumsq = self.grad_sumsq(all_params) if max_norm and max_norm > 0 else None
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            grads, ms, vs = [], [], []
            step = None
            dev = None
            if self.capturable:
                for p in params:
                    self._init_state(p)
                dev = self._group_state(gi, all_params[0].device)
                if not torch.cuda.is_current_stream_capturing():
                    self.push_lr()
                step = 0
            for p in params:
                st = self._init_state(p)
                if not self.capturable:
                    st["step"] += 1
                    step = int(st["step"].item()) if step is None else step
                grads.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            b1, b2 = group["betas"]
            sizes = (C.c_longlong * len(params))(*[p.numel() for p in params])
            check(lib.tg_adam(stream_ptr(), len(params), self._ptr_array(params), self._ptr_array(grads),
                              self._ptr_array(ms), self._ptr_array(vs), sizes, ptr(sumsq), float(max_norm or 0.0),
                              float(group["lr"]), float(b1), float(b2), float(group["eps"]), step,
                              float(grad_scale), ptr(dev)), "tg_adam")
        return sumsq

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.clip_and_step(0.0)
        return loss


def clip_and_step(opt, params: Iterable[torch.Tensor], max_norm: float):
    """clip_grad_norm_(params, max_norm) + opt.step(): fused when `opt` is a FusedAdam, else the torch calls."""
    if isinstance(opt, FusedAdam):
        opt.clip_and_step(max_norm)
    else:
        torch.nn.utils.clip_grad_norm_(list(params), max_norm)
        opt.step()

"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in
the CPU tests).  Net-new relative to the reference, which is single-process (SURVEY.md sections 2.3, 8e).

Convention used everywhere in this package under DP:
  * every rank holds a contiguous shard of the GLOBAL batch;
  * every loss VALUE is the global-batch loss (identical on all ranks): batch statistics that are non-linear
    in the batch (recon's sqrt of the global MSE, the covariance / ACF moments, the throttle accuracy) are
    made exact by all-reducing their small sufficient statistics (`allreduce_stats`, ~10-25 KB per step);
  * each rank's backward therefore yields the contribution of ITS samples to the global gradient, so the
    parameter gradients are combined with a SUM all-reduce (`allreduce_grads`), bucketed per module and
    issued on a side stream as soon as a stack's weight gradients are complete, overlapping the BPTT of
    the remaining stacks;
  * clip_grad_norm_ + Adam then run on the reduced gradients, identically on every rank.
"""
import os
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as td

_GROUP = None          # process group in use (None = single process)
_ENABLED = False
_COMM_STREAM = None    # side stream for gradient buckets


def init(backend: Optional[str] = None, device: Optional[torch.device] = None):
    """Initialise from torchrun's environment (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*). Returns (rank, world)."""
    global _GROUP, _ENABLED
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        _ENABLED = False
        return 0, 1
    if not td.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        td.init_process_group(backend=backend, **kw)
    _GROUP = td.group.WORLD
    _ENABLED = True
    return td.get_rank(), td.get_world_size()


def enable(group=None):
    """Use an already-initialised process group (tests)."""
    global _GROUP, _ENABLED
    _GROUP = group if group is not None else td.group.WORLD
    _ENABLED = td.get_world_size(_GROUP) > 1


def disable():
    global _GROUP, _ENABLED, _COMM_STREAM
    _GROUP, _ENABLED, _COMM_STREAM = None, False, None


def is_enabled() -> bool:
    return _ENABLED


def world_size() -> int:
    return td.get_world_size(_GROUP) if _ENABLED else 1


def rank() -> int:
    return td.get_rank(_GROUP) if _ENABLED else 0


def allreduce_stats(t: torch.Tensor, count) -> Tuple[torch.Tensor, float]:
    """SUM-all-reduce a small statistics tensor together with its sample count.

    Returns (reduced tensor, global count).  Single process: identity.  Shards are equal-sized by
    construction (`shard_batch` refuses ragged splits), so the global count is world * count and no host
    synchronisation is needed to learn it.
    """
    if not _ENABLED:
        return t, float(count)
    buf = t.detach().clone().contiguous()
    td.all_reduce(buf, op=td.ReduceOp.SUM, group=_GROUP)
    return buf, float(count) * world_size()


class _GlobalMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, local_sum, local_count):
        s, n = allreduce_stats(local_sum.reshape(1), local_count)
        ctx.n = n
        return (s / n).reshape(())

    @staticmethod
    def backward(ctx, g):
        return g / ctx.n, None


def global_mean(local_sum: torch.Tensor, local_count) -> torch.Tensor:
    """mean over the global batch of a quantity whose local sum is `local_sum` (differentiable)."""
    if not _ENABLED:
        return local_sum / float(local_count)
    return _GlobalMean.apply(local_sum, local_count)


def comm_stream() -> "torch.cuda.Stream":
    global _COMM_STREAM
    if _COMM_STREAM is None:
        _COMM_STREAM = torch.cuda.Stream()
    return _COMM_STREAM


class GradBuckets:
    """Bucketed SUM all-reduce of parameter gradients on a side stream.

    launch(params) may be called as soon as those parameters' gradients are final (e.g. right after a stack's
    weight-gradient GEMMs); wait() joins the side stream before clip+Adam.
    """

    def __init__(self):
        self._pending: List[Tuple[torch.Tensor, List[torch.Tensor]]] = []
        self._work = []

    def launch(self, params: Iterable[torch.Tensor]):
        if not _ENABLED:
            return
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        if grads[0].is_cuda:
            cs = comm_stream()
            cs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cs):
                flat = torch.cat([g.reshape(-1) for g in grads])
                td.all_reduce(flat, op=td.ReduceOp.SUM, group=_GROUP)
                for g in grads:
                    g.record_stream(cs)
            self._pending.append((flat, grads))
        else:
            flat = torch.cat([g.reshape(-1) for g in grads])
            td.all_reduce(flat, op=td.ReduceOp.SUM, group=_GROUP)
            self._pending.append((flat, grads))

    @staticmethod
    def _scatter_back(flat, grads):
        views, o = [], 0
        for g in grads:
            n = g.numel()
            views.append(flat[o:o + n].view_as(g))
            o += n
        torch._foreach_copy_(grads, views)     # one multi-tensor kernel instead of one copy per parameter

    def wait(self):
        if not self._pending:
            return
        cuda = self._pending[0][0].is_cuda
        if cuda:
            cs = comm_stream()
            with torch.cuda.stream(cs):
                for flat, grads in self._pending:
                    self._scatter_back(flat, grads)
            torch.cuda.current_stream().wait_stream(cs)
        else:
            for flat, grads in self._pending:
                self._scatter_back(flat, grads)
        self._pending.clear()


class GradReducer:
    """Overlaps the gradient all-reduce with the rest of the backward pass.

    `buckets` is a list of parameter lists (one per network).  A post-accumulate-grad hook on every parameter
    counts down its bucket; when the last gradient of a bucket has been written, the bucket's SUM all-reduce is
    launched on the communication stream (GradBuckets.launch) while autograd keeps running BPTT for the
    remaining networks.  arm() before backward, finish() after it (launches whatever did not fire, then joins).
    Single process: every call is a no-op."""

    def __init__(self, buckets):
        self.buckets = [list(b) for b in buckets]
        self.where = {}
        self.armed = False
        self.left = []
        self.launched = []
        self.gb = GradBuckets()
        self.handles = []
        for bi, b in enumerate(self.buckets):
            for p in b:
                self.where[id(p)] = bi
                self.handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def arm(self):
        self.armed = _ENABLED
        self.left = [len(b) for b in self.buckets]
        self.launched = [False] * len(self.buckets)

    def _hook(self, p):
        if not self.armed:
            return
        bi = self.where[id(p)]
        self.left[bi] -= 1
        if self.left[bi] == 0 and not self.launched[bi]:
            self.launched[bi] = True
            self.gb.launch(self.buckets[bi])

    def finish(self):
        if not self.armed:
            return
        for bi, b in enumerate(self.buckets):
            if not self.launched[bi]:
                self.launched[bi] = True
                self.gb.launch(b)
        self.gb.wait()
        self.armed = False

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []


_WARNED_RAGGED = False


def shard_batch(x: torch.Tensor) -> torch.Tensor:
    """Contiguous shard of a global batch for this rank.  Every rank must hold the same number of sequences
    (allreduce_stats relies on it), so a ragged tail (n % world != 0, e.g. the last batch of an epoch with
    drop_last=False, tt:33-37) is trimmed to the largest multiple of the world size."""
    global _WARNED_RAGGED
    if not _ENABLED:
        return x
    w, r = world_size(), rank()
    n = x.shape[0]
    if n < w:
        raise ValueError(f"global batch {n} is smaller than the world size {w}")
    per = n // w
    if n % w != 0 and not _WARNED_RAGGED:
        _WARNED_RAGGED = True
        if r == 0:
            print(f"timegan_b200.dist: global batch {n} not divisible by {w} ranks; using the first {per * w} sequences")
    return x[r * per:(r + 1) * per]

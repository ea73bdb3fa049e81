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
_PEER = None           # PeerComm: all-reduce kernels over NVLink peer memory (CUDA ranks of one box)


def local_device() -> Optional[torch.device]:
    """The GPU of this rank: cuda:LOCAL_RANK (torchrun's convention, one process per GPU).  Makes it the current
    device, so that `device_autoselect()` and every allocation that follows land on it -- without this every rank
    of `torchrun -m timegan_b200.train_timegan` would sit on cuda:0."""
    if not torch.cuda.is_available():
        return None
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if local >= torch.cuda.device_count():
        raise RuntimeError(f"LOCAL_RANK={local} but only {torch.cuda.device_count()} CUDA device(s) are visible")
    torch.cuda.set_device(local)
    return torch.device("cuda", local)


def init(backend: Optional[str] = None, device: Optional[torch.device] = None):
    """Initialise from torchrun's environment (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*). Returns (rank, world).
    `device=None` under torchrun selects (and makes current) cuda:LOCAL_RANK."""
    global _GROUP, _ENABLED
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        _ENABLED = False
        return 0, 1
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if device is None and backend == "nccl":
        device = local_device()
    if not td.is_initialized():
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        td.init_process_group(backend=backend, **kw)
    _GROUP = td.group.WORLD
    _ENABLED = True
    if backend == "nccl" and device is not None and os.environ.get("TIMEGAN_B200_COMM", "peer") == "peer":
        enable_peer(device)
    return td.get_rank(), td.get_world_size()


def enable_peer(device, region_mb: Optional[int] = None):
    """Route every all-reduce of this package through csrc/peer_allreduce.cu (one multi-tensor kernel per bucket
    over NVLink peer memory, graph-capturable) instead of torch.distributed.  The process group stays in use for
    the rendezvous (IPC handle exchange) and barriers.  TIMEGAN_B200_COMM=nccl keeps the NCCL path."""
    global _PEER
    if _PEER is None:
        mb = int(region_mb or os.environ.get("TIMEGAN_B200_PEER_MB", "256"))
        _PEER = PeerComm(torch.device(device), _GROUP, mb << 20)
    return _PEER


def peer_comm():
    return _PEER


def begin_step(tag: str):
    """Start of one optimiser step (`tag` names the step function).  The k-th all-reduce issued after
    begin_step(tag) is the same call site in every step and on every rank -- that is what lets the peer
    all-reduce keep one staging area / flag set / replay counter per site."""
    if _PEER is not None:
        _PEER.begin_step(tag)


def enable(group=None):
    """Use an already-initialised process group (tests)."""
    global _GROUP, _ENABLED
    _GROUP = group if group is not None else td.group.WORLD
    _ENABLED = td.get_world_size(_GROUP) > 1


def disable():
    """Back to single-process behaviour.  The peer region (if any) stays mapped; enable() re-activates it."""
    global _GROUP, _ENABLED, _COMM_STREAM, _SHARD
    _GROUP, _ENABLED, _COMM_STREAM, _SHARD = None, False, None, None


def shutdown():
    """Collective teardown of the peer-memory communicator (before destroy_process_group)."""
    global _PEER
    if _PEER is not None:
        _PEER.check_status()
        _PEER.close()
        _PEER = None


def is_enabled() -> bool:
    return _ENABLED


def world_size() -> int:
    return td.get_world_size(_GROUP) if _ENABLED else 1


def rank() -> int:
    return td.get_rank(_GROUP) if _ENABLED else 0


_SHARD = None          # (sequences on this rank, sequences in the global batch) of the batch in flight, or None


def global_count(count) -> float:
    """Global-batch size of a quantity whose LOCAL size is `count` (sequences, rows, elements: anything
    proportional to the number of sequences).  Every rank sees the global batch before it is sharded
    (`shard_batch`), so the ratio global / local is known on the host without a synchronisation -- also for the
    ragged last batch of an epoch (drop_last=False, tt:33-37), where the ranks hold shards of different sizes."""
    if not _ENABLED:
        return float(count)
    if _SHARD is not None:
        return float(count) * _SHARD[1] / _SHARD[0]
    return float(count) * world_size()


def allreduce_stats(t: torch.Tensor, count) -> Tuple[torch.Tensor, float]:
    """SUM-all-reduce a small statistics tensor together with its sample count.

    Returns (reduced tensor, global count).  Single process: identity.  Sums are count-weighted by construction
    (every rank adds the statistics of the sequences it holds), and the global count follows from the shard
    bookkeeping of `shard_batch` (`global_count`), so ragged shards need no extra collective.
    """
    if not _ENABLED:
        return t, float(count)
    buf = t.detach().clone().contiguous()
    if _PEER is not None and buf.is_cuda:
        _PEER.allreduce_([buf])
    else:
        td.all_reduce(buf, op=td.ReduceOp.SUM, group=_GROUP)
    return buf, global_count(count)


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


class PeerComm:
    """Host side of csrc/peer_allreduce.cu for the ranks of ONE box.

    Every rank cudaMallocs one region, the 64-byte IPC handles travel through the process group once, and from
    then on an all-reduce is one kernel launch on the current stream.  Region layout (identical on every rank):
    [flag words | staging areas]; a call site -- key (step tag, index of the call within the step, tensor
    sizes) -- gets its slice of both on first use, plus a local {epoch, counter} pair in `self.epochs`."""

    FLAG_BYTES = 4 << 20
    MAX_SITES = 1024

    def __init__(self, device: torch.device, group, region_bytes: int):
        import ctypes as C
        from ._lib import lib, check
        self._C, self._lib, self._check = C, lib, check
        self.device = device
        self.group = group
        self.rank, self.world = td.get_rank(group), td.get_world_size(group)
        if self.world > 8:
            raise ValueError("peer all-reduce covers the (<= 8) GPUs of one NVSwitch box")
        self.region_bytes = int(region_bytes)
        with torch.cuda.device(device):
            own = C.c_void_p()
            check(lib.tg_peer_alloc(C.byref(own), self.region_bytes), "tg_peer_alloc")
            handle = C.create_string_buffer(64)
            check(lib.tg_peer_export(own, handle), "tg_peer_export")
            handles = [None] * self.world
            td.all_gather_object(handles, handle.raw, group=group)
            self.regions = (C.c_void_p * self.world)()
            self._own = own
            for r in range(self.world):
                if r == self.rank:
                    self.regions[r] = own.value
                else:
                    peer = C.c_void_p()
                    check(lib.tg_peer_open(handles[r], C.byref(peer)), f"tg_peer_open(rank {r})")
                    self.regions[r] = peer.value
        td.barrier(group=group)
        self.epochs = torch.zeros(self.MAX_SITES, 2, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.sites = {}
        self.flag_cursor, self.data_cursor = 0, self.FLAG_BYTES
        self.tag, self.k = "", 0
        self.chunk = int(lib.tg_peer_chunk_floats())

    def begin_step(self, tag: str):
        self.tag, self.k = tag, 0

    def _site(self, sizes):
        C = self._C
        key = (self.tag, self.k, tuple(sizes))
        self.k += 1
        st = self.sites.get(key)
        if st is None:
            arr = (C.c_longlong * len(sizes))(*sizes)
            fb = C.c_size_t()
            db = int(self._lib.tg_peer_site_bytes(len(sizes), arr, self.world, C.byref(fb)))
            fb = (int(fb.value) + 255) // 256 * 256
            if len(self.sites) >= self.MAX_SITES or self.flag_cursor + fb > self.FLAG_BYTES or \
                    self.data_cursor + db > self.region_bytes:
                raise RuntimeError(
                    f"peer region exhausted ({len(self.sites)} call sites, {self.data_cursor + db} of "
                    f"{self.region_bytes} B): call dist.begin_step() once per optimiser step, or raise "
                    "TIMEGAN_B200_PEER_MB")
            st = dict(idx=len(self.sites), data_off=self.data_cursor, flag_off=self.flag_cursor, sizes=arr)
            self.flag_cursor += fb
            self.data_cursor += (db + 255) // 256 * 256
            self.sites[key] = st
        return st

    def allreduce_(self, tensors):
        """SUM `tensors` (contiguous fp32 CUDA tensors) in place across the ranks, on the current stream."""
        C = self._C
        for i0 in range(0, len(tensors), 48):
            group = tensors[i0:i0 + 48]
            for t in group:
                if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                    raise RuntimeError("peer all-reduce needs contiguous fp32 CUDA tensors")
            st = self._site([t.numel() for t in group])
            ptrs = (C.c_void_p * len(group))(*[t.data_ptr() for t in group])
            self._check(self._lib.tg_peer_allreduce(
                torch.cuda.current_stream().cuda_stream, self.rank, self.world, self.regions, st["data_off"],
                st["flag_off"], self.epochs[st["idx"]].data_ptr(), self.status.data_ptr(), len(group), ptrs,
                st["sizes"]), "tg_peer_allreduce")

    def check_status(self):
        """Raises if a kernel gave up waiting for a peer (synchronises)."""
        if int(self.status.item()) != 0:
            raise RuntimeError("peer all-reduce timed out waiting for another rank")

    def close(self):
        """Unmap the peers' regions and free this rank's (collective: every rank must call it; no all-reduce may be
        in flight).  Safe to call twice."""
        if self.regions is None:
            return
        torch.cuda.synchronize(self.device)
        td.barrier(group=self.group)          # nobody is still reading this rank's staging areas
        with torch.cuda.device(self.device):
            for r in range(self.world):
                if r != self.rank and self.regions[r]:
                    self._check(self._lib.tg_peer_close(self.regions[r]), "tg_peer_close")
            td.barrier(group=self.group)      # every mapping of this rank's region is gone
            self._check(self._lib.tg_peer_free(self._own), "tg_peer_free")
        self.regions = None
        self.sites.clear()


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
        if grads[0].is_cuda and _PEER is not None:
            # one multi-tensor kernel: gathers the gradients out of their own tensors, exchanges them over NVLink
            # peer memory and writes the sums back in place -- nothing to flatten or scatter back
            cs = comm_stream()
            cs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cs):
                _PEER.allreduce_(grads)
                for g in grads:
                    g.record_stream(cs)
            self._pending.append((None, grads))
        elif grads[0].is_cuda:
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
        cuda = self._pending[0][1][0].is_cuda
        if cuda:
            cs = comm_stream()
            with torch.cuda.stream(cs):
                for flat, grads in self._pending:
                    if flat is not None:
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


def shard_bounds(n: int, w: int, r: int) -> Tuple[int, int]:
    """[start, end) of rank r's contiguous shard of n sequences over w ranks: the first n % w ranks hold one more."""
    per, extra = divmod(n, w)
    start = r * per + min(r, extra)
    return start, start + per + (1 if r < extra else 0)


def shard_batch(x: torch.Tensor) -> Optional[torch.Tensor]:
    """Contiguous shard of a global batch for this rank.  Every sequence of the batch is used, like the
    reference's single process does with drop_last=False (tt:33-37): a ragged batch (n % world != 0) gives the
    first n % world ranks one sequence more, and the loss / gradient conventions stay exact because every
    statistic is a count-weighted sum over the global count (`global_count`).  A batch with fewer sequences
    than ranks returns None on EVERY rank (all ranks see the same n): the callers skip it -- a rank without a
    single sequence cannot take part in the step's kernels."""
    global _SHARD
    if not _ENABLED:
        return x
    w, r = world_size(), rank()
    n = x.shape[0]
    if n < w:
        _SHARD = None
        return None
    a, b = shard_bounds(n, w, r)
    _SHARD = (b - a, n)
    return x[a:b]


def clear_shard():
    """Forget the shard bookkeeping (callers that hand every rank its own equal-sized batch, e.g. bench.py)."""
    global _SHARD
    _SHARD = None

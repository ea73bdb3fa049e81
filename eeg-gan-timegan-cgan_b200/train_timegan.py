"""B200-native drop-in for timeGAN/train_timegan.py: same function names, signatures, defaults, CLI flags,
log/checkpoint schema -- every GRU pass, loss and optimiser update runs in the sm_100a kernels of csrc/.

Reference map (file:line under /root/reference/timeGAN/train_timegan.py):
    set_seeds 21-26 · device_autoselect 29-30 · make_loader 33-37 · smooth_labels 40-43 · add_instance_noise 46-47
    adaptive_dims 50-55 · save_ckpt 58-61 · sample_noise 64-65 · losses 70-126 · phase_autoencoder 131-144
    phase_supervisor 147-163 · disc_step 166-225 · gen_step 228-276 · train_single_npz 281-422 · main 427-495

What is different by design (B200-first, results identical):
  * disc_step never builds a double-backward graph: R1 (tt:198-202) is evaluated as dX-only BPTT -> tangent
    forward -> reverse-over-tangent, real and fake halves share one 2B-batch D forward.
  * no host synchronisation inside a step: the throttle scale (tt:204-216) is computed on the device; with
    `sync=False` the step functions return device scalars instead of Python floats.
  * gen_step treats the discriminator's weights as constants (the reference computes and discards their
    gradients, tt:267 / SURVEY.md App. D item 8).
  * optional keyword-only extras, all defaulting to the reference's behaviour: `noise=` (replay host draws for
    parity), `sync=`, `z_dim=/hidden_dim=` overrides of adaptive_dims, `log_every=`, data-parallel training when
    launched under torchrun.
There is no CPU path: a CPU tensor or a missing libtimegan_b200.so raises.
"""
import csv
import math
import os
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader, TensorDataset

from . import dist as _dist
from . import head as _head
from . import losses as _losses
from . import noise as _noise
from . import ops
from .optim import FusedAdam, clip_and_step
from .timegan_model import TimeGAN


# ---------------------- Utilities (tt:21-65) ----------------------

def set_seeds(seed: int = 42):
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def device_autoselect():
    if not torch.cuda.is_available():
        raise RuntimeError("timegan_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def make_loader(X: np.ndarray, batch_size: int) -> DataLoader:
    tens = torch.tensor(X, dtype=torch.float32)
    ds = TensorDataset(tens)
    return DataLoader(ds, batch_size=batch_size, shuffle=True, drop_last=False, pin_memory=True, num_workers=0)


class DeviceLoader:
    """`make_loader` (tt:33-37) with the dataset resident in HBM.

    At B200 speeds a joint step takes ~24 ms, and the whole dataset (N x 768 x 14 fp32: 11 MB for N=256, 1.4 GB for
    N=32 768) fits the device many times over, so the per-batch collate + pin + H2D copy of the DataLoader is pure
    overhead: X is uploaded once and a batch is one on-device gather.  The ORDER is the reference's: an epoch's
    permutation is drawn exactly like DataLoader(shuffle=True) draws it -- `_base_seed` then the RandomSampler's
    seed, two int64 draws from the global CPU generator, then randperm from a private generator -- so runs that
    replay the CPU random stream (noise="host") see the same batches AND leave the global generator in the same
    state as the reference.  drop_last=False: the ragged last batch is yielded as is.  Yields 1-tuples like a
    DataLoader over a TensorDataset."""

    def __init__(self, X, batch_size: int, device):
        self.X = torch.as_tensor(X, dtype=torch.float32).to(device).contiguous()
        self.batch_size = int(batch_size)
        self.device = self.X.device

    def __len__(self):
        return (self.X.shape[0] + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.X.shape[0]
        torch.empty((), dtype=torch.int64).random_()                       # DataLoader iterator's _base_seed
        seed = int(torch.empty((), dtype=torch.int64).random_().item())    # RandomSampler's private seed
        gen = torch.Generator()
        gen.manual_seed(seed)
        perm = torch.randperm(n, generator=gen).to(self.device, non_blocking=True)
        for i in range(0, n, self.batch_size):
            yield (self.X.index_select(0, perm[i:i + self.batch_size]),)


class HostReplayNoise:
    """Draws every random tensor from torch's global CPU generator with the reference's calls, in the
    reference's order, then moves it to the device.  A run seeded like a CPU run of the
    reference therefore sees bit-identical noise -- the parity mode of the tests and golden curves."""

    def __init__(self, device):
        self.device = device

    def rand(self, *shape):
        return torch.rand(*shape).to(self.device)

    def begin(self):
        pass

    def end(self):
        pass

    def randn_like(self, h, time_major: bool = False):
        """torch.randn_like on a CPU tensor with the strides the reference's tensor has at that call site:
        nn.GRU(batch_first=True) returns a transposed view of a (T,B,H) buffer on the CPU, and randn_like
        fills a non-contiguous tensor through a different (serial) sampling path than a contiguous one, so
        the replay has to reproduce the layout, not just the shape."""
        if time_major and h.dim() == 3:
            B, T, H = h.shape
            like = torch.empty(T, B, H).transpose(0, 1)
        else:
            like = torch.empty(h.shape)
        return torch.randn_like(like).to(self.device)


class _DeviceNoiseAdapter:
    """Production noise: Philox kernels of csrc/rng.cu (statistically equivalent, no host work)."""

    def __init__(self, src: "_noise.DeviceNoise"):
        self.src, self.device = src, src.device

    def rand(self, *shape):
        return self.src.rand(*shape)

    def randn_like(self, h, time_major: bool = False):
        return self.src.randn(h.shape)

    def add_randn(self, h, std, out=None):
        """h + std * N(0,1) as one kernel (same Philox positions as randn_like of the same shape)."""
        return self.src.add_randn(h, std, out=out)

    def begin(self):
        self.src.begin()

    def end(self):
        self.src.end()

    def hold(self):
        self.src.hold()

    def release(self):
        self.src.release()


_DEFAULT_NOISE: Dict[str, _DeviceNoiseAdapter] = {}


def device_noise(seed: int, device):
    """An independent on-device noise source (Philox stream `seed`) usable as the `noise=` argument."""
    return _DeviceNoiseAdapter(_noise.DeviceNoise(seed, device))


def _noise_source(noise, device):
    if noise is not None:
        return noise
    key = str(device)
    if key not in _DEFAULT_NOISE:
        _DEFAULT_NOISE[key] = _DeviceNoiseAdapter(_noise.DeviceNoise(torch.initial_seed() + 7919 * _dist.rank(), device))
    return _DEFAULT_NOISE[key]


def smooth_labels(size: int, smooth: float, device, noise=None) -> Tuple[torch.Tensor, torch.Tensor]:
    nz = _noise_source(noise, device)
    real = (1.0 - smooth) + smooth * nz.rand(size, 1)
    fake = smooth * nz.rand(size, 1)
    return real, fake


class _AddNoise(torch.autograd.Function):
    """h + std * N(0,1) drawn and added by one kernel; d/dh = identity, the noise is a constant (tt:46-47)."""

    @staticmethod
    def forward(ctx, h, std, src):
        return src.add_randn(h.detach(), std)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


def add_instance_noise(h: torch.Tensor, std: float, noise=None, time_major: bool = False, out=None) -> torch.Tensor:
    """tt:46-47.  `time_major` only matters to host-replayed noise (layout of the reference's tensor).
    `std` may be a 0-dim device tensor (CUDA-graph replay: the value changes every step, the graph does not).
    `out` (no-grad callers only): a contiguous view that receives the result, e.g. one half of D's 2B batch."""
    src = _noise_source(noise, h.device)
    if hasattr(src, "add_randn"):                       # on-device Philox: draw + scale + add in one launch
        if out is not None or not (torch.is_grad_enabled() and h.requires_grad):
            return src.add_randn(h, std, out=out)
        if not torch.is_tensor(std) and std <= 0:
            return h
        return _AddNoise.apply(h, std, src)
    if not torch.is_tensor(std) and std <= 0:
        res = h
    else:
        res = h + std * src.randn_like(h, time_major)
    if out is not None:
        out.copy_(res)
        return out
    return res


def _latent_is_gru_view(model) -> bool:
    """True when supervisor output is the raw GRU output (proj is Identity, tm:79), i.e. a time-major view in
    the reference; with a Linear projection the reference's tensor is batch-major contiguous."""
    return isinstance(model.supervisor.proj, nn.Identity)


def adaptive_dims(x_dim: int, seq_len: int) -> Tuple[int, int]:
    z = max(16, min(64, x_dim * 2))
    h = max(32, min(128, x_dim * 4))
    if seq_len > 800:
        z = min(64, z + 8)
        h = min(128, h + 16)
    return z, h


def save_ckpt(path: Path, model: TimeGAN, optG, optD, step: int, meta: Dict):
    state = {"step": step, "model": model.state_dict(), "optG": optG.state_dict(), "optD": optD.state_dict(),
             "meta": meta}
    torch.save(state, path)


class AsyncCheckpointer:
    """`save_ckpt` (tt:58-61) without stalling the training stream.

    save() snapshots every tensor of the checkpoint dict into pinned host memory with non-blocking copies on the
    current stream (so the snapshot is the state at that point of the stream, whatever the GPU does next), records
    an event and hands the dict to a writer thread, which waits for the event and calls torch.save -- same dict,
    same schema {"step","model","optG","optD","meta"}, loadable by the reference's tools.  One write at a time:
    a new save() first joins the previous one.  wait() joins the writer (call before reading the file)."""

    def __init__(self):
        self._thread = None
        self._error = None

    @staticmethod
    def _snapshot(obj):
        if torch.is_tensor(obj):
            if obj.is_cuda:
                host = torch.empty(obj.shape, dtype=obj.dtype, pin_memory=True)
                host.copy_(obj.detach(), non_blocking=True)
                return host
            return obj.detach().clone()
        if isinstance(obj, dict):
            return type(obj)((k, AsyncCheckpointer._snapshot(v)) for k, v in obj.items())
        if isinstance(obj, (list, tuple)):
            return type(obj)(AsyncCheckpointer._snapshot(v) for v in obj)
        return obj

    def save(self, path: Path, model: TimeGAN, optG, optD, step: int, meta: Dict):
        self.save_state(path, {"step": step, "model": model.state_dict(), "optG": optG.state_dict(),
                               "optD": optD.state_dict(), "meta": dict(meta)})

    def save_state(self, path: Path, state: Dict):
        """Write an already assembled checkpoint dict (same schema) behind the training stream."""
        self.wait()
        state = self._snapshot(state)
        ev = None
        if torch.cuda.is_available():
            ev = torch.cuda.Event()
            ev.record()

        def _write():
            try:
                if ev is not None:
                    ev.synchronize()
                tmp = Path(str(path) + ".tmp")
                torch.save(state, tmp)
                os.replace(tmp, path)          # never leave a half-written checkpoint behind
            except Exception as e:  # pragma: no cover
                self._error = e

        import threading
        self._thread = threading.Thread(target=_write, daemon=True)
        self._thread.start()

    def wait(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self._error is not None:
            e, self._error = self._error, None
            raise e


class BestSnapshot:
    """The reference's best-checkpoint rule (tt:410-413: `if g_total < best_ckpt_loss: save_ckpt(best_path, ...)`,
    evaluated after EVERY step) without a host round trip per step.

    A second copy of everything a checkpoint holds -- weights, spectral-norm buffers, both optimisers' Adam moments
    -- lives in HBM.  update(g_total, step) is two launches (csrc/optim.cu `snapshot_if_better`): when the step's
    generator loss beats the best so far, the live tensors are copied into the snapshot and (best, best_step) are
    updated, all on the device.  The host reads `best_step` together with the logged scalars whenever it flushes
    them, and only then -- if it moved -- hands the snapshot to the AsyncCheckpointer.  The file therefore holds
    exactly the weights the reference would have saved, written at most once per flush window."""

    def __init__(self, model: TimeGAN, optG, optD, device, best: float = math.inf):
        import ctypes as C
        self._C = C
        for opt in (optG, optD):
            for g in opt.param_groups:
                for p in g["params"]:
                    opt._init_state(p)
        self.model, self.optG, self.optD = model, optG, optD
        self.src = [t for t in model.state_dict().values() if t.is_cuda and t.dtype == torch.float32]
        self.model_keys = [k for k, t in model.state_dict().items() if t.is_cuda and t.dtype == torch.float32]
        self.n_model = len(self.src)
        self.moment_owner = []
        for name, opt in (("optG", optG), ("optD", optD)):
            for g in opt.param_groups:
                for p in g["params"]:
                    for k in ("exp_avg", "exp_avg_sq"):
                        self.src.append(opt.state[p][k])
                        self.moment_owner.append((name, p, k))
        self.dst = [torch.empty_like(t) for t in self.src]
        n = len(self.src)
        self._dst_ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in self.dst])
        self._src_ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in self.src])
        self._sizes = (C.c_longlong * n)(*[t.numel() for t in self.src])
        self.best = torch.full((1,), float(best), dtype=torch.float32, device=device)
        self.best_step = torch.full((1,), -1.0, dtype=torch.float32, device=device)
        self.written_step = -1

    def update(self, g_total: torch.Tensor, step: int):
        from ._lib import lib, check, ptr, stream_ptr
        val = g_total.detach().reshape(1).float().contiguous()
        check(lib.tg_snapshot_if_better(stream_ptr(), len(self.src), self._dst_ptrs, self._src_ptrs, self._sizes,
                                        ptr(val), ptr(self.best), ptr(self.best_step), float(step)),
              "tg_snapshot_if_better")

    def checkpoint(self, best_step: int, rewind: int, meta: Dict, lr_at) -> Dict:
        """The checkpoint dict of the snapshot (schema of tt:58-61).  `rewind` = optimiser updates performed since the
        snapshot was taken: the live per-parameter `step` counters are rewound by it; `lr_at(name, step)` gives the
        scheduled learning rate of optimiser `name` after `step` GAN steps."""
        snap = dict(zip(self.model_keys, self.dst[:self.n_model]))
        model_sd = type(self.model.state_dict())((k, snap.get(k, v)) for k, v in self.model.state_dict().items())
        moments = {(name, id(p), k): t for (name, p, k), t in zip(self.moment_owner, self.dst[self.n_model:])}
        out = {"step": best_step, "model": model_sd, "meta": dict(meta, best=True)}
        for name, opt in (("optG", self.optG), ("optD", self.optD)):
            sd = opt.state_dict()
            idx = 0
            for g, g_live in zip(sd["param_groups"], opt.param_groups):
                g["lr"] = lr_at(name, best_step)
                for p in g_live["params"]:
                    st = dict(sd["state"].get(idx, {}))
                    if st:
                        st["exp_avg"] = moments[(name, id(p), "exp_avg")]
                        st["exp_avg_sq"] = moments[(name, id(p), "exp_avg_sq")]
                        st["step"] = torch.tensor(float(st["step"]) - float(rewind), dtype=torch.float32)
                        sd["state"][idx] = st
                    idx += 1
            out[name] = sd
        return out


def sample_noise(batch_size: int, seq_len: int, z_dim: int, device, noise=None):
    return _noise_source(noise, device).rand(batch_size, seq_len, z_dim)


# ---------------------- Losses (tt:70-126) ----------------------

bce = _losses.bce
recon_loss = _losses.recon_loss
sup_loss = _losses.sup_loss
sup_loss_fake = _losses.sup_loss_fake


def acf_loss_torch(x_gen: torch.Tensor, x_real: torch.Tensor, max_lag: int) -> torch.Tensor:
    return _losses.cov_acf_losses(x_gen, x_real, max_lag, need_cov=False, need_acf=True)[1]


def _params(*modules):
    out = []
    for m in modules:
        out += list(m.parameters())
    return out


def _zero_grads(opt):
    opt.zero_grad(set_to_none=True)


def _reducer(model, key: str, modules):
    """Per-(model, parameter set) gradient reducer with one bucket per network (DP only; hooks are registered once)."""
    if not _dist.is_enabled():
        return None
    cache = model.__dict__.setdefault("_tg_reducers", {})
    if key not in cache:
        cache[key] = _dist.GradReducer([list(m.parameters()) for m in modules])
    return cache[key]


def _reduce_and_step(opt, params, clip, reducer=None):
    """DP: SUM-all-reduce the local gradient contributions (dist.py), then clip_grad_norm_ + Adam.  With a
    reducer the buckets were launched from autograd hooks while BPTT was still running; here they are joined."""
    if reducer is not None:
        reducer.finish()
    elif _dist.is_enabled():
        b = _dist.GradBuckets()
        b.launch(params)
        b.wait()
    clip_and_step(opt, params, clip)


_SIDE_STREAMS: Dict[str, list] = {}
_CONCURRENT = True   # run the independent GRU stacks of a step on side streams (set False to serialise everything)


def set_concurrency(on: bool):
    global _CONCURRENT
    _CONCURRENT = bool(on)


# GraphedJointStep: let gen_step's D-independent forward passes overlap disc_step (TIMEGAN_B200_OVERLAP_STEPS=0: off)
_OVERLAP_STEPS = os.environ.get("TIMEGAN_B200_OVERLAP_STEPS", "1") != "0"


def set_step_overlap(on: bool):
    global _OVERLAP_STEPS
    _OVERLAP_STEPS = bool(on)


# side streams whose work sits on the step's critical path get a higher CUDA priority: when more CTAs are pending than
# the SMs have slots (three or four recurrent kernels of 256 CTAs in flight), theirs are dispatched first
# (pool index -> priority; 0 is CUDA's lowest).  4 = D's path inside gen_step, 5 = disc_step's main line.  A third level
# (the recon path of gen_step below everything else, its backward issued from a helper stream so that S and G need not
# wait for it) measured the same 20.5 ms/step: the step is bound by SM throughput, not by the order of the fill-in work.
_STREAM_PRIORITY = {4: -1, 5: -1} if os.environ.get("TIMEGAN_B200_STREAM_PRIO", "1") != "0" else {}


def _side_stream(device, k: int):
    pool = _SIDE_STREAMS.setdefault(str(device), [])
    while len(pool) <= k:
        pool.append(torch.cuda.Stream(device=device, priority=_STREAM_PRIORITY.get(len(pool), 0)))
    return pool[k]


class _Fork:
    """`with _Fork(device, k):` runs the block on side stream k after it has caught up with the current stream.
    The recurrent kernels are latency-bound (768 dependent steps per layer pass) and leave most of each SM idle,
    so independent stacks of a step (E || G->S, D || R, and their BPTTs, which autograd replays on the stream of
    the forward) overlap almost for free.  join() makes the current stream wait for every forked stream."""

    def __init__(self, device, k: int):
        self.main = torch.cuda.current_stream(device)
        if _CONCURRENT:
            self.side = _side_stream(device, k)
        else:
            self.side = self.main
        self.ctx = None

    def __enter__(self):
        if self.side is not self.main:
            self.side.wait_stream(self.main)
            ops.fork_begin()               # until join(): recurrent passes on both streams run beside each other
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        return self.ctx.__exit__(*a)

    def join(self):
        if self.side is not self.main:
            self.main.wait_stream(self.side)
            ops.fork_end()


def _as_float(t, sync):
    return t.item() if sync else t.detach()


# ---------------------- Training phases ----------------------

def phase_autoencoder(model: TimeGAN, loader: DataLoader, device, optER: optim.Optimizer, clip: float, epochs: int,
                      log):
    """tt:131-144."""
    model.train()
    params = _params(model.embedder, model.recovery)
    for ep in range(1, epochs + 1):
        epoch_loss = torch.zeros((), device=device)
        n = 0
        for (x_batch,) in loader:
            x = _dist.shard_batch(x_batch.to(device, non_blocking=True))
            if x is None:              # fewer sequences than ranks: skipped on every rank alike
                continue
            _dist.begin_step("AE")
            x_tilde = model.reconstruct(x)
            loss = recon_loss(x, x_tilde)
            _zero_grads(optER)
            red = _reducer(model, "ER", (model.recovery, model.embedder))
            if red is not None:
                red.arm()
            loss.backward()
            _reduce_and_step(optER, params, clip, red)
            epoch_loss += loss.detach() * x_batch.size(0)
            n += x_batch.size(0)
        log(f"[AE] epoch {ep}/{epochs}  recon={epoch_loss.item() / n:.5f}")


def phase_supervisor(model: TimeGAN, loader: DataLoader, device, optS: optim.Optimizer, clip: float, epochs: int, log):
    """tt:147-163."""
    model.train()
    params = _params(model.supervisor)
    for ep in range(1, epochs + 1):
        epoch_loss = torch.zeros((), device=device)
        n = 0
        for (x_batch,) in loader:
            x = _dist.shard_batch(x_batch.to(device, non_blocking=True))
            if x is None:
                continue
            _dist.begin_step("SUP")
            with torch.no_grad():
                h = model.encode(x)
            h_in, h_tgt = h[:, :-1, :].contiguous(), h[:, 1:, :].contiguous()
            h_pred = model.supervisor(h_in)
            loss = _losses.mse_loss(h_pred, h_tgt)
            _zero_grads(optS)
            red = _reducer(model, "S", (model.supervisor,))
            if red is not None:
                red.arm()
            loss.backward()
            _reduce_and_step(optS, params, clip, red)
            epoch_loss += loss.detach() * x_batch.size(0)
            n += x_batch.size(0)
        log(f"[SUP] epoch {ep}/{epochs}  sup={epoch_loss.item() / n:.5f}")


def disc_step(model: TimeGAN, x, device, optD, label_smooth, inst_noise_std, clip, schedulerD=None,
              r1_gamma: float = 1.0, target_acc: float = 0.55, band: float = 0.10, *, noise=None, sync: bool = True):
    """Discriminator update with R1 and the soft throttle (tt:166-225).  Returns (loss, acc)."""
    D = model.discriminator
    D.train()
    _dist.begin_step("D")
    nz = _noise_source(noise, device)
    nz.begin()
    B, T = x.size(0), x.size(1)
    gru = D.rnn.rnn
    wd = [w.detach() for w in gru.layer_weights()]

    with torch.no_grad():
        z = nz.rand(B, T, model.embedder.rnn.rnn.hidden_size)                        # tt:179
        fork_e = _Fork(device, 0)
        with fork_e:
            h_real = model.encode(x)                                                 # tt:175-176  (|| G -> S)
        h_fake = model.refine_latent(model.gen_latent(z))                            # tt:180-181 (forward only)
        fork_e.join()
        # one 2B-batch pass of the D stack over [real ; fake] (tt:191-193): the noisy latents are written straight
        # into the two halves of its input
        h_in = torch.empty((2 * B,) + tuple(h_real.shape[1:]), dtype=torch.float32, device=h_real.device)
        h_real_n = add_instance_noise(h_real, inst_noise_std, nz, True, out=h_in[:B])              # tt:184
        h_fake_n = add_instance_noise(h_fake, inst_noise_std, nz, _latent_is_gru_view(model), out=h_in[B:])   # tt:185
        y_real, y_fake = smooth_labels(B, label_smooth, device, nz)                  # tt:188
        masks = gru.dropout_masks(h_in)
        y_all, saves = ops.stack_forward(h_in, wd, save=True, masks=masks)
        last = y_all[:, -1, :]                       # (2B,H) view: row stride T*H, read in place by the head kernels
        H = last.shape[1]

        # head forward (tm:96-98 for both calls, tt:196 BCE sums, tt:205-208 accuracy counts): ONE kernel; the legacy
        # spectral-norm hook's two power iterations (one per D call) update fc.weight_u / fc.weight_v in place
        hs = _head.forward(D, last, torch.cat([y_real.reshape(-1), y_fake.reshape(-1)]), 2)
        stats, _ = _dist.allreduce_stats(hs.stats, B)          # data parallel: the sums of the global batch
        Bg = _dist.global_count(B)
        # throttle scale (tt:209-215) on the device, the R1 seed d(sum d_real)/d y_last (tt:200) and the fake half's
        # input gradient (it needs nothing else, so its BPTT starts right away, beside the whole R1 chain)
        scal, seed, gyf = _head.seed(hs, stats, Bg, target_acc, band, need_seed=r1_gamma > 0.0)
        grads = ops.alloc_like_flat(wd)
        r1 = None
        if r1_gamma > 0.0:                                                           # tt:199-202
            saves_r = [sv.narrow(0, B) for sv in saves]
            masks_r = None if masks is None else [m[:B] for m in masks]
            saves_f = [sv.narrow(B, B) for sv in saves]
            masks_f = None if masks is None else [m[B:] for m in masks]
            grads_f = ops.alloc_like_flat(wd)
            fork_f = _Fork(device, 0)
            with fork_f:                           # BPTT of the fake half || (dX-only BPTT -> tangent forward -> reverse)
                ops.stack_backward(gyf, saves_f, wd, need_dx=False, need_dw=True, dy_last=True, grads=grads_f,
                                   accumulate=False, masks=masks_f)
            v, _ = ops.stack_backward(seed, saves_r, wd, need_dx=True, need_dw=False, dy_last=True, masks=masks_r)
            r1 = _dist.global_mean(_losses.sumsq(v), B)
            ydot, tsaves = ops.stack_jvp_forward(v, saves_r, wd, masks=masks_r)
            hd_last = ydot[:, -1, :]
            # d(scale * (loss_bce + 0.5*gamma*r1))/d{y_last(real), tangent, fc}: 0.5*gamma*r1 enters through
            # (gamma/Bg) * sdot, sdot = sum_b p(1-p) (hd_b . wbar)   
            gyr, ghd, gw, gb, loss_val = _head.backward(D, hs, last, hd_last, scal, r1, Bg, r1_gamma)
            grads.flat.zero_()            # one launch for the whole stack (the buffers are views of one allocation)
            ops.stack_jvp_backward(gyr, ghd, saves_r, tsaves, wd, grads, accumulate=True, masks=masks_r)
            fork_f.join()
            grads.flat.add_(grads_f.flat)
        else:
            gyr, _, gw, gb, loss_val = _head.backward(D, hs, last, None, scal, None, Bg, 0.0)
            ops.stack_backward(torch.cat([gyr, gyf], 0), saves, wd, need_dx=False, need_dw=True, dy_last=True,
                               grads=grads, accumulate=False, masks=masks)
        acc = scal[1]
        loss_val = loss_val.reshape(())
    _zero_grads(optD)                                                                # tt:218-221
    for p, g in zip(gru.layer_weights(), grads):
        p.grad = g
    D.fc.weight_orig.grad, D.fc.bias.grad = gw, gb
    _reduce_and_step(optD, _params(D), clip)
    nz.end()
    if schedulerD is not None:
        schedulerD.step()
    return _as_float(loss_val, sync), _as_float(acc, sync)


def gen_step(model: TimeGAN, x, device, optG, alpha_sup, beta_rec, inst_noise_std, clip, schedulerG=None,
             gamma_cov: float = 0.0, gamma_acf: float = 0.0, acf_max_lag: int = 32, *, noise=None,
             sync: bool = True, d_ready=None, fork_base: int = 0):
    """Generator/supervisor/embedder/recovery update (tt:228-276).  Returns the six logged losses.

    `d_ready` (a CUDA event) marks the point where the discriminator's weights are final: everything before the D
    forward pass -- E -> R, G -> S -> R, the moment losses -- does not read them, so a caller that runs this function
    on its own stream (GraphedJointStep) lets that part overlap the tail of disc_step and passes the event of D's
    optimiser step here.  `fork_base`: first side-stream index this call may use."""
    model.generator.train(); model.supervisor.train(); model.embedder.train(); model.recovery.train()
    _dist.begin_step("G")
    nz = _noise_source(noise, device)
    nz.begin()
    B, T = x.size(0), x.size(1)

    z = nz.rand(B, T, model.embedder.rnn.rnn.hidden_size)                            # tt:235
    fork_rec = _Fork(device, fork_base)
    with fork_rec:                                   # E -> R on the real batch runs beside G -> S
        x_tilde = model.reconstruct(x)                                               # tt:247-248
        g_rec = recon_loss(x, x_tilde)
    e_hat = model.gen_latent(z)
    h_hat = model.refine_latent(e_hat)
    # Split backward (only when the caller overlaps this step with disc_step, d_ready given): everything that does not
    # pass through D -- recon (R <- E), the moment losses (R on the generated latents) and the first-difference term --
    # is back-propagated BEFORE waiting for D's update, down to h_hat, where its gradient is parked; D's path adds to it
    # afterwards and only then S and G are unrolled (one BPTT each, as in the single backward).  9 of gen_step's 18 BPTT
    # passes and their weight gradients thereby run beside disc_step's R1 chain instead of after it.
    split = d_ready is not None
    h_use = h_hat.detach().requires_grad_(True) if split else h_hat
    if split:
        # D's path gets its own cut point, created on the stream D will run on: autograd accumulates a leaf's gradient
        # on the stream the leaf was first used on, and that must not be this function's main line (it would then wait
        # for D's backward before starting the early backward below)
        fork_adv = _Fork(device, fork_base + 2)
        with fork_adv:
            h_adv = h_hat.detach().requires_grad_(True)
            d_in = add_instance_noise(h_adv, inst_noise_std, nz, _latent_is_gru_view(model))   # tt:240 (noise drawn here)
    else:
        d_in = add_instance_noise(h_hat, inst_noise_std, nz, _latent_is_gru_view(model))       # tt:240 (noise drawn here)
    cov_term = torch.zeros((), device=device)
    acf_term = torch.zeros((), device=device)
    fork_dec = _Fork(device, fork_base + 1)
    with fork_dec:                                   # R on the generated latents runs beside D
        x_hat = model.decode(h_use)                                                  # tt:251
        if gamma_cov > 0 or gamma_acf > 0:                                           # tt:254-263
            cov_term, acf_term = _losses.cov_acf_losses(x_hat, x, acf_max_lag, need_cov=gamma_cov > 0,
                                                        need_acf=gamma_acf > 0)
    g_sup = sup_loss_fake(h_use)                                                     # tt:244
    # buckets in the order their BPTT finishes (recovery, embedder first; then supervisor, generator): each
    # bucket's all-reduce overlaps the BPTT of the networks that are still running
    red = _reducer(model, "G", (model.recovery, model.embedder, model.supervisor, model.generator))
    if split:
        # D's forward and dX-only backward on their own stream: they wait for D's update and for d_in, not for the
        # early backward below (whatever is left of it runs beside them)
        with fork_adv:
            fork_adv.side.wait_event(d_ready)
            g_adv = model.discriminator.adv_loss(d_in)   # bce(D(d_in), ones) with D frozen (tt:240-241)
            g_adv.backward()
        _zero_grads(optG)
        if red is not None:
            red.arm()
        fork_rec.join()
        fork_dec.join()
        early = alpha_sup * g_sup + beta_rec * g_rec + gamma_cov * cov_term + gamma_acf * acf_term
        early.backward()
        fork_adv.join()
        h_hat.backward(h_use.grad + h_adv.grad)
        with torch.no_grad():
            g_total = g_adv + alpha_sup * g_sup + beta_rec * g_rec + gamma_cov * cov_term + gamma_acf * acf_term
    else:
        g_adv = model.discriminator.adv_loss(d_in)   # bce(D(d_in), ones) with D frozen: one head kernel (tt:240-241)
        fork_rec.join()
        fork_dec.join()
        g_total = g_adv + alpha_sup * g_sup + beta_rec * g_rec + gamma_cov * cov_term + gamma_acf * acf_term
        _zero_grads(optG)
        if red is not None:
            red.arm()
        g_total.backward()
    params = _params(model.generator, model.supervisor, model.embedder, model.recovery)
    _reduce_and_step(optG, params, clip, red)
    nz.end()
    if schedulerG is not None:
        schedulerG.step()
    vals = (g_total, g_adv, g_sup, g_rec, cov_term, acf_term)
    return tuple(_as_float(v, sync) for v in vals)


class GraphedJointStep:
    """One joint training step (disc_step + gen_step, tt:379-395) captured in a CUDA graph.

    A joint step issues ~3200 kernel launches and ~30 ms of Python/ctypes work; once the GPU finishes a step
    faster than the host can issue it, the host is the bottleneck.  Capturing removes it: every replay re-runs
    the same launches with the same device pointers, while everything that changes from step to step lives in
    device memory -- the input batch (`self.x`), the instance-noise std (`self.std`), the Philox position
    (noise.DeviceNoise.ctr), Adam's step counter / lr / bias corrections (FusedAdam(capturable=True)).
    The first `warmup` calls run eagerly (they are real training steps); the next call captures and replays.
    Needs on-device noise (noise=None) and capturable optimisers; LR schedulers are stepped here, outside the graph."""

    def __init__(self, model, optD, optG, device, *, label_smooth, clip, r1_gamma, target_acc, band, alpha_sup,
                 beta_rec, gamma_cov, gamma_acf, acf_max_lag, schedulerD=None, schedulerG=None, warmup: int = 3,
                 noise=None):
        if not (getattr(optD, "capturable", False) and getattr(optG, "capturable", False)):
            raise ValueError("GraphedJointStep needs FusedAdam(capturable=True) optimisers")
        self.model, self.optD, self.optG, self.device = model, optD, optG, torch.device(device)
        self.kw = dict(label_smooth=label_smooth, clip=clip, r1_gamma=r1_gamma, target_acc=target_acc, band=band,
                       alpha_sup=alpha_sup, beta_rec=beta_rec, gamma_cov=gamma_cov, gamma_acf=gamma_acf,
                       acf_max_lag=acf_max_lag)
        self.schedulerD, self.schedulerG = schedulerD, schedulerG
        if isinstance(noise, HostReplayNoise):
            raise ValueError("host-replayed noise cannot be captured in a CUDA graph")
        self.noise = noise
        self.warmup, self.calls = int(warmup), 0
        self.x = None
        self.std = torch.zeros((), dtype=torch.float32, device=self.device)
        self.graph = None
        self.out = None
        self._noise_on = None

    def _joint(self):
        k = self.kw
        std = self.std if self._noise_on else 0.0
        if not (_CONCURRENT and _OVERLAP_STEPS):
            d = disc_step(self.model, self.x, self.device, self.optD, k["label_smooth"], std, k["clip"], None,
                          k["r1_gamma"], target_acc=k["target_acc"], band=k["band"], noise=self.noise, sync=False)
            g = gen_step(self.model, self.x, self.device, self.optG, k["alpha_sup"], k["beta_rec"], std, k["clip"],
                         None, k["gamma_cov"], k["gamma_acf"], k["acf_max_lag"], noise=self.noise, sync=False)
            return torch.stack([v.float().reshape(()) for v in tuple(d) + tuple(g)])
        # gen_step reads the discriminator only from its D forward pass on (tt:240); E, G, S, R are not touched by
        # disc_step's optimiser.  So gen_step runs on its own stream that depends on the START of the joint step and
        # waits for D's update right before it needs D: its E -> R / G -> S -> R forward passes (15 of the step's 54
        # layer passes) fill the SMs beside disc_step's R1 chain.  Issue order, noise positions and all-reduce call
        # sites are those of the sequential step; only the dependencies the CUDA graph records differ.
        main = torch.cuda.current_stream(self.device)
        crit, gstream = _side_stream(self.device, 5), _side_stream(self.device, 6)
        nz = _noise_source(self.noise, self.device)
        hold = hasattr(nz, "hold")
        if hold:
            nz.hold()
        start = torch.cuda.Event()
        start.record(main)
        crit.wait_event(start)
        with torch.cuda.stream(crit):         # disc_step's main line (G -> S -> D -> R1 chain) is the step's critical path
            d = disc_step(self.model, self.x, self.device, self.optD, k["label_smooth"], std, k["clip"], None,
                          k["r1_gamma"], target_acc=k["target_acc"], band=k["band"], noise=self.noise, sync=False)
            d_ready = torch.cuda.Event()
            d_ready.record(crit)
        gstream.wait_event(start)
        with torch.cuda.stream(gstream):
            g = gen_step(self.model, self.x, self.device, self.optG, k["alpha_sup"], k["beta_rec"], std, k["clip"],
                         None, k["gamma_cov"], k["gamma_acf"], k["acf_max_lag"], noise=self.noise, sync=False,
                         d_ready=d_ready, fork_base=2)
        main.wait_stream(crit)
        main.wait_stream(gstream)
        if hold:
            nz.release()
        return torch.stack([v.float().reshape(()) for v in tuple(d) + tuple(g)])

    def __call__(self, x, inst_noise_std: float):
        """x: (B,T,C) device or pinned-host tensor.  Returns the 8 logged scalars as one device tensor
        (loss_D, acc_D, loss_G, adv, sup, rec, cov, acf); it is overwritten by the next call."""
        if self.x is None:
            self.x = torch.empty(x.shape, dtype=torch.float32, device=self.device)
            self._noise_on = inst_noise_std > 0
        if tuple(x.shape) != tuple(self.x.shape):
            raise ValueError(f"GraphedJointStep was built for batches of shape {tuple(self.x.shape)}, got {tuple(x.shape)}")
        if (inst_noise_std > 0) != self._noise_on:
            raise ValueError("instance noise cannot be switched on/off after capture")
        self.x.copy_(x, non_blocking=True)
        self.std.fill_(float(inst_noise_std))
        self.optD.push_lr()
        self.optG.push_lr()
        if self.calls < self.warmup:
            out = self._joint()
        else:
            if self.graph is None:
                torch.cuda.synchronize(self.device)
                self.graph = torch.cuda.CUDAGraph()
                # thread_local: the autograd worker and (under DP) the NCCL watchdog issue CUDA calls from other
                # threads while this thread captures
                with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                    self.out = self._joint()
            self.graph.replay()
            out = self.out
        self.calls += 1
        for sch in (self.schedulerD, self.schedulerG):
            if sch is not None:
                sch.step()
        return out


# ---------------------- Full training (tt:281-422) ----------------------

def train_single_npz(npz_path: Path, out_dir: Path,
                     batch_size=64,
                     ae_epochs=120,
                     sup_epochs=150,
                     gan_steps=8000,
                     lr_g=1e-3, lr_d=2e-4,
                     betas=(0.5, 0.9),
                     alpha_sup=5.0,
                     beta_rec=0.2,
                     label_smooth=0.2,
                     inst_noise_start=0.3,
                     inst_noise_end=0.1,
                     grad_clip=0.5,
                     layers=1,
                     dropout=0.2,
                     seed=42,
                     r1_gamma=1.0,
                     d_min_acc=0.45,
                     d_max_acc=0.60,
                     gamma_cov=0.05,
                     gamma_acf=0.05,
                     acf_max_lag=64,
                     device=None,
                     *,
                     z_dim: Optional[int] = None,
                     hidden_dim: Optional[int] = None,
                     noise: Optional[str] = None,
                     proj_dtype: str = "fp32",
                     log_every: Optional[int] = None,
                     graph: Optional[bool] = None,
                     resident_data: bool = True,
                     resume: bool = False,
                     ckpt_every: int = 500,
                     stop_after: Optional[int] = None):
    """Same schedule, logs and artefacts as the reference.  Extras (keyword-only): `z_dim`/`hidden_dim`
    override adaptive_dims; `noise="host"` replays the reference's CPU random stream (parity runs), default is
    on-device Philox; `proj_dtype` "fp32" | "bf16" (bf16 input projections with a bf16 gi tensor, one TF32 pass
    for dX / weight gradients) | "tf32" (one TF32 pass everywhere).
    The default path is the fast path: every step still gets its CSV row and the reference's per-step
    best-checkpoint rule is still applied to every step (on the device, `BestSnapshot`), but the host only
    synchronises with the GPU every `log_every` steps (default 25; 1 with host-replayed noise) to write the rows
    and hand a changed best snapshot to the asynchronous writer; `graph` (default: on whenever the noise is
    on-device and, under data parallelism, the all-reduces are the peer-memory kernels) replays the joint step
    from a CUDA graph (GraphedJointStep; full-size batches only -- the ragged last batch of an epoch runs
    eagerly); `resident_data` keeps the dataset in HBM and gathers
    batches on the device in the reference's shuffle order (DeviceLoader); `resume=True` continues the GAN phase
    from `out_dir/ckpt_latest.pt` (weights, both optimisers, LR schedules, noise decay, step counter) instead of
    starting over -- the reference can only start over; `ckpt_every` is the reference's hard-coded 500 (tt:406);
    `stop_after=k` ends the run after GAN step k with a checkpoint written (pre-emption / tests)."""
    npz_path, out_dir = Path(npz_path), Path(out_dir)
    set_seeds(seed)
    device = device or device_autoselect()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("timegan_b200.train_single_npz: device must be CUDA; there is no CPU path")
    rank0 = _dist.rank() == 0
    out_dir.mkdir(parents=True, exist_ok=True)
    ops.set_proj_mode(proj_dtype)

    data = np.load(npz_path)
    X = data["X"].astype(np.float32)
    N, T, C = X.shape
    zd, hd = adaptive_dims(C, T)
    z_dim = int(z_dim) if z_dim is not None else zd
    h_dim = int(hidden_dim) if hidden_dim is not None else hd

    log_file = out_dir / "train_log.csv"
    ckpt_path, best_path = out_dir / "ckpt_latest.pt", out_dir / "ckpt_best.pt"
    resume_state = None
    if resume and ckpt_path.exists():
        resume_state = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    if rank0 and not (resume_state is not None and log_file.exists()):
        with open(log_file, "w", newline="") as f:
            csv.writer(f).writerow(["step", "phase", "loss_D", "acc_D", "loss_G", "loss_adv", "loss_sup", "loss_rec",
                                    "loss_cov", "loss_acf"])

    def LOG(msg):
        if rank0:
            print(msg, flush=True)

    LOG(f"==> {npz_path.name} | N={N} T={T} C={C}  z_dim={z_dim} h_dim={h_dim}  device={device}")

    loader = DeviceLoader(X, batch_size, device) if resident_data else make_loader(X, batch_size)
    model = TimeGAN(x_dim=C, z_dim=z_dim, hidden_dim=h_dim, num_layers=layers, dropout=dropout).to(device)
    nz = HostReplayNoise(device) if noise == "host" else None

    start_step = 0
    if resume_state is not None:
        # the checkpoint was written during the GAN phase: phases 1 and 2 are behind us
        model.load_state_dict(resume_state["model"])
        start_step = int(resume_state["step"])
        LOG(f"[resume] {ckpt_path.name}: continuing the GAN phase after step {start_step}/{gan_steps}")
    else:
        optER = FusedAdam(_params(model.embedder, model.recovery), lr=lr_g, betas=betas)
        phase_autoencoder(model, loader, device, optER, grad_clip, ae_epochs, LOG)

        optS = FusedAdam(model.supervisor.parameters(), lr=lr_g, betas=betas)
        phase_supervisor(model, loader, device, optS, grad_clip, sup_epochs, LOG)

    # CUDA-graph replay is the default whenever it is possible: on-device noise, and under data parallelism the
    # all-reduces must be this package's own peer-memory kernels (dist.PeerComm); NCCL collectives are issued eagerly
    can_graph = nz is None and (not _dist.is_enabled() or _dist.peer_comm() is not None)
    use_graph = can_graph if graph is None else (bool(graph) and can_graph)
    if log_every is None:
        log_every = 1 if nz is not None else 25
    optD = FusedAdam(model.discriminator.parameters(), lr=lr_d, betas=betas, capturable=use_graph)
    optG = FusedAdam(_params(model.generator, model.supervisor, model.embedder, model.recovery), lr=lr_g, betas=betas,
                     capturable=use_graph)
    if resume_state is not None:
        optG.load_state_dict(resume_state["optG"])
        optD.load_state_dict(resume_state["optD"])
        for opt in (optG, optD):       # Adam moments back onto the device; the base LR is re-derived below
            for st in opt.state.values():
                for k in ("exp_avg", "exp_avg_sq"):
                    if k in st:
                        st[k] = st[k].to(device)
            for g in opt.param_groups:
                g["lr"] = g.get("initial_lr", g["lr"])
    milestones = [gan_steps // 2, int(gan_steps * 0.75)]
    schedulerG = optim.lr_scheduler.MultiStepLR(optG, milestones=milestones, gamma=0.5)
    schedulerD = optim.lr_scheduler.MultiStepLR(optD, milestones=milestones, gamma=0.5)
    if start_step:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for _ in range(start_step):     # one scheduler tick per completed GAN step (tt:348-349)
                schedulerG.step()
                schedulerD.step()

    def lr_at(name, st):
        base = lr_g if name == "optG" else lr_d
        return base * (0.5 ** sum(1 for m in milestones if st >= m))

    # on-device noise: one Philox stream per call, keyed by (seed, rank, step the run starts at) -- a second call in
    # the same process (main()'s loop over files, another seed) starts its own stream instead of continuing the first
    step_noise = nz
    if nz is None:
        step_noise = device_noise(int(seed) * 1000003 + 7919 * _dist.rank() + 104729 * start_step, device)
    loader_iter = iter(loader)
    noise_decay = (inst_noise_start - inst_noise_end) / max(1, gan_steps)
    inst_noise = max(inst_noise_end, inst_noise_start - noise_decay * start_step)
    best_ckpt_loss = math.inf
    if resume_state is not None:        # the best loss seen before the interruption (kept in the checkpoint's meta)
        best_ckpt_loss = float(resume_state.get("meta", {}).get("best_loss", math.inf))
    saver, best_saver = AsyncCheckpointer(), AsyncCheckpointer()
    meta = {"npz": npz_path.name, "z_dim": z_dim, "h_dim": h_dim}
    target = 0.5 * (d_min_acc + d_max_acc)
    band = max(0.0, d_max_acc - d_min_acc)
    pending = []   # (step, device scalars) awaiting a flush
    skipped = []   # GAN steps without an optimiser update (data-parallel batches smaller than the world size)
    snap = BestSnapshot(model, optG, optD, device, best_ckpt_loss) if rank0 else None
    graphed = None
    if use_graph:
        graphed = GraphedJointStep(model, optD, optG, device, label_smooth=label_smooth, clip=grad_clip,
                                   r1_gamma=r1_gamma, target_acc=target, band=band, alpha_sup=alpha_sup,
                                   beta_rec=beta_rec, gamma_cov=gamma_cov, gamma_acf=gamma_acf, acf_max_lag=acf_max_lag,
                                   schedulerD=schedulerD, schedulerG=schedulerG, noise=step_noise)

    def flush(now_step):
        """ONE device->host transfer for every step logged since the last flush: CSV rows, the [GAN] progress lines,
        the best-checkpoint bookkeeping and the health checks of the data-parallel transport."""
        nonlocal best_ckpt_loss
        if not pending:
            return
        rows_dev = [torch.stack([v.float().reshape(()) for v in row]) for _, row in pending]
        tail = [snap.best.reshape(()), snap.best_step.reshape(())] if snap is not None else []
        flat = torch.cat([torch.stack(rows_dev).reshape(-1)] + [torch.stack(tail)] if tail else
                         [torch.stack(rows_dev).reshape(-1)]).cpu().tolist()
        vals = [flat[i * 8:(i + 1) * 8] for i in range(len(pending))]
        rows = []
        for (st, _), r in zip(pending, vals):
            rows.append([st, "GAN"] + r)
            if st % 100 == 0:
                LOG(f"[GAN] step {st}/{gan_steps}  D:loss={r[0]:.4f} acc≈{r[1]:.2f}  G:total={r[2]:.4f} "
                    f"(adv={r[3]:.4f}, sup={r[4]:.4f}, rec={r[5]:.4f}, cov={r[6]:.4f}, acf={r[7]:.4f})")
        if rank0:
            with open(log_file, "a", newline="") as f:
                csv.writer(f).writerows(rows)
        if _dist.peer_comm() is not None:
            _dist.peer_comm().check_status()        # a rank that timed out waiting for a peer poisons its sums: stop here
        if snap is not None:
            best_ckpt_loss, best_step = float(flat[-2]), int(flat[-1])
            if best_step >= 0 and best_step != snap.written_step:
                # tt:410-413: the weights after the step with the lowest g_total so far -- captured on the device at
                # that step, written now (behind the training stream)
                rewind = (now_step - best_step) - sum(1 for st in skipped if best_step < st <= now_step)
                best_saver.save_state(best_path, snap.checkpoint(best_step, rewind, meta, lr_at))
                snap.written_step = best_step
        pending.clear()

    for step in range(start_step + 1, gan_steps + 1):
        try:
            (x_batch,) = next(loader_iter)
        except StopIteration:
            loader_iter = iter(loader)
            (x_batch,) = next(loader_iter)
        x = _dist.shard_batch(x_batch.to(device, non_blocking=True))
        if x is None:
            # fewer sequences than ranks (tail of an epoch under data parallelism): every rank sees the same count and
            # skips the batch; the step still counts so that schedules and checkpoints stay aligned with the reference
            schedulerD.step(); schedulerG.step()
            inst_noise = max(inst_noise_end, inst_noise - noise_decay)
            skipped.append(step)
            continue

        if graphed is not None and x.shape[0] * _dist.world_size() == batch_size and inst_noise > 0:
            vals = graphed(x, inst_noise).clone()           # the graph's output buffer is reused by the next replay
            row = tuple(vals.unbind(0))
        else:
            d_loss, d_acc = disc_step(model, x, device, optD, label_smooth, inst_noise, grad_clip, schedulerD, r1_gamma,
                                      target_acc=target, band=band, noise=step_noise, sync=False)
            g_vals = gen_step(model, x, device, optG, alpha_sup, beta_rec, inst_noise, grad_clip, schedulerG, gamma_cov,
                              gamma_acf, acf_max_lag, noise=step_noise, sync=False)
            row = (d_loss, d_acc) + tuple(g_vals)
        pending.append((step, row))
        if snap is not None:
            snap.update(row[2], step)
        stopping = stop_after is not None and step >= stop_after
        ckpt_now = step % max(1, ckpt_every) == 0 or step == gan_steps or stopping
        if len(pending) >= max(1, log_every) or ckpt_now:
            flush(step)

        inst_noise = max(inst_noise_end, inst_noise - noise_decay)
        if ckpt_now and rank0:
            saver.save(ckpt_path, model, optG, optD, step, dict(meta, best_loss=best_ckpt_loss))   # behind the stream
        if stopping:
            saver.wait()
            best_saver.wait()
            LOG(f"[stop] stop_after={stop_after}: checkpoint at step {step}, resume with resume=True")
            return
    saver.wait()
    best_saver.wait()
    dump = os.environ.get("TIMEGAN_B200_DUMP_WEIGHT_SUM")
    if dump:    # DP self-check hook: every rank records an exact fingerprint of its final weights (replicas must agree)
        fp = [float(p.detach().double().sum().item()) for p in model.parameters()]
        Path(dump, f"weight_sum_rank{_dist.rank()}.txt").write_text(repr(fp))

    model.eval()
    from .generate_long_synth import generate_windows
    X_hat = generate_windows(model, N, T, z_dim, device, noise=nz)
    if rank0:
        np.savez_compressed(out_dir / "synthetic.npz", X=X_hat)
        print(f"Saved synthetic: {out_dir / 'synthetic.npz'}")
    return True


# ---------------------- Entry point (tt:427-495) ----------------------

def build_argparser():
    import argparse
    ap = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--data_dir", type=str, default="./preprocessed",
                    help="Folder with postureX_with_exo.npz / postureX_no_exo.npz")
    ap.add_argument("--out_dir", type=str, default="./timegan_runs", help="Output root folder")
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--ae_epochs", type=int, default=120)
    ap.add_argument("--sup_epochs", type=int, default=150)
    ap.add_argument("--gan_steps", type=int, default=8000)
    ap.add_argument("--lr_g", type=float, default=1e-3)
    ap.add_argument("--lr_d", type=float, default=2e-4)
    ap.add_argument("--beta1", type=float, default=0.5)
    ap.add_argument("--beta2", type=float, default=0.9)
    ap.add_argument("--alpha_sup", type=float, default=5.0)
    ap.add_argument("--beta_rec", type=float, default=0.2)
    ap.add_argument("--label_smooth", type=float, default=0.2)
    ap.add_argument("--inst_noise_start", type=float, default=0.3)
    ap.add_argument("--inst_noise_end", type=float, default=0.1)
    ap.add_argument("--grad_clip", type=float, default=0.5)
    ap.add_argument("--layers", type=int, default=1)
    ap.add_argument("--dropout", type=float, default=0.2)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--r1_gamma", type=float, default=1.0)
    ap.add_argument("--d_min_acc", type=float, default=0.45)
    ap.add_argument("--d_max_acc", type=float, default=0.60)
    ap.add_argument("--gamma_cov", type=float, default=0.05)
    ap.add_argument("--gamma_acf", type=float, default=0.05)
    ap.add_argument("--acf_max_lag", type=int, default=64)
    # extras of this implementation (defaults = reference behaviour)
    ap.add_argument("--z_dim", type=int, default=None, help="override adaptive_dims' latent size")
    ap.add_argument("--hidden_dim", type=int, default=None, help="override adaptive_dims' hidden size")
    ap.add_argument("--proj_dtype", type=str, default="fp32", choices=["fp32", "bf16", "tf32"])
    ap.add_argument("--noise", type=str, default=None, choices=[None, "host"])
    ap.add_argument("--log_every", type=int, default=None,
                    help="GAN steps between host synchronisations (CSV rows are per step either way); default 25")
    ap.add_argument("--graph", dest="graph", action="store_true", default=None,
                    help="replay the joint step from a CUDA graph (default: on whenever possible)")
    ap.add_argument("--no_graph", dest="graph", action="store_false", help="issue every step eagerly")
    ap.add_argument("--resume", action="store_true", help="continue the GAN phase from <out_dir>/ckpt_latest.pt")
    ap.add_argument("--ckpt_every", type=int, default=500, help="GAN steps between ckpt_latest.pt writes")
    return ap


def main(argv=None):
    args = build_argparser().parse_args(argv)
    _dist.init()                      # under torchrun: cuda:LOCAL_RANK becomes the current device, peer transport on
    device = device_autoselect()
    print(f"Using device: {device}")
    out_root = Path(args.out_dir)
    out_root.mkdir(parents=True, exist_ok=True)
    files = sorted(Path(args.data_dir).glob("posture*_*.npz"))
    if not files:
        raise SystemExit(f"No NPZs found in {args.data_dir}. Run preprocessing first.")
    for fp in files:
        train_single_npz(
            fp, out_root / fp.stem,
            batch_size=args.batch_size, ae_epochs=args.ae_epochs, sup_epochs=args.sup_epochs,
            gan_steps=args.gan_steps, lr_g=args.lr_g, lr_d=args.lr_d, betas=(args.beta1, args.beta2),
            alpha_sup=args.alpha_sup, beta_rec=args.beta_rec, label_smooth=args.label_smooth,
            inst_noise_start=args.inst_noise_start, inst_noise_end=args.inst_noise_end, grad_clip=args.grad_clip,
            layers=args.layers, dropout=args.dropout, seed=args.seed, r1_gamma=args.r1_gamma,
            d_min_acc=args.d_min_acc, d_max_acc=args.d_max_acc, gamma_cov=args.gamma_cov, gamma_acf=args.gamma_acf,
            acf_max_lag=args.acf_max_lag, device=device, z_dim=args.z_dim, hidden_dim=args.hidden_dim,
            proj_dtype=args.proj_dtype, noise=args.noise, log_every=args.log_every, graph=args.graph,
            resume=args.resume, ckpt_every=args.ckpt_every)
    _dist.shutdown()


if __name__ == "__main__":
    main()

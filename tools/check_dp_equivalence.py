#!/usr/bin/env python3
"""torchrun --nproc-per-node 2 tools/check_dp_equivalence.py
Two data-parallel ranks, each holding its shard of a global batch, must reproduce the single-process joint steps
on the whole batch: same losses (the batch statistics are all-reduced) and same updated weights --
including a RAGGED global batch (7 sequences -> shards of 4 and 3: the count-weighted statistics of dist.py, the
reference's drop_last=False tail of tt:33-37)."""
import copy, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as td


class ShardedReplayNoise:
    """Draws the GLOBAL-batch tensor from the seeded CPU generator (reference order) and keeps this rank's slice."""
    def __init__(self, device, world, rank):
        self.device, self.world, self.rank = device, world, rank
        self.n_global = None          # set before every step: sequences in the global batch
    def begin(self): pass
    def end(self): pass
    def _slice(self, t):
        from timegan_b200 import dist as D
        a, b = D.shard_bounds(self.n_global, self.world, self.rank)
        return t[a:b].contiguous().to(self.device)
    def rand(self, *shape):
        return self._slice(torch.rand(self.n_global, *shape[1:]))
    def randn_like(self, h, time_major=False):
        if time_major and h.dim() == 3:
            like = torch.empty(h.shape[1], self.n_global, h.shape[2]).transpose(0, 1)
        else:
            like = torch.empty(self.n_global, *h.shape[1:])
        return self._slice(torch.randn_like(like))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/null")
    import timegan_b200 as tg
    from timegan_b200 import dist as D, train_timegan as tt
    D.init(backend="nccl", device=dev)
    torch.manual_seed(0)
    base = tg.TimeGAN(14, 24, 24, 2, 0.0)
    Bg, T = 8, 96
    xs = [torch.rand(Bg, T, 14), torch.rand(7, T, 14), torch.rand(Bg, T, 14)]      # the middle batch is ragged
    hp = dict(label_smooth=0.2, std=0.3, clip=0.5, r1=1.0, target=0.525, band=0.15, a=5.0, b=0.2, gc=0.05, ga=0.05, lag=32)

    def run(model, noise, shard):
        P = tt._params
        oD = tg.FusedAdam(model.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9))
        oG = tg.FusedAdam(P(model.generator, model.supervisor, model.embedder, model.recovery), lr=1e-3, betas=(0.5, 0.9))
        out = []
        for i, xg in enumerate(xs):
            torch.manual_seed(100 + i)
            if hasattr(noise, "n_global"):
                noise.n_global = xg.shape[0]
            x = (D.shard_batch(xg) if shard else xg).to(dev)
            d = tt.disc_step(model, x, dev, oD, hp["label_smooth"], hp["std"], hp["clip"], None, hp["r1"],
                             target_acc=hp["target"], band=hp["band"], noise=noise)
            g = tt.gen_step(model, x, dev, oG, hp["a"], hp["b"], hp["std"], hp["clip"], None, hp["gc"], hp["ga"], hp["lag"],
                            noise=noise)
            out.append(list(d) + list(g))
        return torch.tensor(out, dtype=torch.float64)

    m_dp = copy.deepcopy(base).to(dev)
    res_dp = run(m_dp, ShardedReplayNoise(dev, world, rank), shard=True)
    D.disable()                                   # single-process reference on the whole batch (every rank, no comms)
    m_sp = copy.deepcopy(base).to(dev)
    res_sp = run(m_sp, tt.HostReplayNoise(dev), shard=False)
    rel = ((res_dp - res_sp).abs() / res_sp.abs().clamp_min(1e-3)).max().item()
    wdiff = max((a.detach() - b.detach()).abs().max().item() for a, b in zip(m_dp.parameters(), m_sp.parameters()))
    ok = rel < 2e-4 and wdiff < 2e-5
    print(f"rank {rank}: max rel loss deviation {rel:.2e}, max weight deviation {wdiff:.2e} -> {'OK' if ok else 'MISMATCH'}", flush=True)
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

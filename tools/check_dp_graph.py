#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_dp_graph.py
Data-parallel joint steps replayed from the CUDA graph (peer-memory all-reduce kernels inside the graph) must equal
the same DP steps issued eagerly: same weights, same per-rank device noise stream, 8 steps incl. an LR change."""
import copy, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as td


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/null")
    import timegan_b200 as tg
    from timegan_b200 import dist as D, train_timegan as tt
    D.init(backend="nccl", device=dev)
    assert D.peer_comm() is not None
    torch.manual_seed(5)
    base = tg.TimeGAN(14, 24, 24, 2, 0.0).to(dev)
    g = torch.Generator().manual_seed(100)
    xs = [D.shard_batch(torch.rand(6 * world, 48, 14, generator=g)).to(dev) for _ in range(8)]
    hp = dict(label_smooth=0.2, clip=0.5, r1_gamma=1.0, target_acc=0.525, band=0.15, alpha_sup=5.0, beta_rec=0.2,
              gamma_cov=0.05, gamma_acf=0.05, acf_max_lag=16)
    P = tt._params
    outs, finals = {}, {}
    for mode in ("eager", "graph"):
        m = copy.deepcopy(base)
        cap = mode == "graph"
        oD = tg.FusedAdam(m.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9), capturable=cap)
        oG = tg.FusedAdam(P(m.generator, m.supervisor, m.embedder, m.recovery), lr=1e-3, betas=(0.5, 0.9), capturable=cap)
        sD = torch.optim.lr_scheduler.MultiStepLR(oD, milestones=[4, 6], gamma=0.5)
        sG = torch.optim.lr_scheduler.MultiStepLR(oG, milestones=[4, 6], gamma=0.5)
        nz = tt.device_noise(1234 + rank, dev)
        rows = []
        if cap:
            step = tt.GraphedJointStep(m, oD, oG, dev, schedulerD=sD, schedulerG=sG, warmup=2, noise=nz, **hp)
            for i, x in enumerate(xs):
                rows.append(step(x, 0.3 - 0.01 * i).clone())
            assert step.graph is not None
        else:
            for i, x in enumerate(xs):
                d = tt.disc_step(m, x, dev, oD, hp["label_smooth"], 0.3 - 0.01 * i, hp["clip"], sD, hp["r1_gamma"],
                                 target_acc=hp["target_acc"], band=hp["band"], noise=nz, sync=False)
                gq = tt.gen_step(m, x, dev, oG, hp["alpha_sup"], hp["beta_rec"], 0.3 - 0.01 * i, hp["clip"], sG,
                                 hp["gamma_cov"], hp["gamma_acf"], hp["acf_max_lag"], noise=nz, sync=False)
                rows.append(torch.stack([v.float().reshape(()) for v in d + gq]))
        torch.cuda.synchronize()
        outs[mode] = torch.stack(rows).cpu()
        finals[mode] = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    D.peer_comm().check_status()
    dl = (outs["graph"] - outs["eager"]).abs().max().item()
    dw = max((finals["graph"][k] - finals["eager"][k]).abs().max().item() for k in finals["eager"])
    # every rank must hold the same weights after DP steps (bit-identical sums + identical Adam)
    flat = torch.cat([v.reshape(-1).float() for v in finals["graph"].values()]).to(dev)
    lo, hi = flat.clone(), flat.clone()
    td.all_reduce(lo, op=td.ReduceOp.MIN); td.all_reduce(hi, op=td.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    ok = torch.isfinite(outs["graph"]).all().item() and torch.allclose(outs["graph"], outs["eager"], rtol=2e-4, atol=1e-6) \
        and dw < 2e-5 and same
    print(f"rank {rank}: graph vs eager DP: max loss diff {dl:.2e}, max weight diff {dw:.2e}, "
          f"weights identical across ranks: {same} -> {'OK' if ok else 'MISMATCH'}", flush=True)
    D.shutdown()
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

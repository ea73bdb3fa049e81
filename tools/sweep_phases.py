#!/usr/bin/env python3
"""BASELINE config c4: hidden sweep over the three scheduled phases (AE step, SUP step, joint step) on one GPU
(launch under torchrun for N > 1: each rank then holds B sequences and gradients are all-reduced).
Prints one JSON line per hidden size:  python tools/sweep_phases.py [--hidden 24 64 128] [--batch 256] [--steps 8]"""
import argparse, json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hidden", type=int, nargs="+", default=[24, 64, 128])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--proj", type=str, default="fp32")
    ap.add_argument("--eager", action="store_true", help="issue the joint step eagerly instead of replaying the CUDA graph")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/null")
    import timegan_b200 as tg
    from timegan_b200 import dist as D, ops, train_timegan as tt
    if world > 1:
        D.init(backend="nccl", device=dev)
    ops.set_proj_mode(a.proj)
    B, T = a.batch, 768
    xs = [torch.rand(B, T, 14, device=dev) for _ in range(4)]        # per-rank batches for the step functions
    # the scheduled phases shard the batch they are given (dist.shard_batch), like train_single_npz: give them the
    # GLOBAL batch (B * world sequences, identical on every rank) so that every GPU still works on B sequences
    gq = torch.Generator(device=dev).manual_seed(7)
    xg = [torch.rand(B * world, T, 14, device=dev, generator=gq) for _ in range(2)] if world > 1 else xs

    def timed(fn, data=None):
        data = xs if data is None else data
        for i in range(3):
            fn(data[i % len(data)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            fn(data[i % len(data)])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    for H in a.hidden:
        torch.manual_seed(42)
        m = tg.TimeGAN(14, H, H, a.layers, 0.0).to(dev)
        P = tt._params
        oER = tg.FusedAdam(P(m.embedder, m.recovery), lr=1e-3, betas=(0.5, 0.9))
        oS = tg.FusedAdam(m.supervisor.parameters(), lr=1e-3, betas=(0.5, 0.9))
        oD = tg.FusedAdam(m.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9))
        oG = tg.FusedAdam(P(m.generator, m.supervisor, m.embedder, m.recovery), lr=1e-3, betas=(0.5, 0.9))
        log = lambda s: None
        ae = timed(lambda x: tt.phase_autoencoder(m, [(x,)], dev, oER, 0.5, 1, log), xg)
        sup = timed(lambda x: tt.phase_supervisor(m, [(x,)], dev, oS, 0.5, 1, log), xg)

        # the joint phase as train_single_npz runs it by default: replayed from the CUDA graph (capturable optimisers)
        oDg = tg.FusedAdam(m.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9), capturable=True)
        oGg = tg.FusedAdam(P(m.generator, m.supervisor, m.embedder, m.recovery), lr=1e-3, betas=(0.5, 0.9), capturable=True)
        use_graph = not a.eager and (world == 1 or D.peer_comm() is not None)
        if use_graph:
            gj = tt.GraphedJointStep(m, oDg, oGg, dev, label_smooth=0.2, clip=0.5, r1_gamma=1.0, target_acc=0.525, band=0.15,
                                     alpha_sup=5.0, beta_rec=0.2, gamma_cov=0.05, gamma_acf=0.05, acf_max_lag=64, warmup=2)
            joint = lambda x: gj(x, 0.3)
        else:
            def joint(x):
                tt.disc_step(m, x, dev, oD, 0.2, 0.3, 0.5, None, 1.0, target_acc=0.525, band=0.15, sync=False)
                tt.gen_step(m, x, dev, oG, 5.0, 0.2, 0.3, 0.5, None, 0.05, 0.05, 64, sync=False)
        jt = timed(joint)
        del oDg, oGg
        if rank == 0:
            g = B * world
            print(json.dumps({"hidden": H, "layers": a.layers, "batch_per_gpu": B, "n_gpus": world, "proj": a.proj,
                              "issue": "AE/SUP eager, joint " + ("cuda-graph replay" if use_graph else "eager"),
                              "ae_seq_s": round(g / ae * 1e3, 1), "sup_seq_s": round(g / sup * 1e3, 1),
                              "joint_seq_s": round(g / jt * 1e3, 1), "ae_ms": round(ae, 2), "sup_ms": round(sup, 2),
                              "joint_ms": round(jt, 2)}), flush=True)
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


if __name__ == "__main__":
    main()

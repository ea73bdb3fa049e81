#!/usr/bin/env python3
"""Kernel timeline of ONE replayed joint step (CUDA graph), from CUPTI activity records (torch.profiler):

    python tools/trace_overlap.py [--hidden 64] [--out gpurun_out/timeline.csv]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/trace_overlap.py ...

Evidence for (a) the side-stream structure of the step -- which recurrent passes run beside each other, incl. gen_step's
forward passes beside disc_step's R1 chain -- and (b) under data parallelism, the gradient-bucket all-reduces
(peer_allreduce_kernel, csrc/peer_allreduce.cu) running WHILE BPTT kernels of the networks that are still going are on the
SMs.  Rank 0 writes every kernel of the traced replay (name, stream, start, duration) as CSV and prints a
summary: per stream busy time, concurrency histogram, and for every all-reduce launch the kernels that overlap it.
Timings taken under the profiler are NOT benchmark numbers; only order and overlap are read from them."""
import argparse
import os
import sys
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as td


def short(name: str) -> str:
    for key in ("gru_fwd_kernel", "gru_bwd_kernel", "gru_bwd_pair_kernel", "gru_jvp_fwd_kernel", "gru_jvp_bwd_kernel",
                "gru_cl_fwd_kernel", "gru_cl_bwd_kernel", "gru_cl_jvp_bwd_kernel", "tc_gemm_tn_kernel", "tc_gemm_bf16_kernel",
                "tc_wgrad_layer_kernel", "tc_wgrad_kernel", "peer_allreduce_kernel", "sgemm_kernel", "adam_multi",
                "sumsq_multi", "bigh_"):
        if key in name:
            return key
    name = name.replace("(anonymous namespace)::", "").replace("void ", "")
    return name.split("(")[0].split("<")[0][-40:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--proj", default="fp32")
    ap.add_argument("--out", default="gpurun_out/timeline.csv")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/null")
    import timegan_b200 as tg
    from timegan_b200 import dist as D, ops, train_timegan as tt
    if world > 1:
        D.init(backend="nccl", device=dev)
    ops.set_proj_mode(a.proj)
    torch.manual_seed(43)
    H = a.hidden
    m = tg.TimeGAN(14, H, H, 3, 0.0).to(dev)
    oD = tg.FusedAdam(m.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9), capturable=True)
    oG = tg.FusedAdam(tt._params(m.generator, m.supervisor, m.embedder, m.recovery), lr=1e-3, betas=(0.5, 0.9),
                      capturable=True)
    step = tt.GraphedJointStep(m, oD, oG, dev, label_smooth=0.2, clip=0.5, r1_gamma=1.0, target_acc=0.525, band=0.15,
                               alpha_sup=5.0, beta_rec=0.2, gamma_cov=0.05, gamma_acf=0.05, acf_max_lag=64, warmup=2)
    g = torch.Generator().manual_seed(7 + rank)
    xs = [torch.rand(a.batch, 768, 14, generator=g).to(dev) for _ in range(2)]
    for i in range(5):
        step(xs[i % 2], 0.3)
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    from torch.profiler import profile, ProfilerActivity
    # two replays under the profiler: the first absorbs CUPTI's start-up (the ranks enter it at different times, so its
    # all-reduces wait for the slowest rank); the SECOND one, entered after a barrier, is the one that is analysed
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(xs[0], 0.3)
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        step(xs[1], 0.3)
        torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    ok = True
    if rank == 0:
        import json, tempfile
        tmp = Path(tempfile.mkdtemp()) / "trace.json"
        prof.export_chrome_trace(str(tmp))
        tr = json.loads(tmp.read_text())
        ks = sorted(((float(e["ts"]), float(e["ts"]) + float(e["dur"]), int(e.get("args", {}).get("stream", -1)), e["name"])
                     for e in tr.get("traceEvents", []) if e.get("cat") == "kernel" and e.get("dur", 0) > 0),
                    key=lambda t: t[0])
        ks = [k for k in ks if "ncclDevKernel" not in k[3]]      # the barrier between the two replays
        ks = ks[len(ks) // 2:]                                   # same graph twice -> the second half is the second replay
        if not ks:
            print("no kernel records (CUPTI unavailable?)")
            ok = False
        else:
            t0 = ks[0][0]
            out = Path(a.out)
            out.parent.mkdir(parents=True, exist_ok=True)
            with open(out, "w") as f:
                f.write("start_us,dur_us,stream,kernel\n")
                for s, e, st, n in ks:
                    f.write(f"{s - t0:.1f},{e - s:.1f},{st},{short(n)}\n")
            span = ks[-1][1] - t0
            print(f"[timeline] world={world} hidden={H} proj={a.proj}: {len(ks)} kernels, span {span / 1e3:.2f} ms "
                  f"(under the profiler), written to {out}")
            # concurrency histogram over the recurrent kernels (sweep line)
            rec = [(s, e) for s, e, st, n in ks if "gru_" in n or "bigh_" in n]
            pts = sorted([(s, 1) for s, e in rec] + [(e, -1) for s, e in rec])
            hist, lvl, last = defaultdict(float), 0, t0
            for t, d in pts:
                hist[lvl] += t - last
                lvl, last = lvl + d, t
            tot = sum(hist.values()) or 1.0
            print("[timeline] recurrent kernels in flight -> share of the step: " +
                  ", ".join(f"{k}: {100 * v / span:.1f}%" for k, v in sorted(hist.items())))
            busy = defaultdict(float)
            for s, e, st, n in ks:
                busy[st] += e - s
            print("[timeline] busy time per stream (ms): " + ", ".join(f"s{k}: {v / 1e3:.2f}" for k, v in sorted(busy.items(), key=lambda kv: -kv[1])))
            ars = [(s, e, st) for s, e, st, n in ks if "peer_allreduce" in n]
            if ars:
                n_overl, tot_ar, tot_ov = 0, 0.0, 0.0
                for s, e, st in ars:
                    others = [(max(s, s2), min(e, e2), short(n2)) for s2, e2, st2, n2 in ks
                              if st2 != st and e2 > s and s2 < e and "peer_allreduce" not in n2]
                    ov = sum(max(0.0, b - a_) for a_, b, _ in others)
                    names = sorted({n for _, _, n in others})
                    tot_ar += e - s
                    tot_ov += min(ov, e - s)
                    n_overl += bool(others)
                    if e - s > 15.0:
                        print(f"[timeline] all-reduce at {(s - t0) / 1e3:7.3f} ms, {e - s:6.1f} us on s{st}: beside {names if names else 'nothing'}")
                print(f"[timeline] {len(ars)} all-reduce launches, {n_overl} with compute kernels of other streams in flight; "
                      f"{100 * tot_ov / max(tot_ar, 1e-9):.0f}% of all-reduce time covered")
    if world > 1:
        D.shutdown()
        td.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

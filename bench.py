#!/usr/bin/env python3
"""bench.py -- TimeGAN joint-training throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hidden H] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" is ONE joint training step (disc_step + gen_step, train_timegan.py:379-395) over one batch of
synthetic (B,768,14) windows.  Workload at every N: BASELINE config c2 per GPU -- z = hidden = 64, 3-layer GRU
stacks, batch 256 per GPU, fp32 (weak scaling: global batch = 256 N, gradients all-reduced over NCCL).

One JSON line on rank 0:
  value        sequences/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e          the same step driven from HOST buffers: pinned-host batch -> H2D copy -> disc_step/gen_step ->
               D2H read of the 8 logged loss scalars, all inside the timed region
  roofline     dominant kernel family of the step (device time share), its algorithmic bytes / launch duration
               against the measured HBM copy bandwidth; `ffma` gives the same family against the fp32 FMA peak
  cpu_baseline oracle/timegan_ref.py (the CPU restatement of the reference; kind "port") on this box's host cores:
               2 timed FULL-LENGTH joint steps of the same workload (B x 768 x 14; no truncation, no rescaling)
  also.c3      BASELINE config c3 (h = 128, reduced-precision projections) measured the same way in the same run
  dp_check     (N > 1) every replica holds bit-identical weights after all steps of the run
  --impl reference : only the CPU leg, K timed + W warm-up full-length steps (rank 0 only under torchrun).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

T_LEN, X_DIM = 768, 14
HP = dict(label_smooth=0.2, inst_noise=0.3, clip=0.5, r1_gamma=1.0, target=0.525, band=0.15, alpha_sup=5.0,
          beta_rec=0.2, gamma_cov=0.05, gamma_acf=0.05, acf_max_lag=64, lr_g=1e-3, lr_d=2e-4, betas=(0.5, 0.9))
METRIC = "TimeGAN train seq/sec (T=768,C=14)"


# stdout carries exactly ONE line (the JSON); anything a library prints on fd 1 (e.g. NCCL's version banner under
# torchrun) is sent to stderr: fd 1 is pointed at fd 2 for the whole run and the line is written to the saved fd
_REAL_STDOUT = None


def _guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="sequences per GPU")
    ap.add_argument("--proj", type=str, default="fp32", choices=["fp32", "bf16", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also-c3", dest="also_c3", action="store_false",
                    help="skip the secondary BASELINE config c3 (h=128, reduced-precision projections) line under `also`")
    ap.add_argument("--serial", action="store_true", help="disable side-stream concurrency inside the step")
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--dp-graph", action="store_true", help="experimental: capture the NCCL all-reduces too (N > 1)")
    ap.add_argument("--wgrad-ctas", type=int, default=-1,
                    help="SMs given to the side-stream weight-gradient kernels (0 = in line; -1 = package default)")
    ap.add_argument("--bt", type=int, default=0, choices=[0, 1, 2, 4],
                    help="sequences per CTA in the recurrent kernels (0 = the library's heuristic)")
    return ap.parse_args()


PROJ_LABEL = {"fp32": "fp32 (3xTF32) projections",
              "bf16": "bf16 input projections (bf16 operands on tcgen05 kind::f16, bf16 gi read by the recurrence; dX and "
                      "weight gradients one TF32 pass)",
              "tf32": "reduced-precision projections (one TF32 pass over fp32 operands)"}


def workload_name(a):
    cfg = "c2" if (a.hidden, a.layers, a.batch) == (64, 3, 256) else "custom"
    return (f"{cfg}: TimeGAN joint step (disc_step+gen_step) z=h={a.hidden} L={a.layers} B={a.batch}/GPU "
            f"T={T_LEN} C={X_DIM} {PROJ_LABEL[a.proj]}")


# ------------------------------------------------------------------------------------------------
# CPU leg: oracle port of the reference on the host cores -- FULL-LENGTH (T = 768) joint steps, no extrapolation
# ------------------------------------------------------------------------------------------------
class CpuArm:
    """The reference's joint step (tt:379-395 -> disc_step + gen_step) as restated by oracle/timegan_ref.py, on the
    host cores.  One model / optimiser pair lives for the whole arm, like a training run."""

    def __init__(self, a, threads):
        import torch
        from oracle import timegan_ref as R
        self.R, self.a = R, a
        torch.set_num_threads(threads)
        torch.manual_seed(0)
        self.model = R.build_model(X_DIM, a.hidden, a.hidden, a.layers, 0.0)
        self.opts = R.make_optimizers(self.model, HP["lr_g"], HP["lr_d"], HP["betas"])
        self.nz = R.TorchNoise()
        self.threads = threads

    def step(self, x):
        R = self.R
        t0 = time.perf_counter()
        R.d_step(self.model, x, self.opts["D"], self.nz, HP["label_smooth"], HP["inst_noise"], HP["clip"],
                 HP["r1_gamma"], HP["target"], HP["band"])
        R.g_step(self.model, x, self.opts["G"], self.nz, HP["alpha_sup"], HP["beta_rec"], HP["inst_noise"], HP["clip"],
                 HP["gamma_cov"], HP["gamma_acf"], HP["acf_max_lag"])
        return time.perf_counter() - t0


def pick_cpu_threads(a):
    """Fastest thread count of {1, 8, all} on this host, from one short (T = 96) probe step each -- the probes only
    choose the thread count, they are never part of a reported number."""
    import torch
    cores = os.cpu_count() or 1
    probe = {}
    for th in sorted({1, min(cores, 8), cores}):
        arm = CpuArm(a, th)
        x = torch.rand(a.batch, 96, X_DIM)
        arm.step(x)
        probe[th] = round(arm.step(x), 3)
    return min(probe, key=probe.get), probe, cores


def cpu_full_steps(a, steps, warmup, threads, short_warmup=False):
    """`warmup` untimed + `steps` timed joint steps on (B, 768, 14) batches.  Returns the list of step times (s).
    short_warmup: the untimed steps run on the first 96 timesteps only (they exist to spin up the thread pool and the
    allocator; used by the in-line cpu_baseline of the GPU arm, never by the reference arm)."""
    import torch
    arm = CpuArm(a, threads)
    g = torch.Generator().manual_seed(1234)
    xs = [torch.rand(a.batch, T_LEN, X_DIM, generator=g) for _ in range(2)]
    times = []
    for i in range(warmup + steps):
        x = xs[i % 2]
        if i < warmup and short_warmup:
            x = x[:, :96].contiguous()
        dt = arm.step(x)
        if i >= warmup:
            times.append(dt)
    return times


def cpu_baseline(a):
    import torch
    th, probe, cores = pick_cpu_threads(a)
    times = cpu_full_steps(a, 2, 1, th, short_warmup=True)
    per_step = sum(times) / len(times)
    return {"value": round(a.batch / per_step, 4), "unit": "seq/s", "cores": th, "kind": "port", "host_cores": cores,
            "sample": f"2 timed (after 1 short warm-up step) full-length joint steps of oracle/timegan_ref.py (torch "
                      f"{torch.__version__}, CPU) on {a.batch} x {T_LEN} x {X_DIM} batches -- the bench workload, no "
                      f"rescaling; thread count chosen from T=96 probe steps {probe} (s/step)",
            "s_per_step": round(per_step, 3)}


def bench_config(a, world):
    """`config` of the JSON line -- identical in both arms (the driver compares them)."""
    return {"workload": workload_name(a), "global_batch": a.batch * world, "parallelism": f"dp{world}",
            "l2": "8 rotating input batches; each step streams >2 GB of activations (>> 126 MB L2)"}


def run_reference(a):
    """--impl reference: K timed + W warm-up FULL joint steps (B x 768 x 14) of the CPU restatement of the reference.
    Under torchrun only rank 0 works.  At N > 1 the arm's config is the same weak-scaling workload (global batch
    B*N); the host processes it one rank-share (B sequences, full T) per step -- CPU seq/s does not depend on how many
    shares follow, and a 2048-sequence step would take 25 x 90 s -- and says so in `sample`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    th, probe, cores = pick_cpu_threads(a)
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    times = cpu_full_steps(a, steps, warmup, th)
    per_step = sum(times) / len(times)
    v = a.batch / per_step
    share = "" if world == 1 else (f"; each step is one rank's share ({a.batch} of the {a.batch * world} sequences of "
                                   f"the global batch), seq/s is per host and independent of the number of shares")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "seq/s", "n_gpus": a.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(per_step * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": bench_config(a, world),
        "cpu_baseline": {"value": round(v, 4), "unit": "seq/s", "cores": th, "kind": "port", "host_cores": cores,
                         "sample": f"every step = 1 full-length joint step of oracle/timegan_ref.py on a {a.batch} x "
                                   f"{T_LEN} x {X_DIM} batch (no time truncation, no rescaling){share}; thread count "
                                   f"chosen from T=96 probe steps {probe} (s/step)"},
        "e2e": {"value": round(v, 4), "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: sample DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
def secondary_config(a, dev, world, rank, hidden, proj, label, steps=6, warmup=3):
    """Device-resident joint-step throughput of ANOTHER BASELINE.json configuration inside the same run (same
    timing rules: warm-up, barrier + synchronize on both sides, CUDA events, max over ranks).  Reported under
    `also` -- the headline `value` stays the c2 workload at every N so that the driver's scaling efficiency
    compares like with like."""
    import torch
    import torch.distributed as td
    import timegan_b200 as tg
    from timegan_b200 import ops, dist as tdist, train_timegan as tt
    old = ops.get_proj_mode()
    ops.set_proj_mode(proj)
    try:
        torch.manual_seed(43)
        model = tg.TimeGAN(X_DIM, hidden, hidden, a.layers, 0.0).to(dev)
        use_graph = (not a.no_graph) and (world == 1 or tdist.peer_comm() is not None)
        optD = tg.FusedAdam(model.discriminator.parameters(), lr=HP["lr_d"], betas=HP["betas"], capturable=use_graph)
        optG = tg.FusedAdam(tt._params(model.generator, model.supervisor, model.embedder, model.recovery),
                            lr=HP["lr_g"], betas=HP["betas"], capturable=use_graph)
        g = torch.Generator().manual_seed(4321 + rank)
        xs = [torch.rand(a.batch, T_LEN, X_DIM, generator=g).to(dev) for _ in range(4)]
        if use_graph:
            step = tt.GraphedJointStep(model, optD, optG, dev, label_smooth=HP["label_smooth"], clip=HP["clip"],
                                       r1_gamma=HP["r1_gamma"], target_acc=HP["target"], band=HP["band"],
                                       alpha_sup=HP["alpha_sup"], beta_rec=HP["beta_rec"], gamma_cov=HP["gamma_cov"],
                                       gamma_acf=HP["gamma_acf"], acf_max_lag=HP["acf_max_lag"], warmup=2)
            joint = lambda x: step(x, HP["inst_noise"])
        else:
            def joint(x):
                d = tt.disc_step(model, x, dev, optD, HP["label_smooth"], HP["inst_noise"], HP["clip"], None,
                                 HP["r1_gamma"], target_acc=HP["target"], band=HP["band"], sync=False)
                q = tt.gen_step(model, x, dev, optG, HP["alpha_sup"], HP["beta_rec"], HP["inst_noise"], HP["clip"],
                                None, HP["gamma_cov"], HP["gamma_acf"], HP["acf_max_lag"], sync=False)
                return torch.stack([v.float().reshape(()) for v in d + q])

        def barrier():
            if world > 1:
                td.barrier()
            torch.cuda.synchronize()
        for i in range(warmup + (1 if use_graph else 0)):
            joint(xs[i % 4])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            out = joint(xs[i % 4])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            ms = float(t.item())
        ok = bool(torch.isfinite(out).all().item())
        del model, optD, optG, xs
        torch.cuda.empty_cache()
        return {"workload": f"{label}: TimeGAN joint step z=h={hidden} L={a.layers} B={a.batch}/GPU T={T_LEN} C={X_DIM} "
                            f"{PROJ_LABEL[proj]}", "global_batch": a.batch * world, "value": round(a.batch * world * steps / (ms * 1e-3), 2),
                "unit": "seq/s", "ms_per_step": round(ms / steps, 3), "steps": steps, "warmup": warmup,
                "issue": "cuda-graph replay" if use_graph else "eager", "finite": ok}
    finally:
        ops.set_proj_mode(old)


def run_ours(a):
    import torch
    import torch.distributed as td
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the TimeGAN hot path only exists as sm_100a kernels")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import timegan_b200 as tg
    from timegan_b200 import _lib, ops, dist as tdist, train_timegan as tt
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/null")   # keep NCCL's version banner off stdout (one JSON line)
        tdist.init(backend="nccl", device=dev)
    ops.set_proj_mode(a.proj)
    ops.set_bt_override(a.bt)
    if a.wgrad_ctas >= 0:
        ops.set_wgrad_overlap(a.wgrad_ctas)
    tt.set_concurrency(not a.serial)

    torch.manual_seed(42)
    model = tg.TimeGAN(X_DIM, a.hidden, a.hidden, a.layers, 0.0).to(dev)
    P = tt._params
    # CUDA-graph replay everywhere: under data parallelism the gradient / statistics all-reduces are this
    # package's own peer-memory kernels (csrc/peer_allreduce.cu), which are ordinary launches and replay with the
    # rest of the step.  TIMEGAN_B200_COMM=nccl falls back to eagerly issued NCCL collectives.
    use_graph = (not a.no_graph) and (world == 1 or tdist.peer_comm() is not None or a.dp_graph)
    optD = tg.FusedAdam(model.discriminator.parameters(), lr=HP["lr_d"], betas=HP["betas"], capturable=use_graph)
    optG = tg.FusedAdam(P(model.generator, model.supervisor, model.embedder, model.recovery), lr=HP["lr_g"],
                        betas=HP["betas"], capturable=use_graph)
    B = a.batch
    n_batches = 8
    g = torch.Generator().manual_seed(1234 + rank)
    host = [torch.rand(B, T_LEN, X_DIM, generator=g).pin_memory() for _ in range(n_batches)]
    resident = [h.to(dev) for h in host]

    def joint_eager(x):
        d = tt.disc_step(model, x, dev, optD, HP["label_smooth"], HP["inst_noise"], HP["clip"], None, HP["r1_gamma"],
                         target_acc=HP["target"], band=HP["band"], sync=False)
        gq = tt.gen_step(model, x, dev, optG, HP["alpha_sup"], HP["beta_rec"], HP["inst_noise"], HP["clip"], None,
                         HP["gamma_cov"], HP["gamma_acf"], HP["acf_max_lag"], sync=False)
        return torch.stack([v.float().reshape(()) for v in d + gq])

    graphed = None
    if use_graph:
        graphed = tt.GraphedJointStep(model, optD, optG, dev, label_smooth=HP["label_smooth"], clip=HP["clip"],
                                      r1_gamma=HP["r1_gamma"], target_acc=HP["target"], band=HP["band"],
                                      alpha_sup=HP["alpha_sup"], beta_rec=HP["beta_rec"], gamma_cov=HP["gamma_cov"],
                                      gamma_acf=HP["gamma_acf"], acf_max_lag=HP["acf_max_lag"], warmup=2)

    def joint(x):
        # the public API a user calls: the graphed step (train_single_npz(graph=True)) or the eager step functions
        return graphed(x, HP["inst_noise"]) if graphed is not None else joint_eager(x)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(a.warmup, 3) + (1 if use_graph else 0)):
        joint(resident[i % n_batches])
    barrier()

    # ---- timed region 1: inputs resident in HBM ----
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    _lib.prof_reset()
    _lib.prof_enable(not use_graph)     # per-call event pairs cannot be recorded into a graph; see region 3
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(a.steps):
        out = joint(resident[i % n_batches])
    host_issue_ms = (time.perf_counter() - t_host0) * 1e3   # CPU time to ISSUE the steps (no sync inside)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - l0
    _lib.prof_enable(False)
    prof = _lib.prof_read()
    clk = clocks.stop() if rank == 0 else None
    launches_per_step = launches / a.steps

    # ---- timed region 2: end to end from host buffers ----
    scal = torch.empty(8, dtype=torch.float32).pin_memory()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(a.steps):
        if graphed is not None:
            out = joint(host[i % n_batches])      # pinned host batch -> H2D copy into the graph's input buffer
        else:
            out = joint(host[i % n_batches].to(dev, non_blocking=True))
        scal.copy_(out, non_blocking=False)
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    assert all(v == v for v in scal.tolist()), f"non-finite losses in bench: {scal.tolist()}"

    # ---- region 3 (graph mode only): the same steps issued eagerly with per-call CUDA-event pairs, to attribute
    #      device time to kernel families (roofline bookkeeping) and to count launches per step ----
    if use_graph:
        n_prof = min(a.steps, 5)
        tt.set_concurrency(False)          # serialise the stacks so an event pair brackets ONE kernel family
        joint_eager(resident[0])
        _lib.prof_reset()
        _lib.prof_enable(True)
        l0 = _lib.launch_count()
        barrier()
        for i in range(n_prof):
            joint_eager(resident[i % n_batches])
        barrier()
        tt.set_concurrency(not a.serial)
        launches_per_step = (_lib.launch_count() - l0) / n_prof
        _lib.prof_enable(False)
        prof = _lib.prof_read()
        prof_steps = n_prof
    else:
        prof_steps = a.steps

    # ---- data-parallel self-check: after all these optimiser steps every replica must hold bit-identical weights
    #      (each rank saw DIFFERENT data; only a correct gradient / statistics all-reduce keeps them in lock step) ----
    dp_check = None
    if world > 1:
        fp = torch.stack([p.detach().double().sum() for p in model.parameters()])
        lo, hi = fp.clone(), fp.clone()
        td.all_reduce(lo, op=td.ReduceOp.MIN)
        td.all_reduce(hi, op=td.ReduceOp.MAX)
        dp_check = {"replicas_bit_identical": bool(torch.equal(lo, hi)),
                    "max_abs_param_sum_spread": float((hi - lo).abs().max().item()),
                    "steps_checked": int(max(a.warmup, 3) + 2 * a.steps + 6)}
    also = {}
    if a.also_c3 and workload_name(a).startswith("c2"):
        del graphed
        torch.cuda.empty_cache()
        also["c3"] = secondary_config(a, dev, world, rank, 128, "bf16", "c3")

    comm_name = (None if world == 1 else "peer-memory all-reduce kernels over NVLink (csrc/peer_allreduce.cu)"
                 if tdist.peer_comm() is not None else "NCCL all_reduce")
    tdist.shutdown()          # checks the peer all-reduce status word, unmaps / frees the peer regions
    if world > 1:
        td.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    # DRAM traffic of the dominant family's representative launch, from the committed `ncu --set full` capture
    try:
        tr = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        if workload_name(a).startswith("c2") and tr.get("workload") == "c2":
            traffic_tab = tr["kernels"]
        else:
            traffic_tab = {}
    except Exception:
        traffic_tab = {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    total_dev_ms = sum(v["ms"] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    ach_gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    sm_mhz = (clk or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    ffma_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    ach_tf = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
    seqs = B * world * a.steps
    line = {
        "metric": METRIC, "value": round(seqs / (ms * 1e-3), 2), "unit": "seq/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "fp32", "tf32": "fp32 (tf32 projections)",
                                                       "bf16": "fp32 recurrence, bf16 input projections"}[a.proj],
        "data": "synthetic",
        "config": bench_config(a, world),
        "e2e": {"value": round(seqs / (ms_e2e * 1e-3), 2), "unit": "seq/s",
                "h2d_bytes_per_step": B * T_LEN * X_DIM * 4, "d2h_bytes_per_step": 8 * 4,
                "ms_per_step": round(ms_e2e / a.steps, 3)},
        "gpu_launches": int(round(launches_per_step * a.steps)),
        "issue": "cuda-graph replay (1 graph launch per step)" if use_graph else "eager",
        "comm": comm_name,
        "host_issue_ms_per_step": round(host_issue_ms / a.steps, 3),
        "clocks": clk,
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": round(ach_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                     "frac": round(ach_gbs / hbm_peak, 4),
                     "traffic": (traffic_tab.get(dom) or {}).get("dram_bytes_per_launch"),
                     "traffic_note": (traffic_tab.get(dom) or {}).get("note"),
                     "algorithmic_bytes_per_launch": round(d["bytes"] / max(d["calls"], 1)),
                     "peak_source": peak_src,
                     "avg_launch_ms": round(d["ms"] / max(d["calls"], 1), 4), "launches": d["calls"],
                     "share_of_device_time": round(d["ms"] / total_dev_ms, 4),
                     "timing": "CUDA-event pair around every C-ABI call of the family, summed over eagerly issued, "
                               "serialised steps inside this bench run (graph replays cannot carry event pairs)",
                     "ffma": {"achieved": round(ach_tf, 2), "peak": round(ffma_peak, 1), "unit": "TFLOP/s",
                              "frac": round(ach_tf / ffma_peak, 4),
                              "note": "W_hh h runs on fp32 FMA pipes by design; peak = 148 SM x 128 FMA x 2 x sampled clock"}},
        "families": {k: {"ms_per_step": round(v["ms"] / prof_steps, 3), "calls_per_step": round(v["calls"] / prof_steps, 1),
                         "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0,
                         "TFLOPs": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 else 0.0}
                     for k, v in prof.items() if v["calls"]},
    }
    if dp_check is not None:
        line["dp_check"] = dp_check
    if also:
        line["also"] = also
    if not a.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(a)
    emit(line)


def main():
    a = parse()
    _guard_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

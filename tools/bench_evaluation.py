#!/usr/bin/env python3
"""the metric block of timeGAN/evaluation.py (discriminative + TSTR/TRTS predictive + statistical)
for one posture-sized pair of real / synthetic sets on one B200, next to the CPU oracle port on a bounded sample.
    python tools/bench_evaluation.py [--n 1200] [--cpu-n 48]      prints one JSON line"""
import argparse, json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1200, help="windows per domain (a posture has ~600-1300 real windows)")
    ap.add_argument("--t", type=int, default=768)
    ap.add_argument("--cpu-n", type=int, default=48)
    a = ap.parse_args()
    from timegan_b200 import evaluation as ev
    from oracle import eval_ref as E
    from oracle.make_golden_eval import make_inputs
    real, fake = make_inputs(seed=1, n=a.n, T=a.t)
    ev.evaluate_pair(real[:64], fake[:64])                      # warm-up (kernel images, cuFFT plan)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    torch.manual_seed(0)
    m = ev.evaluate_pair(real, fake)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # CPU port on a bounded sample; every part is linear in the number of windows
    r, f = real[:a.cpu_n], fake[:a.cpu_n]
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    t1 = time.perf_counter()
    torch.manual_seed(0)
    E.discriminative_score(r, f)
    E.predictive_score(f[:, :-1], f[:, -1], r[:, :-1], r[:, -1])
    E.predictive_score(r[:, :-1], r[:, -1], f[:, :-1], f[:, -1])
    t2 = time.perf_counter()
    E.statistical_similarity(r, f)
    t3 = time.perf_counter()
    cpu_s = (t3 - t1) * a.n / a.cpu_n
    print(json.dumps({"metric": "evaluation metric block (ev:190-216) windows/sec, real+synthetic", "value": round(2 * a.n / dt, 1),
                      "unit": "windows/s", "seconds": round(dt, 3), "n_per_domain": a.n, "T": a.t,
                      "metrics": {k: (round(float(v), 6) if isinstance(v, (float, np.floating)) else v) for k, v in m.items()},
                      "cpu_baseline": {"value": round(2 * a.n / cpu_s, 2), "unit": "windows/s", "kind": "port",
                                       "cores": min(8, os.cpu_count() or 1),
                                       "sample": f"oracle/eval_ref.py on {a.cpu_n} windows per domain, scaled x{a.n / a.cpu_n:g}",
                                       "networks_s": round((t2 - t1) * a.n / a.cpu_n, 1),
                                       "statistics_s": round((t3 - t2) * a.n / a.cpu_n, 1)},
                      "speedup_vs_cpu_port": round(cpu_s / dt, 1)}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""BASELINE config c5: generate_long_synth-style inference -- decode(refine_latent(gen_latent(Z))) for N synthetic
768x14 windows, chunked on one B200 (timegan_b200.generate_long_synth.generate_windows), next to the CPU oracle
port on a bounded sample.  Prints one JSON line.   python tools/bench_inference.py [--n 100000] [--hidden 56 --z 28 --layers 1]"""
import argparse, json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--t", type=int, default=768)
    ap.add_argument("--z", type=int, default=28)
    ap.add_argument("--hidden", type=int, default=56)
    ap.add_argument("--layers", type=int, default=1)
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--cpu-n", type=int, default=256)
    a = ap.parse_args()
    import timegan_b200 as tg
    from timegan_b200.generate_long_synth import generate_windows
    from oracle import timegan_ref as R
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    port = R.build_model(14, a.z, a.hidden, a.layers, 0.2)
    model = tg.TimeGAN(14, a.z, a.hidden, a.layers, 0.2)
    model.load_state_dict(port.state_dict())
    model = model.to(dev).eval()
    port.eval()
    generate_windows(model, min(a.n, a.chunk), a.t, a.z, dev, chunk=a.chunk)      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = generate_windows(model, a.n, a.t, a.z, dev, chunk=a.chunk)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert out.shape == (a.n, a.t, 14) and bool((out == out).all())
    best = None
    for th in sorted({1, min(8, os.cpu_count() or 1), os.cpu_count() or 1}):
        torch.set_num_threads(th)
        z = torch.rand(a.cpu_n, a.t, a.z)
        t1 = time.perf_counter()
        R.generate(port, z)
        c = a.cpu_n / (time.perf_counter() - t1)
        if best is None or c > best[0]:
            best = (c, th)
    print(json.dumps({"metric": "TimeGAN inference windows/sec (T=768,C=14), host array out", "value": round(a.n / dt, 1),
                      "unit": "windows/s", "n_windows": a.n, "seconds": round(dt, 3),
                      "config": {"workload": f"c5: G->S->R eval forward, z={a.z} h={a.hidden} L={a.layers}, chunk {a.chunk}, "
                                             "device Philox noise, D2H overlapped, result in host memory"},
                      "cpu_baseline": {"value": round(best[0], 1), "unit": "windows/s", "cores": best[1], "kind": "port",
                                       "sample": f"{a.cpu_n} windows, one batch, oracle/timegan_ref.generate"},
                      "speedup_vs_cpu_port": round(a.n / dt / best[0], 1)}))


if __name__ == "__main__":
    main()

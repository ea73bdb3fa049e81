#!/usr/bin/env python3
"""Golden values of the evaluation metrics, produced by the UNMODIFIED reference (timeGAN/evaluation.py) on CPU.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference); writes tests/golden/eval_small.npz.
The reference imports matplotlib at module top (ev:30), which this image does not have; a stub module named
matplotlib is placed in sys.modules so that the file imports UNCHANGED -- none of the metric functions touch it.
Inputs are regenerated from the recorded seed by the consumers (make_inputs below), so only seeds + outputs are stored.
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/timeGAN")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def make_inputs(seed=0, n=40, T=320, C=14):
    """Band-limited 'real' windows in [0,1] and a noisier, slightly shifted 'synthetic' set (both float32)."""
    rng = np.random.default_rng(seed)
    t = np.arange(T)[None, :, None] / 128.0
    f = rng.uniform(4.0, 30.0, size=(n, 1, C))
    ph = rng.uniform(0, 2 * np.pi, size=(n, 1, C))
    real = 0.5 + 0.3 * np.sin(2 * np.pi * f * t + ph) + 0.05 * rng.standard_normal((n, T, C))
    f2 = rng.uniform(6.0, 36.0, size=(n, 1, C))
    fake = 0.52 + 0.25 * np.sin(2 * np.pi * f2 * t + ph[::-1]) + 0.09 * rng.standard_normal((n, T, C))
    return np.clip(real, 0, 1).astype(np.float32), np.clip(fake, 0, 1).astype(np.float32)


def main():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, str(REF))
    import evaluation as ev            # the reference, unmodified

    real, fake = make_inputs()
    out = {"seed": 0, "n": real.shape[0], "T": real.shape[1], "C": real.shape[2]}
    torch.manual_seed(0)
    acc, auc = ev.discriminative_score(real, fake)
    out["disc"] = np.array([acc, auc])
    torch.manual_seed(1)
    out["pred_tstr"] = np.array(ev.predictive_score(fake[:, :-1], fake[:, -1], real[:, :-1], real[:, -1]))
    torch.manual_seed(2)
    out["pred_trts"] = np.array(ev.predictive_score(real[:, :-1], real[:, -1], fake[:, :-1], fake[:, -1]))
    out["stat"] = np.array(ev.statistical_similarity(real, fake, fs=128.0))
    out["acf_seq"] = np.array([ev.autocorr_seq(real[i, :, c], 96) for i in range(4) for c in range(3)])
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez(OUT / "eval_small.npz", **out)
    print({k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()

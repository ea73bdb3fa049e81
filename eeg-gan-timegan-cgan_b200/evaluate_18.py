"""Per-(posture, condition) evaluation -- drop-in for timeGAN/evaluate_18.py (18 models), on the GPU.

Same metric functions as `evaluation.py` of this package (the reference duplicates them verbatim, e18:44-144 ==
ev:41-139); what this file adds is the reference's pairing rule and outputs:
    find_synth_npz             e18:146-152   synthetic_long.npz, else synthetic.npz, else the first *.npz of the run
    load_pairs_by_condition    e18:154-172   (posture, condition) -> (real[:m], fake[:m])
    main                       e18:175-306   metrics_per_posture_condition.csv + metrics_global.csv (no figures)
"""
import argparse
from pathlib import Path

import numpy as np
import torch

from .evaluation import (RNNClassifier, RNNPredictor, autocorr_seq, discriminative_score, evaluate_pair,  # noqa: F401
                         predictive_score, statistical_similarity)


def find_synth_npz(run_dir: Path):
    for name in ("synthetic_long.npz", "synthetic.npz"):
        if (run_dir / name).exists():
            return run_dir / name
    others = sorted(run_dir.glob("*.npz"))
    return others[0] if others else None


def load_pairs_by_condition(real_dir: Path, synth_dir: Path):
    pairs = {}
    for p in range(1, 10):
        for cond in ("with_exo", "no_exo"):
            rfp = real_dir / f"posture{p}_{cond}.npz"
            sfp = find_synth_npz(synth_dir / f"posture{p}_{cond}")
            if rfp.exists() and sfp is not None and sfp.exists():
                r = np.load(rfp)["X"].astype(np.float32)
                f = np.load(sfp)["X"].astype(np.float32)
                m = min(len(r), len(f))
                if m > 0:
                    pairs[(p, cond)] = (r[:m], f[:m])
    return pairs


def main(argv=None):
    import pandas as pd
    ap = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--real_dir", type=str, default="./preprocessed")
    ap.add_argument("--synth_dir", type=str, default="./timegan_runs")
    ap.add_argument("--out", type=str, default="./eval_out")
    ap.add_argument("--fs", type=float, default=128.0)
    ap.add_argument("--tsne_max", type=int, default=6000, help="accepted for CLI compatibility (no figures here)")
    args = ap.parse_args(argv)
    np.random.seed(0)
    torch.manual_seed(0)
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    pairs = load_pairs_by_condition(Path(args.real_dir), Path(args.synth_dir))
    if not pairs:
        raise SystemExit("No (posture, condition) pairs found with matching real and synthetic.")
    rows, all_real, all_fake = [], [], []
    for (posture, cond) in sorted(pairs.keys()):
        real, fake = pairs[(posture, cond)]
        rows.append(dict({"posture": posture, "condition": cond}, **evaluate_pair(real, fake, fs=args.fs)))
        all_real.append(real)
        all_fake.append(fake)
    fp = out / "metrics_per_posture_condition.csv"
    pd.DataFrame(rows).sort_values(["posture", "condition"]).to_csv(fp, index=False)
    print(f"Wrote {fp}")
    R, F = np.concatenate(all_real, axis=0), np.concatenate(all_fake, axis=0)
    pd.DataFrame([evaluate_pair(R, F, fs=args.fs)]).to_csv(out / "metrics_global.csv", index=False)
    print(f"Wrote {out / 'metrics_global.csv'}")


if __name__ == "__main__":
    main()

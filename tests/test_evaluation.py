"""the evaluation metrics of timeGAN/evaluation.py.
not-gpu: the oracle restatement (oracle/eval_ref.py) against golden values produced by the UNMODIFIED reference
(oracle/make_golden_eval.py -> tests/golden/eval_small.npz).  gpu: the device implementation against both."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden" / "eval_small.npz"


def _inputs():
    sys.path.insert(0, str(ROOT / "oracle"))
    from oracle.make_golden_eval import make_inputs
    return make_inputs()


def test_oracle_port_matches_reference_goldens():
    from oracle import eval_ref as E
    g = np.load(GOLD)
    real, fake = _inputs()
    assert real.shape == (int(g["n"]), int(g["T"]), int(g["C"]))
    torch.manual_seed(0)
    assert np.allclose(E.discriminative_score(real, fake), g["disc"], rtol=0, atol=1e-12)
    torch.manual_seed(1)
    assert np.allclose(E.predictive_score(fake[:, :-1], fake[:, -1], real[:, :-1], real[:, -1]), g["pred_tstr"], rtol=1e-9)
    assert np.allclose(E.statistical_similarity(real, fake, fs=128.0), g["stat"], rtol=1e-12)
    acf = [E.autocorr_seq(real[i, :, c], 96) for i in range(4) for c in range(3)]
    assert np.allclose(acf, g["acf_seq"], rtol=1e-12)


@pytest.mark.gpu
def test_statistics_match_reference():
    from timegan_b200 import evaluation as ev
    g = np.load(GOLD)
    real, fake = _inputs()
    got = ev.statistical_similarity(real, fake, fs=128.0)
    assert np.allclose(got, g["stat"], rtol=1e-6), (got, g["stat"])
    dev = torch.device("cuda:0")
    sc = ev.acf_scores(torch.tensor(real[:4, :, :3]).to(dev), 96).cpu().numpy().reshape(-1)
    assert np.allclose(sc, g["acf_seq"], rtol=1e-9, atol=1e-12)
    assert abs(ev.autocorr_seq(real[1, :, 2], 96) - g["acf_seq"][5]) < 1e-12
    # edge cases of ev:63-71: constant series -> 0 ; window shorter than maxlag -> lags 1..T-1 only
    const = torch.full((1, 50, 1), 0.25, device=dev)
    assert float(ev.acf_scores(const, 96)) == 0.0
    from oracle import eval_ref as E
    short = np.random.default_rng(3).random(20).astype(np.float32)
    assert abs(ev.autocorr_seq(short, 12) - E.autocorr_seq(short, 12)) < 1e-12
    # ... and when the last lag leaves ONE sample, np.corrcoef (hence the reference's mean) is NaN: so is ours
    with np.errstate(all="ignore"):
        assert np.isnan(E.autocorr_seq(short, 96)) and np.isnan(ev.autocorr_seq(short, 96))
    # Welch periodogram against scipy on an odd-length, short window
    import scipy.signal as sig
    x = np.random.default_rng(4).random((3, 300, 2)).astype(np.float32)
    _, ref = sig.welch(x, fs=128.0, axis=1, nperseg=256)
    mine = ev.welch_psd(torch.tensor(x).to(dev), 128.0).cpu().numpy()
    assert np.allclose(mine, ref, rtol=1e-5, atol=1e-8)       # scipy keeps float32 for float32 input; ours is fp64


@pytest.mark.gpu
def test_posthoc_networks_match_reference():
    """Same seed -> same initial weights (FusedGRU / nn.Linear draw like nn.GRU / nn.Linear) -> the 20 / 50 full-batch
    Adam epochs follow the CPU reference to fp32 training noise."""
    from timegan_b200 import evaluation as ev
    from oracle import eval_ref as E
    g = np.load(GOLD)
    real, fake = _inputs()
    torch.manual_seed(0)
    acc, auc, p = ev.discriminative_score(real, fake, return_probs=True)
    torch.manual_seed(0)
    _, _, p_ref = E.discriminative_score(real, fake, return_probs=True)
    assert np.abs(p - p_ref).max() < 2e-4, np.abs(p - p_ref).max()
    assert abs(auc - g["disc"][1]) < 0.02 and abs(acc - g["disc"][0]) < 0.05
    torch.manual_seed(1)
    rmse, r2 = ev.predictive_score(fake[:, :-1], fake[:, -1], real[:, :-1], real[:, -1])
    assert abs(rmse - g["pred_tstr"][0]) < 1e-3 * g["pred_tstr"][0] and abs(r2 - g["pred_tstr"][1]) < 5e-3
    torch.manual_seed(2)
    rmse, r2 = ev.predictive_score(real[:, :-1], real[:, -1], fake[:, :-1], fake[:, -1])
    assert abs(rmse - g["pred_trts"][0]) < 1e-3 * g["pred_trts"][0] and abs(r2 - g["pred_trts"][1]) < 5e-3


@pytest.mark.gpu
def test_evaluation_entry_points_write_the_reference_csvs(tmp_path):
    """evaluation.main (ev:170-238) and evaluate_18.main (e18:175-262) on a two-posture toy layout: same file names
    and columns as the reference's CSVs; evaluate_18 prefers synthetic_long.npz (e18:146-152)."""
    import pandas as pd
    from timegan_b200 import evaluation as ev, evaluate_18 as e18
    from oracle.make_golden_eval import make_inputs
    real_dir, synth_dir = tmp_path / "preprocessed", tmp_path / "timegan_runs"
    real_dir.mkdir()
    for p, cond, seed in ((1, "with_exo", 1), (1, "no_exo", 2), (3, "no_exo", 3)):
        r, f = make_inputs(seed=seed, n=14, T=300)
        np.savez(real_dir / f"posture{p}_{cond}.npz", X=r, fs=128.0)
        run = synth_dir / f"posture{p}_{cond}"
        run.mkdir(parents=True)
        np.savez(run / "synthetic.npz", X=f)
    np.savez(synth_dir / "posture3_no_exo" / "synthetic_long.npz", X=make_inputs(seed=9, n=20, T=300)[1])
    cols = ["disc_acc", "disc_auc", "rmse_tstr", "r2_tstr", "rmse_trts", "r2_trts", "psd_diff", "acf_diff", "coh_diff",
            "n_real", "n_fake", "seq_len", "n_ch"]
    ev.main(["--real_dir", str(real_dir), "--synth_dir", str(synth_dir), "--out", str(tmp_path / "o1")])
    per = pd.read_csv(tmp_path / "o1" / "metrics_per_posture.csv")
    assert list(per.columns) == ["posture"] + cols and list(per["posture"]) == [1, 3]
    assert list(per["n_real"]) == [28, 14]                      # posture 1: both conditions concatenated
    glob = pd.read_csv(tmp_path / "o1" / "metrics_global.csv")
    assert list(glob.columns) == cols and int(glob["n_real"][0]) == 42 and np.isfinite(glob.values).all()
    e18.main(["--real_dir", str(real_dir), "--synth_dir", str(synth_dir), "--out", str(tmp_path / "o2")])
    per = pd.read_csv(tmp_path / "o2" / "metrics_per_posture_condition.csv")
    assert list(per.columns) == ["posture", "condition"] + cols and len(per) == 3
    assert e18.find_synth_npz(synth_dir / "posture3_no_exo").name == "synthetic_long.npz"

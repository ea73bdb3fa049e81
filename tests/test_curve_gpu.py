"""Loss-curve parity over the full schedule at BASELINE config c1 (z=h=24, 3 layers, batch 32, N=256x768x14
synthetic, 5 AE + 5 SUP epochs, 200 joint steps): timegan_b200.train_single_npz on the B200, with the reference's
CPU random stream replayed (noise="host"), against tests/golden/c1_curve/ -- the log the UNMODIFIED reference
wrote for the same seed on the CPU (oracle/make_golden_curve.py).

A GAN trajectory amplifies rounding differences, so the check is "tracks": the pre-training curves and the first
joint steps agree to ~1e-3, every logged column stays within a few percent of the reference over all 200 steps."""
import csv
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "c1_curve"


def _read_csv(p):
    with open(p) as f:
        rows = list(csv.DictReader(f))
    cols = ["loss_D", "acc_D", "loss_G", "loss_adv", "loss_sup", "loss_rec", "loss_cov", "loss_acf"]
    return {c: np.array([float(r[c]) for r in rows]) for c in cols}


def test_five_phase_curves_track_reference(tmp_path, capsys):
    from timegan_b200 import train_timegan as tt
    cfg = json.loads((GOLD / "config.json").read_text())
    X = np.random.default_rng(0).random((cfg["N"], cfg["T"], cfg["C"]), dtype=np.float32)
    npz = tmp_path / "posture1_synth.npz"
    np.savez(npz, X=X, fs=128.0)
    tt.train_single_npz(npz, tmp_path / "run", batch_size=cfg["batch_size"], ae_epochs=cfg["ae_epochs"],
                        sup_epochs=cfg["sup_epochs"], gan_steps=cfg["gan_steps"], layers=cfg["layers"],
                        dropout=cfg["dropout"], seed=cfg["seed"], device=torch.device("cuda:0"),
                        z_dim=cfg["z_dim"], hidden_dim=cfg["hidden_dim"], noise="host")
    out = capsys.readouterr().out
    # ---- phases 1 and 2: epoch means printed with 5 decimals (tt:144,163) ----
    gold_pre = (GOLD / "pretrain_log.txt").read_text()
    num = lambda key, text: [float(v) for v in re.findall(key + r"=(\d+\.\d{5})", text)]
    ref_ae, ref_sup = num("recon", gold_pre), num("sup", gold_pre)
    got_ae, got_sup = num("recon", out), num("sup", out)
    assert len(got_ae) == len(ref_ae) == cfg["ae_epochs"] and len(got_sup) == len(ref_sup) == cfg["sup_epochs"]
    np.testing.assert_allclose(got_ae, ref_ae, rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(got_sup, ref_sup, rtol=2e-2, atol=2e-5)
    # ---- phase 3: the per-step log (tt:318-319) ----
    ref, got = _read_csv(GOLD / "train_log.csv"), _read_csv(tmp_path / "run" / "train_log.csv")
    report = {}
    # the throttle (tt:211-215) multiplies loss_D by a function of acc_D, which moves in quanta of 1/64: a step
    # where one thresholded probability flips legitimately changes loss_D by ~10 %, so loss_D is compared on
    # the steps where both runs counted the same accuracy (the flips themselves are bounded below)
    same_acc = np.abs(got["acc_D"] - ref["acc_D"]) < 1e-6
    for c in ref:
        assert got[c].shape == ref[c].shape == (cfg["gan_steps"],)
        scale = np.maximum(np.abs(ref[c]), 1e-2 if c != "loss_sup" else 1e-5)
        rel = np.abs(got[c] - ref[c]) / scale
        if c == "loss_D":
            rel = np.where(same_acc, rel, 0.0)
        report[c] = (float(rel[:10].max()), float(rel.max()), float(rel.mean()))
    report["acc_flips"] = int((~same_acc).sum())
    print("curve deviation (first-10 max, max, mean):", json.dumps(report))
    (tmp_path / "report.json").write_text(json.dumps(report))
    import os
    if os.environ.get("GRAFT_REPO_ROOT"):
        Path(os.environ["GRAFT_REPO_ROOT"], "gpurun_out").mkdir(exist_ok=True)
        Path(os.environ["GRAFT_REPO_ROOT"], "gpurun_out", "curve_report.json").write_text(json.dumps(report))
        import shutil
        shutil.copy(tmp_path / "run" / "train_log.csv",
                    Path(os.environ["GRAFT_REPO_ROOT"], "gpurun_out", "curve_train_log.csv"))
    # late in the run D hovers at p ~ 0.5 where the thresholded count is noise-like; each flip is one quantum
    assert report["acc_flips"] <= cfg["gan_steps"] // 2
    for c, v in report.items():
        if c == "acc_flips":
            continue
        first10, worst, mean = v
        if c == "acc_D":
            # a count of 64 probabilities thresholded at 0.5 (tt:206-207).  Late in this run the reference's D
            # outputs sit within ~1e-6 of 0.5, so the count there is the sign of rounding noise (the continuous
            # losses stay within 1 %); require exact agreement while D is decisive and 3/4 agreement overall.
            assert np.abs(got[c] - ref[c])[:40].max() <= 1e-6, (c, report[c])
            assert (np.abs(got[c] - ref[c]) <= 1.0 / 64 + 1e-6).mean() >= 0.75, (c, report[c])
            continue
        assert first10 < 5e-3, (c, report[c])
        assert mean < 2e-2 and worst < 1e-1, (c, report[c])
    # artefacts of the reference schedule exist with the reference schema
    ck = torch.load(tmp_path / "run" / "ckpt_latest.pt", map_location="cpu")
    assert set(ck) == {"step", "model", "optG", "optD", "meta"} and ck["step"] == cfg["gan_steps"]
    assert {k: ck["meta"][k] for k in ("npz", "z_dim", "h_dim")} == {"npz": "posture1_synth.npz", "z_dim": 24, "h_dim": 24}
    syn = np.load(tmp_path / "run" / "synthetic.npz")["X"]
    assert syn.shape == X.shape and syn.dtype == np.float32

"""train_single_npz end to end on the GPU with production settings (on-device noise): eager vs CUDA-graph replay
give the same log, artefacts follow the reference's schema (train_timegan.py:315-320, 58-61, 416-420), ragged last
batches and the deferred-logging path work, and main.py's config route reaches the same function."""
import csv
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rows(p):
    with open(p) as f:
        return list(csv.DictReader(f))


def test_train_single_npz_eager_equals_graph_and_writes_reference_artefacts(tmp_path, capsys):
    from timegan_b200 import train_timegan as tt
    X = np.random.default_rng(1).random((44, 64, 14), dtype=np.float32)     # 44 = 5 full batches of 8 + 4 ragged
    npz = tmp_path / "posture3_with_exo.npz"
    np.savez(npz, X=X, fs=128.0)
    kw = dict(batch_size=8, ae_epochs=1, sup_epochs=1, gan_steps=14, layers=2, dropout=0.0, seed=7,
              device=torch.device("cuda:0"), z_dim=16, hidden_dim=16, acf_max_lag=16)
    logs = {}
    for mode in ("eager", "graph"):
        tt._DEFAULT_NOISE.clear()           # same Philox stream for both runs
        out = tmp_path / mode
        assert tt.train_single_npz(npz, out, log_every=5, graph=(mode == "graph"), **kw) is True
        rows = _rows(out / "train_log.csv")
        assert [int(r["step"]) for r in rows] == list(range(1, 15))
        assert list(rows[0].keys()) == ["step", "phase", "loss_D", "acc_D", "loss_G", "loss_adv", "loss_sup",
                                        "loss_rec", "loss_cov", "loss_acf"]
        logs[mode] = np.array([[float(r[c]) for c in list(r.keys())[2:]] for r in rows])
        assert np.isfinite(logs[mode]).all()
        ck = torch.load(out / "ckpt_latest.pt", map_location="cpu")
        assert ck["step"] == 14
        assert {k: ck["meta"][k] for k in ("npz", "z_dim", "h_dim")} == {"npz": npz.name, "z_dim": 16, "h_dim": 16}
        assert ck["meta"]["best_loss"] == pytest.approx(min(float(r["loss_G"]) for r in rows), rel=1e-6)
        assert set(ck["optG"]["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(ck["optG"]["state"][0]["step"]) == 14
        assert (out / "ckpt_best.pt").exists()
        syn = np.load(out / "synthetic.npz")["X"]
        assert syn.shape == X.shape and syn.dtype == np.float32 and np.isfinite(syn).all()
    np.testing.assert_allclose(logs["graph"], logs["eager"], rtol=5e-3, atol=2e-5)
    text = capsys.readouterr().out
    assert "[AE] epoch 1/1" in text and "[SUP] epoch 1/1" in text and "Saved synthetic" in text


def test_snapshot_if_better_kernel():
    """csrc/optim.cu `snapshot_if_better`: the multi-tensor copy happens only when value < best, and (best, step)
    follow (tt:410-413 evaluated on the device)."""
    import ctypes as C
    from timegan_b200._lib import lib, check, ptr, stream_ptr
    dev = torch.device("cuda:0")
    src = [torch.randn(n, device=dev) for n in (5, 4096, 10001)] * 20          # 60 tensors: two launches of <= 48
    dst = [torch.zeros_like(t) for t in src]
    n = len(src)
    dp = (C.c_void_p * n)(*[t.data_ptr() for t in dst])
    sp = (C.c_void_p * n)(*[t.data_ptr() for t in src])
    sz = (C.c_longlong * n)(*[t.numel() for t in src])
    best = torch.full((1,), float("inf"), device=dev)
    bstep = torch.full((1,), -1.0, device=dev)

    def call(v, step):
        val = torch.tensor([v], device=dev)
        check(lib.tg_snapshot_if_better(stream_ptr(), n, dp, sp, sz, ptr(val), ptr(best), ptr(bstep), float(step)), "snap")
    call(2.0, 1)
    assert best.item() == 2.0 and bstep.item() == 1.0 and all(torch.equal(a, b) for a, b in zip(dst, src))
    old = [t.clone() for t in src]
    for t in src:
        t.add_(1.0)
    call(3.0, 2)                                             # worse: nothing moves
    assert best.item() == 2.0 and bstep.item() == 1.0 and all(torch.equal(a, b) for a, b in zip(dst, old))
    call(2.0, 3)                                             # equal is not better (strict <, like the reference)
    assert bstep.item() == 1.0
    call(float("nan"), 4)                                    # NaN never wins
    assert bstep.item() == 1.0
    call(1.5, 5)
    assert best.item() == 1.5 and bstep.item() == 5.0 and all(torch.equal(a, b) for a, b in zip(dst, src))


def test_default_path_is_the_fast_path_and_keeps_the_per_step_best_rule(tmp_path):
    """train_single_npz with DEFAULT extras: the joint step replays from the CUDA graph (inter-layer dropout 0.2
    inside the captured step), the host synchronises every 25 steps only, every step still has its CSV row, and
    ckpt_best.pt holds the weights AFTER THE STEP WITH THE LOWEST loss_G -- the reference's per-step rule (tt:410-413)
    -- which a second run stopped at exactly that step reproduces."""
    from timegan_b200 import train_timegan as tt
    X = np.random.default_rng(5).random((32, 48, 14), dtype=np.float32)
    npz = tmp_path / "posture2_no_exo.npz"
    np.savez(npz, X=X, fs=128.0)
    kw = dict(batch_size=8, ae_epochs=1, sup_epochs=1, gan_steps=40, layers=2, dropout=0.2, seed=11,
              device=torch.device("cuda:0"), z_dim=16, hidden_dim=16, acf_max_lag=16)
    seen = []
    orig = tt.GraphedJointStep.__call__

    def spy(self, x, std):
        out = orig(self, x, std)
        seen.append(self.graph is not None)
        return out
    tt.GraphedJointStep.__call__ = spy
    try:
        assert tt.train_single_npz(npz, tmp_path / "a", **kw) is True
    finally:
        tt.GraphedJointStep.__call__ = orig
    assert len(seen) == 40 and sum(seen) >= 36                 # graph replay is what runs by default
    rows = _rows(tmp_path / "a" / "train_log.csv")
    assert [int(r["step"]) for r in rows] == list(range(1, 41))
    g = np.array([float(r["loss_G"]) for r in rows])
    assert np.isfinite(g).all()
    best_step = int(np.argmin(g)) + 1                            # first step that reaches the minimum
    ck = torch.load(tmp_path / "a" / "ckpt_best.pt", map_location="cpu", weights_only=False)
    assert ck["step"] == best_step and ck["meta"]["best"] is True
    assert float(ck["optG"]["state"][0]["step"]) == best_step and float(ck["optD"]["state"][0]["step"]) == best_step
    lr = 1e-3 * 0.5 ** sum(best_step >= m for m in (20, 30))
    assert ck["optG"]["param_groups"][0]["lr"] == pytest.approx(lr)
    # same seed, stopped right after that step: its ckpt_latest is what the reference would have saved as best
    tt.train_single_npz(npz, tmp_path / "b", stop_after=best_step, **kw)
    ref = torch.load(tmp_path / "b" / "ckpt_latest.pt", map_location="cpu", weights_only=False)
    assert ref["step"] == best_step
    for k, v in ref["model"].items():
        assert torch.allclose(ck["model"][k], v, rtol=1e-4, atol=1e-6), k
    for i, st in ref["optG"]["state"].items():
        assert torch.allclose(ck["optG"]["state"][i]["exp_avg"], st["exp_avg"], rtol=1e-3, atol=1e-7), i


def test_main_config_route(tmp_path, monkeypatch):
    from timegan_b200 import main as tmain
    data = tmp_path / "preprocessed"
    data.mkdir()
    np.savez(data / "posture1_no_exo.npz", X=np.random.default_rng(2).random((10, 32, 14), dtype=np.float32))
    cfg = {"data_dir": str(data), "out_dir": str(tmp_path / "runs"), "batch_size": 4, "ae_epochs": 1, "sup_epochs": 1,
           "gan_steps": 3, "layers": 1, "dropout": 0.2, "lr_d": 0.0003, "alpha_sup": 3.0, "beta_rec": 0.15,
           "inst_noise_start": 0.25, "inst_noise_end": 0.05, "d_max_acc": 0.68, "gamma_cov": 0.03, "gamma_acf": 0.02,
           "acf_max_lag": 48}
    p = tmp_path / "cfg.json"
    p.write_text(json.dumps(cfg))
    tmain.main(["--config", str(p)])
    run = tmp_path / "runs" / "posture1_no_exo"
    assert len(_rows(run / "train_log.csv")) == 3
    ck = torch.load(run / "ckpt_latest.pt", map_location="cpu")
    assert ck["meta"]["z_dim"] == 28 and ck["meta"]["h_dim"] == 56       # adaptive_dims(14, 32) (tt:50-55)
    # the checkpoint loads into the inference entry point (generate_long_synth.py:96-102)
    from timegan_b200.generate_long_synth import load_model, generate_windows
    m = load_model(run / "ckpt_latest.pt", 14, torch.device("cuda:0"))
    out = generate_windows(m, 5, 100, 28, torch.device("cuda:0"))
    assert out.shape == (5, 100, 14) and np.isfinite(out).all()

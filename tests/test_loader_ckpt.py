"""SURVEY.md 8f rows N4 (device-resident loader) and N2 (asynchronous checkpoints + resume)."""
import csv
from pathlib import Path

import numpy as np
import pytest
import torch


def test_device_loader_replays_the_dataloader_shuffle_exactly():
    """DeviceLoader yields the batches of DataLoader(shuffle=True, drop_last=False) (tt:33-37) in the same order,
    ragged tail included, and leaves the global CPU generator in the same state -- so a run that replays the
    reference's random stream sees identical data and identical noise afterwards."""
    from timegan_b200.train_timegan import DeviceLoader, make_loader
    X = np.arange(26 * 3 * 2, dtype=np.float32).reshape(26, 3, 2)        # N=26: the reference's one-short-batch case
    for bs in (8, 26, 64):
        torch.manual_seed(123)
        ref_loader = make_loader(X, bs)
        ref = [[b[0].clone() for b in ref_loader] for _ in range(2)]      # two epochs
        ref_state = torch.get_rng_state()
        ref_draw = torch.rand(3)
        torch.manual_seed(123)
        ours_loader = DeviceLoader(X, bs, "cpu")
        ours = [[b[0].clone() for b in ours_loader] for _ in range(2)]
        assert torch.equal(torch.get_rng_state(), ref_state)
        assert torch.equal(torch.rand(3), ref_draw)
        assert len(ours_loader) == len(ref_loader)
        for e in range(2):
            assert len(ours[e]) == len(ref[e])
            for a, b in zip(ours[e], ref[e]):
                assert a.shape == b.shape and torch.equal(a, b)


@pytest.mark.gpu
def test_async_checkpoint_and_resume(tmp_path):
    """A run stopped at step 6 of 10 and resumed finishes with the step counter, LR schedule, instance-noise decay
    and optimiser clocks of an uninterrupted run; the log is appended to (one header, rows 1..10 once each); the
    checkpoint written behind the stream has the reference's schema and loads into torch.optim.Adam."""
    import timegan_b200 as tg
    from timegan_b200 import train_timegan as tt
    rng = np.random.default_rng(0)
    npz = tmp_path / "posture1_synth.npz"
    np.savez(npz, X=rng.random((24, 32, 14), dtype=np.float32), fs=128.0)
    out = tmp_path / "run"
    kw = dict(batch_size=8, ae_epochs=1, sup_epochs=1, gan_steps=10, layers=2, dropout=0.0, seed=3, z_dim=8,
              hidden_dim=8, acf_max_lag=8, device=torch.device("cuda:0"), ckpt_every=3)
    tt.train_single_npz(npz, out, stop_after=6, **kw)
    ck = torch.load(out / "ckpt_latest.pt", map_location="cpu", weights_only=False)
    assert set(ck) == {"step", "model", "optG", "optD", "meta"} and ck["step"] == 6
    assert not (out / "synthetic.npz").exists()
    # the optimiser state is torch.optim.Adam's
    ref_model = tg.TimeGAN(14, 8, 8, 2, 0.0)
    ref_model.load_state_dict(ck["model"])
    adam = torch.optim.Adam(ref_model.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9))
    adam.load_state_dict(ck["optD"])
    assert float(adam.state_dict()["state"][0]["step"]) == 6
    w6 = {k: v.clone() for k, v in ck["model"].items()}

    tt.train_single_npz(npz, out, resume=True, **kw)
    ck2 = torch.load(out / "ckpt_latest.pt", map_location="cpu", weights_only=False)
    assert ck2["step"] == 10
    assert float(ck2["optG"]["state"][0]["step"]) == 10
    # MultiStepLR milestones [5, 7] at gamma 0.5: both passed -> lr_g/4, exactly as without the interruption
    assert abs(ck2["optG"]["param_groups"][0]["lr"] - 1e-3 * 0.25) < 1e-12
    assert abs(ck2["optD"]["param_groups"][0]["lr"] - 2e-4 * 0.25) < 1e-12
    assert any((ck2["model"][k] - w6[k]).abs().max() > 0 for k in w6)            # it did train on
    assert all(torch.isfinite(v).all() for v in ck2["model"].values())
    rows = list(csv.reader(open(out / "train_log.csv")))
    assert rows[0][0] == "step" and sum(r[0] == "step" for r in rows) == 1
    assert [int(r[0]) for r in rows[1:]] == list(range(1, 11))
    assert (out / "synthetic.npz").exists()

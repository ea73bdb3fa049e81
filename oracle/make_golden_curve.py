#!/usr/bin/env python3
"""Generate the config-1 golden loss curve by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Runs in the GPU-less build container (needs /root/reference);
its outputs are committed under tests/golden/c1_curve/ and read by the -m gpu parity tests.

What it does:
  * synthetic NPZ  X = default_rng(0).random((256,768,14), float32)  -> posture1_synth.npz
  * imports /root/reference/timeGAN/train_timegan.py unmodified, patches ONLY the runtime
    attribute `adaptive_dims` (train_timegan.py:50-55 forces z=28,h=56) to return (24,24)
  * train_single_npz(batch_size=32, ae_epochs=5, sup_epochs=5, gan_steps=200, layers=3,
    dropout=0.0, device=cpu), every other kwarg at its train_timegan.py:281-303 default
  * captures the AE/SUP epoch prints (train_timegan.py:144,163) and train_log.csv (tt:318-319)
"""
import contextlib, io, json, shutil, sys, tempfile, time
from pathlib import Path
import numpy as np
import torch

REF = Path("/root/reference/timeGAN")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "c1_curve"

def main():
    sys.path.insert(0, str(REF))
    import train_timegan as tt  # noqa
    H = 24
    tt.adaptive_dims = lambda C, T: (H, H)
    torch.set_num_threads(int(sys.argv[1]) if len(sys.argv) > 1 else 4)
    cfg = dict(batch_size=32, ae_epochs=5, sup_epochs=5, gan_steps=200, layers=3, dropout=0.0, seed=42)
    tmp = Path(tempfile.mkdtemp(prefix="c1curve_"))
    X = np.random.default_rng(0).random((256, 768, 14), dtype=np.float32)
    np.savez(tmp / "posture1_synth.npz", X=X, fs=128.0)
    lines = []
    class Tee(io.TextIOBase):
        def write(self, s):
            sys.__stdout__.write(s); sys.__stdout__.flush(); lines.append(s); return len(s)
    t0 = time.time()
    with contextlib.redirect_stdout(Tee()):
        tt.train_single_npz(tmp / "posture1_synth.npz", tmp / "run", device=torch.device("cpu"), **cfg)
    wall = time.time() - t0
    OUT.mkdir(parents=True, exist_ok=True)
    shutil.copy(tmp / "run" / "train_log.csv", OUT / "train_log.csv")
    (OUT / "pretrain_log.txt").write_text("".join(l for l in lines if l.startswith("[AE]") or l.startswith("[SUP]")))
    meta = dict(cfg, z_dim=H, hidden_dim=H, N=256, T=768, C=14, data="default_rng(0).random((256,768,14),float32)",
                torch=torch.__version__, threads=torch.get_num_threads(), wall_s=round(wall, 1),
                reference="/root/reference/timeGAN/train_timegan.py:train_single_npz (unmodified; adaptive_dims patched at runtime)")
    (OUT / "config.json").write_text(json.dumps(meta, indent=1))
    shutil.rmtree(tmp, ignore_errors=True)

if __name__ == "__main__":
    main()

"""B200-native drop-in for timeGAN/generate_long_synth.py (gl:43-131) and the generation tail of
train_single_npz (train_timegan.py:416-420): load a checkpoint, run generator -> supervisor -> recovery on
U(0,1) noise for any N and T, optionally de-normalise, write `synthetic*.npz`.

Differences by design: the chain runs in chunks of `chunk` windows (the reference pushes all N windows through
in ONE batch, gl:118 -- 100k windows would need 8.6 GB for Z alone) with the device->host copy of chunk i
overlapping the kernels of chunk i+1, and `num_layers` is inferred from the checkpoint instead of being
hard-coded to 1 (gl:100).  Same CLI flags and output files.
"""
import argparse
import re
from pathlib import Path

import numpy as np
import torch

from .timegan_model import TimeGAN


def device_autoselect():
    if not torch.cuda.is_available():
        raise RuntimeError("timegan_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def sample_noise(batch_size: int, seq_len: int, z_dim: int, device, noise=None):
    from .train_timegan import _noise_source
    return _noise_source(noise, device).rand(batch_size, seq_len, z_dim)


def infer_posture_cond(run_dir_name: str):
    m = re.match(r"posture(\d+)_(with_exo|no_exo)$", run_dir_name)
    if not m:
        return None, None
    return int(m.group(1)), m.group(2)


def layers_in_state_dict(sd) -> int:
    """Number of GRU layers stored in a TimeGAN state_dict (keys embedder.rnn.rnn.weight_ih_l{k})."""
    ks = [int(k.rsplit("_l", 1)[1]) for k in sd if k.startswith("embedder.rnn.rnn.weight_ih_l")]
    return max(ks) + 1 if ks else 1


@torch.no_grad()
def generate_windows(model: TimeGAN, n: int, seq_len: int, z_dim: int, device, chunk: int = 4096, noise=None,
                     out: np.ndarray = None) -> np.ndarray:
    """decode(refine_latent(gen_latent(Z))) for n windows (gl:117-121 == tt:417-419), chunked and pipelined.

    Returns float32 (n, seq_len, x_dim) on the host.  With host-replay noise and chunk >= n the draw is the
    reference's single rand(n, T, z) call."""
    x_dim = model.recovery.out.out_features
    if out is None:
        out = np.empty((n, seq_len, x_dim), dtype=np.float32)
    chunk = max(1, min(chunk, n))
    pinned = [torch.empty((chunk, seq_len, x_dim), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    events = [None, None]
    spans = [None, None]
    copy_stream = torch.cuda.Stream(device=device)
    i = 0
    for k, start in enumerate(range(0, n, chunk)):
        nb = min(chunk, n - start)
        slot = k & 1
        if events[slot] is not None:          # the pinned buffer is being reused: drain it to `out` first
            events[slot].synchronize()
            s0, s1 = spans[slot]
            out[s0:s1] = pinned[slot][: s1 - s0].numpy()
        Z = sample_noise(nb, seq_len, z_dim, device, noise)
        Xh = model.decode(model.refine_latent(model.gen_latent(Z)))
        copy_stream.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(copy_stream):
            pinned[slot][:nb].copy_(Xh, non_blocking=True)
            Xh.record_stream(copy_stream)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        events[slot], spans[slot] = ev, (start, start + nb)
        i += 1
    for slot in ((i & 1), ((i + 1) & 1)):
        if events[slot] is not None:
            events[slot].synchronize()
            s0, s1 = spans[slot]
            out[s0:s1] = pinned[slot][: s1 - s0].numpy()
    return out


def load_model(ckpt_path: Path, x_dim: int, device) -> TimeGAN:
    """Rebuild TimeGAN from a reference-format checkpoint ({"step","model","optG","optD","meta"}, tt:58-61)."""
    state = torch.load(ckpt_path, map_location="cpu")
    meta = state.get("meta", {})
    z_dim, h_dim = int(meta.get("z_dim")), int(meta.get("h_dim"))
    model = TimeGAN(x_dim=x_dim, z_dim=z_dim, hidden_dim=h_dim, num_layers=layers_in_state_dict(state["model"]),
                    dropout=0.2).to(device)
    model.load_state_dict(state["model"])
    model.eval()
    return model


def build_argparser():
    ap = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--runs_dir", type=str, default="./timegan_runs",
                    help="Root containing postureX_with_exo/postureX_no_exo folders")
    ap.add_argument("--real_dir", type=str, default="./preprocessed",
                    help="Where postureX_with_exo.npz etc. live (to get x_dim and fs)")
    ap.add_argument("--out_suffix", type=str, default="synthetic_long.npz",
                    help="Output filename written inside each run folder")
    ap.add_argument("--gen_seconds", type=float, default=None, help="Length to generate in seconds (overrides gen_len).")
    ap.add_argument("--gen_len", type=int, default=None, help="Length to generate in samples. If not set, uses training T.")
    ap.add_argument("--n", type=int, default=None, help="Number of sequences to generate. Default: match real N.")
    ap.add_argument("--prefer_latest", action="store_true", help="Use ckpt_latest.pt instead of ckpt_best.pt if both exist.")
    ap.add_argument("--denorm", action="store_true", help="If set, invert scaling using scale_min/scale_range from real NPZ.")
    ap.add_argument("--chunk", type=int, default=4096, help="windows per device batch (extra of this implementation)")
    return ap


def main(argv=None):
    args = build_argparser().parse_args(argv)
    runs_root, real_root = Path(args.runs_dir), Path(args.real_dir)
    run_dirs = [p for p in sorted(runs_root.iterdir())
                if p.is_dir() and re.match(r"posture\d+_(with_exo|no_exo)$", p.name)]
    if not run_dirs:
        raise SystemExit(f"No run folders found under {runs_root}")
    device = device_autoselect()
    print(f"Using device: {device}")
    for rd in run_dirs:
        posture, cond = infer_posture_cond(rd.name)
        if posture is None:
            continue
        ckpt_best, ckpt_last = rd / "ckpt_best.pt", rd / "ckpt_latest.pt"
        ckpt = ckpt_last if args.prefer_latest and ckpt_last.exists() else (ckpt_best if ckpt_best.exists() else ckpt_last)
        if not ckpt or not ckpt.exists():
            print(f"[SKIP] {rd.name}: no checkpoint found.")
            continue
        real_npz = real_root / f"posture{posture}_{cond}.npz"
        if not real_npz.exists():
            print(f"[SKIP] {rd.name}: real file missing: {real_npz}")
            continue
        real = np.load(real_npz)
        N_real, T_train, C = real["X"].shape
        fs = float(real["fs"]) if "fs" in real.files else 128.0
        model = load_model(ckpt, C, device)
        z_dim = model.embedder.rnn.rnn.hidden_size
        if args.gen_seconds is not None:
            T_out = int(round(args.gen_seconds * fs))
        elif args.gen_len is not None:
            T_out = int(args.gen_len)
        else:
            T_out = int(T_train)
        N_out = int(args.n) if args.n is not None else int(N_real)
        print(f"[{rd.name}] N_out={N_out}  T_out={T_out}  C={C}  z_dim={z_dim}  fs≈{fs:.2f}")
        Xh = generate_windows(model, N_out, T_out, z_dim, device, chunk=args.chunk)
        if args.denorm and "scale_min" in real.files and "scale_range" in real.files:
            mn = real["scale_min"].astype(np.float32)
            rg = real["scale_range"].astype(np.float32)
            Xh = Xh * rg[None, None, :] + mn[None, None, :]
        out_fp = rd / (args.out_suffix if "{" not in args.out_suffix else args.out_suffix.format(T=T_out))
        np.savez_compressed(out_fp, X=Xh)
        print(f"[OK] wrote {out_fp}")


if __name__ == "__main__":
    main()

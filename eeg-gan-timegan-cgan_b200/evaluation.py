"""Evaluation of TimeGAN EEG augmentation on the GPU -- drop-in for timeGAN/evaluation.py.

Same function names, arguments, defaults and return values as the reference (file:line cited per function):
    RNNClassifier / RNNPredictor      evaluation.py:41-59   (post-hoc GRU, hidden 24, reads the last hidden state)
    autocorr_seq                      evaluation.py:63-71
    discriminative_score              evaluation.py:73-101  (20 full-batch Adam epochs, BCE, accuracy + AUC)
    predictive_score                  evaluation.py:103-118 (50 full-batch Adam epochs, MSE, RMSE + R^2; TSTR / TRTS)
    statistical_similarity            evaluation.py:120-139 (Welch PSD, autocorrelation score, channel correlation)
    load_posture_pairs, main          evaluation.py:141-276 (CSV outputs; the PCA / t-SNE figures are not produced:
                                                             plotting is outside the hot-path scope)

What runs where.  The two post-hoc networks train through this package's persistent GRU kernels (forward with
last_only, BPTT, fused weight gradients) and FusedAdam; the autocorrelation score -- N*C*96 np.corrcoef calls in the
reference, minutes per posture -- is one launch of csrc/eval_stats.cu; the Welch periodogram is batched on the device
(segmenting, detrend, Hann window, rFFT via torch.fft) and the channel-correlation matrix is an fp64 Gram on the
device.  Index logic that decides WHICH windows are used (class balancing, stratified split) is the reference's own
numpy / scikit-learn calls, so both sides score the same windows.  No CPU fallback: everything needs CUDA.
"""
import argparse
import ctypes as C
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import lib, check, ptr, stream_ptr, require_cuda
from .optim import FusedAdam
from .timegan_model import FusedGRU


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("timegan_b200.evaluation needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# -------------------------- Models (ev:41-59) --------------------------

class _LastStateRNN(nn.Module):
    def __init__(self, input_dim, hidden_dim, num_layers, out_dim):
        super().__init__()
        # same construction order and initialisers as nn.GRU + nn.Linear => same weights for the same seed
        self.rnn = FusedGRU(input_dim, hidden_dim, num_layers=num_layers, dropout=0.0, batch_first=True)
        self.out = nn.Linear(hidden_dim, out_dim)

    def _head(self, x):
        hn, _ = self.rnn(x, last_only=True)          # hn[-1] of nn.GRU == y[:, -1, :] of the top layer
        return ops.linear(hn, self.out.weight, self.out.bias)


class RNNClassifier(_LastStateRNN):
    def __init__(self, input_dim, hidden_dim=24, num_layers=1):
        super().__init__(input_dim, hidden_dim, num_layers, 1)

    def forward(self, x):
        return torch.sigmoid(self._head(x))


class RNNPredictor(_LastStateRNN):
    def __init__(self, input_dim, hidden_dim=24, num_layers=1, output_dim=None):
        super().__init__(input_dim, hidden_dim, num_layers, output_dim or input_dim)

    def forward(self, x):
        return self._head(x)


# -------------------------- Helpers --------------------------

def acf_scores(x: torch.Tensor, maxlag: int) -> torch.Tensor:
    """autocorr_seq (ev:63-71) for every window and channel at once: x (N,T,C) fp32 CUDA -> (N,C) fp64."""
    require_cuda(x, "acf_scores input")
    x = x.contiguous()
    N, T, Cc = x.shape
    out = torch.empty(N, Cc, dtype=torch.float64, device=x.device)
    check(lib.tg_acf_score(stream_ptr(), ptr(x), N, T, Cc, int(maxlag), ptr(out)), "tg_acf_score")
    return out


def autocorr_seq(x, maxlag, device=None):
    """ev:63-71 for one series (numpy in, float out)."""
    t = torch.as_tensor(np.asarray(x, dtype=np.float32)).reshape(1, -1, 1).to(_device(device))
    return float(acf_scores(t, maxlag)[0, 0].item())


def _split_for_discriminator(real, fake, seed=0):
    """ev:75-80 verbatim index logic (numpy RandomState permutations + sklearn's stratified split)."""
    from sklearn.model_selection import train_test_split
    n = min(len(real), len(fake))
    idx_r = np.random.RandomState(seed).permutation(len(real))[:n]
    idx_f = np.random.RandomState(seed + 1).permutation(len(fake))[:n]
    X = np.concatenate([real[idx_r], fake[idx_f]], axis=0)
    y = np.concatenate([np.ones(n), np.zeros(n)], axis=0)
    return train_test_split(X, y, test_size=0.3, stratify=y, random_state=seed)


def _fit(model, Xt, yt, loss_fn, epochs, lr):
    opt = FusedAdam(model.parameters(), lr=lr)              # torch.optim.Adam defaults (ev:82,105), no clipping
    for _ in range(epochs):
        opt.zero_grad()
        loss = loss_fn(model(Xt), yt)
        loss.backward()
        opt.step()
    return model


def discriminative_score(real, fake, epochs=20, lr=1e-3, hidden=24, seed=0, device=None, return_probs=False):
    """ev:73-101: post-hoc real-vs-synthetic classifier; returns (accuracy, AUC) on the held-out 30 %."""
    from sklearn.metrics import accuracy_score, roc_auc_score
    dev = _device(device)
    Xtr, Xte, ytr, yte = _split_for_discriminator(real, fake, seed)
    clf = RNNClassifier(Xtr.shape[-1], hidden).to(dev)
    Xt = torch.tensor(Xtr, dtype=torch.float32, device=dev)
    yt = torch.tensor(ytr, dtype=torch.float32, device=dev).unsqueeze(1)
    _fit(clf, Xt, yt, nn.functional.binary_cross_entropy, epochs, lr)
    with torch.no_grad():
        p = clf(torch.tensor(Xte, dtype=torch.float32, device=dev)).cpu().numpy().flatten()
    yhat = (p >= 0.5).astype(int)
    acc = accuracy_score(yte, yhat)
    try:
        auc = roc_auc_score(yte, p)
    except ValueError:
        auc = np.nan
    return (acc, auc, p) if return_probs else (acc, auc)


def predictive_score(X_train, y_train, X_test, y_test, epochs=50, lr=1e-3, hidden=24, device=None):
    """ev:103-118: train a next-step predictor on one domain, score RMSE / R^2 on the other."""
    from sklearn.metrics import mean_squared_error, r2_score
    dev = _device(device)
    model = RNNPredictor(X_train.shape[-1], hidden).to(dev)
    Xt = torch.tensor(np.ascontiguousarray(X_train), dtype=torch.float32, device=dev)
    yt = torch.tensor(np.ascontiguousarray(y_train), dtype=torch.float32, device=dev)
    _fit(model, Xt, yt, nn.functional.mse_loss, epochs, lr)
    with torch.no_grad():
        yhat = model(torch.tensor(np.ascontiguousarray(X_test), dtype=torch.float32, device=dev)).cpu().numpy()
    rmse = np.sqrt(mean_squared_error(y_test, yhat))
    r2 = r2_score(y_test, yhat, multioutput="uniform_average")
    return rmse, r2


def welch_psd(x: torch.Tensor, fs: float = 128.0, nperseg: int = 256) -> torch.Tensor:
    """scipy.signal.welch(x, fs, axis=1, nperseg) with its defaults (Hann window, 50 % overlap, constant detrend,
    density scaling, one-sided, mean over segments), batched on the device: (N,T,C) fp32 -> (N, nperseg//2+1, C) fp64."""
    N, T, Cc = x.shape
    nper = min(nperseg, T)
    step = nper - nper // 2
    seg = x.to(torch.float64).permute(0, 2, 1).unfold(2, nper, step)          # (N, C, S, nper)
    seg = seg - seg.mean(dim=-1, keepdim=True)
    win = torch.hann_window(nper, periodic=True, dtype=torch.float64, device=x.device)
    spec = torch.fft.rfft(seg * win, dim=-1)
    p = (spec.real ** 2 + spec.imag ** 2) / (fs * (win * win).sum())
    if nper % 2 == 0:
        p[..., 1:-1] *= 2.0
    else:
        p[..., 1:] *= 2.0
    return p.mean(dim=2).permute(0, 2, 1)                                     # (N, F, C)


def _corrcoef_cols(flat: torch.Tensor) -> torch.Tensor:
    """np.corrcoef(flat, rowvar=False) in fp64 on the device."""
    f = flat.to(torch.float64)
    f = f - f.mean(dim=0, keepdim=True)
    cov = f.t() @ f / (f.shape[0] - 1)
    d = torch.sqrt(torch.diagonal(cov))
    return cov / (d[:, None] * d[None, :])


def statistical_similarity(real, fake, fs=128.0, device=None):
    """ev:120-139: (psd_diff, acf_diff, coh_diff)."""
    dev = _device(device)
    r = torch.as_tensor(np.asarray(real, dtype=np.float32)).to(dev)
    f = torch.as_tensor(np.asarray(fake, dtype=np.float32)).to(dev)
    psd_diff = float((welch_psd(r, fs).mean(dim=0) - welch_psd(f, fs).mean(dim=0)).abs().mean().item())
    maxlag = int(0.75 * fs)
    acf_r = acf_scores(r, maxlag).mean(dim=0)          # per channel: mean over windows
    acf_f = acf_scores(f, maxlag).mean(dim=0)
    acf_diff = float((acf_r - acf_f).abs().mean().item())
    corr_r = _corrcoef_cols(r.reshape(-1, r.shape[-1]))
    corr_f = _corrcoef_cols(f.reshape(-1, f.shape[-1]))
    coh_diff = float((corr_r - corr_f).abs().mean().item())
    return psd_diff, acf_diff, coh_diff


def load_posture_pairs(real_dir: Path, synth_dir: Path):
    """ev:141-166: posture -> (real, fake), conditions concatenated and balanced within each condition."""
    pairs = {}
    for p in range(1, 10):
        real_list, fake_list = [], []
        for cond in ["with_exo", "no_exo"]:
            rfp = real_dir / f"posture{p}_{cond}.npz"
            sfp = synth_dir / f"posture{p}_{cond}" / "synthetic.npz"
            if rfp.exists() and sfp.exists():
                r = np.load(rfp)["X"].astype(np.float32)
                f = np.load(sfp)["X"].astype(np.float32)
                m = min(len(r), len(f))
                if m > 0:
                    real_list.append(r[:m])
                    fake_list.append(f[:m])
        if real_list and fake_list:
            pairs[p] = (np.concatenate(real_list, axis=0), np.concatenate(fake_list, axis=0))
    return pairs


def evaluate_pair(real, fake, fs=128.0, device=None):
    """The metric block main() computes per posture and globally (ev:190-216, 224-238)."""
    acc, auc = discriminative_score(real, fake, device=device)
    Xr_in, yr = real[:, :-1, :], real[:, -1, :]
    Xf_in, yf = fake[:, :-1, :], fake[:, -1, :]
    rmse_tstr, r2_tstr = predictive_score(Xf_in, yf, Xr_in, yr, device=device)
    rmse_trts, r2_trts = predictive_score(Xr_in, yr, Xf_in, yf, device=device)
    psd_diff, acf_diff, coh_diff = statistical_similarity(real, fake, fs=fs, device=device)
    return {"disc_acc": acc, "disc_auc": auc, "rmse_tstr": rmse_tstr, "r2_tstr": r2_tstr, "rmse_trts": rmse_trts,
            "r2_trts": r2_trts, "psd_diff": psd_diff, "acf_diff": acf_diff, "coh_diff": coh_diff,
            "n_real": len(real), "n_fake": len(fake), "seq_len": real.shape[1], "n_ch": real.shape[2]}


def main(argv=None):
    """ev:170-276 without the figures: writes metrics_per_posture.csv and metrics_global.csv."""
    import pandas as pd
    ap = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--real_dir", type=str, default="./preprocessed")
    ap.add_argument("--synth_dir", type=str, default="./timegan_runs")
    ap.add_argument("--out", type=str, default="./eval_out")
    ap.add_argument("--fs", type=float, default=128.0)
    args = ap.parse_args(argv)
    np.random.seed(0)
    torch.manual_seed(0)
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    pairs = load_posture_pairs(Path(args.real_dir), Path(args.synth_dir))
    if not pairs:
        raise SystemExit("No matching posture pairs found. Make sure synthetic.npz exists for each trained model.")
    rows, all_real, all_fake = [], [], []
    for posture in sorted(pairs.keys()):
        real, fake = pairs[posture]
        rows.append(dict({"posture": posture}, **evaluate_pair(real, fake, fs=args.fs)))
        all_real.append(real)
        all_fake.append(fake)
    pd.DataFrame(rows).sort_values("posture").to_csv(out / "metrics_per_posture.csv", index=False)
    print(f"Wrote {out / 'metrics_per_posture.csv'}")
    R, F = np.concatenate(all_real, axis=0), np.concatenate(all_fake, axis=0)
    pd.DataFrame([evaluate_pair(R, F, fs=args.fs)]).to_csv(out / "metrics_global.csv", index=False)
    print(f"Wrote {out / 'metrics_global.csv'}")
    print("PCA / t-SNE figures of the reference (ev:240-273) are not produced by the GPU evaluation.")


if __name__ == "__main__":
    main()

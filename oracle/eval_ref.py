"""CPU restatement of the reference's evaluation metrics (timeGAN/evaluation.py) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and tools' cpu-baseline legs may import this.  The reference module cannot be
imported on the GPU box (it is absent there, and it imports matplotlib at top level, ev:30), so its metric functions
are restated here call for call; `oracle/make_golden_eval.py` pins this restatement against the UNMODIFIED
reference (run with a stub matplotlib) and commits tests/golden/eval_small.npz.

    RNNClassifier / RNNPredictor      ev:41-59
    autocorr_seq                      ev:63-71
    discriminative_score              ev:73-101
    predictive_score                  ev:103-118
    statistical_similarity            ev:120-139
"""
import numpy as np
import scipy.signal as sig
import torch
import torch.nn as nn
import torch.optim as optim
from sklearn.metrics import accuracy_score, mean_squared_error, r2_score, roc_auc_score
from sklearn.model_selection import train_test_split


class RNNClassifier(nn.Module):                      # ev:41-48
    def __init__(self, input_dim, hidden_dim=24, num_layers=1):
        super().__init__()
        self.rnn = nn.GRU(input_dim, hidden_dim, num_layers, batch_first=True)
        self.out = nn.Linear(hidden_dim, 1)

    def forward(self, x):
        _, hn = self.rnn(x)
        return torch.sigmoid(self.out(hn[-1]))


class RNNPredictor(nn.Module):                       # ev:50-59
    def __init__(self, input_dim, hidden_dim=24, num_layers=1, output_dim=None):
        super().__init__()
        output_dim = output_dim or input_dim
        self.rnn = nn.GRU(input_dim, hidden_dim, num_layers, batch_first=True)
        self.out = nn.Linear(hidden_dim, output_dim)

    def forward(self, x):
        _, hn = self.rnn(x)
        return self.out(hn[-1])


def autocorr_seq(x, maxlag):                         # ev:63-71
    if np.std(x) < 1e-8:
        return 0.0
    vals = []
    for lag in range(1, maxlag + 1):
        if lag >= len(x):
            break
        vals.append(np.corrcoef(x[:-lag], x[lag:])[0, 1])
    return float(np.mean(vals)) if vals else 0.0


def split_for_discriminator(real, fake, seed=0):     # ev:75-80 (index logic shared with the GPU implementation)
    n = min(len(real), len(fake))
    idx_r = np.random.RandomState(seed).permutation(len(real))[:n]
    idx_f = np.random.RandomState(seed + 1).permutation(len(fake))[:n]
    X = np.concatenate([real[idx_r], fake[idx_f]], axis=0)
    y = np.concatenate([np.ones(n), np.zeros(n)], axis=0)
    return train_test_split(X, y, test_size=0.3, stratify=y, random_state=seed)


def discriminative_score(real, fake, epochs=20, lr=1e-3, hidden=24, seed=0, return_probs=False):   # ev:73-101
    Xtr, Xte, ytr, yte = split_for_discriminator(real, fake, seed)
    clf = RNNClassifier(Xtr.shape[-1], hidden)
    opt = optim.Adam(clf.parameters(), lr=lr)
    lossf = nn.BCELoss()
    Xt = torch.tensor(Xtr, dtype=torch.float32)
    yt = torch.tensor(ytr, dtype=torch.float32).unsqueeze(1)
    for _ in range(epochs):
        opt.zero_grad()
        p = clf(Xt)
        loss = lossf(p, yt)
        loss.backward()
        opt.step()
    with torch.no_grad():
        p = clf(torch.tensor(Xte, dtype=torch.float32)).numpy().flatten()
    yhat = (p >= 0.5).astype(int)
    acc = accuracy_score(yte, yhat)
    try:
        auc = roc_auc_score(yte, p)
    except ValueError:
        auc = np.nan
    return (acc, auc, p) if return_probs else (acc, auc)


def predictive_score(X_train, y_train, X_test, y_test, epochs=50, lr=1e-3, hidden=24):              # ev:103-118
    model = RNNPredictor(X_train.shape[-1], hidden)
    opt = optim.Adam(model.parameters(), lr=lr)
    lossf = nn.MSELoss()
    Xt = torch.tensor(X_train, dtype=torch.float32)
    yt = torch.tensor(y_train, dtype=torch.float32)
    for _ in range(epochs):
        opt.zero_grad()
        pred = model(Xt)
        loss = lossf(pred, yt)
        loss.backward()
        opt.step()
    with torch.no_grad():
        yhat = model(torch.tensor(X_test, dtype=torch.float32)).numpy()
    rmse = np.sqrt(mean_squared_error(y_test, yhat))
    r2 = r2_score(y_test, yhat, multioutput="uniform_average")
    return rmse, r2


def statistical_similarity(real, fake, fs=128.0):     # ev:120-139
    fr, psd_r = sig.welch(real, fs=fs, axis=1, nperseg=256)
    ff, psd_f = sig.welch(fake, fs=fs, axis=1, nperseg=256)
    psd_diff = float(np.mean(np.abs(psd_r.mean(axis=0) - psd_f.mean(axis=0))))
    maxlag = int(0.75 * fs)
    acf_r, acf_f = [], []
    for ch in range(real.shape[-1]):
        acf_r.append(np.mean([autocorr_seq(seq[:, ch], maxlag) for seq in real]))
        acf_f.append(np.mean([autocorr_seq(seq[:, ch], maxlag) for seq in fake]))
    acf_diff = float(np.mean(np.abs(np.array(acf_r) - np.array(acf_f))))
    r_flat = real.reshape(-1, real.shape[-1])
    f_flat = fake.reshape(-1, fake.shape[-1])
    corr_r = np.corrcoef(r_flat, rowvar=False)
    corr_f = np.corrcoef(f_flat, rowvar=False)
    coh_diff = float(np.mean(np.abs(corr_r - corr_f)))
    return psd_diff, acf_diff, coh_diff

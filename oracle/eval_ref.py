"""CPU restatement of the reference's evaluation metrics (timeGAN/evaluation.py) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and the cpu-baseline legs of the bench tools may import this.  The reference
module cannot be imported on the GPU box (it is absent there, and it imports matplotlib at top level, ev:30), so
the arithmetic of its metric functions is restated here with the same library calls in the same order (which is
what makes the results bit-identical: same initialisers drawing from the same generator, same float64 numpy
paths).  `oracle/make_golden_eval.py` pins this restatement against the UNMODIFIED reference, run with a stub
matplotlib, and commits tests/golden/eval_small.npz.

    post-hoc networks                 ev:41-59    one-layer GRU(hidden 24) + Linear on the LAST hidden state
    autocorr_seq                      ev:63-71    mean over lags of np.corrcoef(x[:-lag], x[lag:])
    discriminative_score              ev:73-101   balanced classes, 70/30 stratified split, 20 Adam epochs, BCE
    predictive_score                  ev:103-118  50 Adam epochs, MSE; RMSE and R^2 on the other domain
    statistical_similarity            ev:120-139  Welch PSD gap, autocorrelation gap, channel-correlation gap
"""
import numpy as np
import scipy.signal
import sklearn.metrics as skm
import torch
from sklearn.model_selection import train_test_split


class _LastState(torch.nn.Module):
    """GRU then Linear, in that construction order (the order fixes which random numbers each one receives)."""

    def __init__(self, n_in, n_hidden, n_out, squash):
        super().__init__()
        self.rnn = torch.nn.GRU(n_in, n_hidden, 1, batch_first=True)
        self.out = torch.nn.Linear(n_hidden, n_out)
        self.squash = squash

    def forward(self, x):
        last = self.rnn(x)[1][-1]
        z = self.out(last)
        return torch.sigmoid(z) if self.squash else z


def RNNClassifier(input_dim, hidden_dim=24):
    return _LastState(input_dim, hidden_dim, 1, True)


def RNNPredictor(input_dim, hidden_dim=24, output_dim=None):
    return _LastState(input_dim, hidden_dim, output_dim or input_dim, False)


def _full_batch_fit(net, inputs, targets, criterion, epochs, lr):
    """`epochs` Adam steps (torch defaults) on the whole training set at once."""
    adam = torch.optim.Adam(net.parameters(), lr=lr)
    x = torch.tensor(inputs, dtype=torch.float32)
    y = torch.tensor(targets, dtype=torch.float32)
    for _ in range(epochs):
        adam.zero_grad()
        criterion(net(x), y).backward()
        adam.step()
    return net


def _predict(net, inputs):
    with torch.no_grad():
        return net(torch.tensor(inputs, dtype=torch.float32)).numpy()


def autocorr_seq(x, maxlag):
    if np.std(x) < 1e-8:
        return 0.0
    lags = range(1, min(maxlag, len(x) - 1) + 1)
    r = [np.corrcoef(x[:-k], x[k:])[0, 1] for k in lags]
    return float(np.mean(r)) if r else 0.0


def split_for_discriminator(real, fake, seed=0):
    """Which windows the classifier trains / is scored on: class balancing by two seeded permutations, then a
    stratified 70/30 split."""
    n = min(len(real), len(fake))
    pick_r = np.random.RandomState(seed).permutation(len(real))[:n]
    pick_f = np.random.RandomState(seed + 1).permutation(len(fake))[:n]
    windows = np.concatenate([real[pick_r], fake[pick_f]], axis=0)
    labels = np.concatenate([np.ones(n), np.zeros(n)], axis=0)
    return train_test_split(windows, labels, test_size=0.3, stratify=labels, random_state=seed)


def discriminative_score(real, fake, epochs=20, lr=1e-3, hidden=24, seed=0, return_probs=False):
    x_fit, x_held, y_fit, y_held = split_for_discriminator(real, fake, seed)
    net = RNNClassifier(x_fit.shape[-1], hidden)
    _full_batch_fit(net, x_fit, y_fit[:, None], torch.nn.BCELoss(), epochs, lr)
    prob = _predict(net, x_held).flatten()
    acc = skm.accuracy_score(y_held, (prob >= 0.5).astype(int))
    try:
        auc = skm.roc_auc_score(y_held, prob)
    except ValueError:
        auc = np.nan
    return (acc, auc, prob) if return_probs else (acc, auc)


def predictive_score(X_train, y_train, X_test, y_test, epochs=50, lr=1e-3, hidden=24):
    net = RNNPredictor(X_train.shape[-1], hidden)
    _full_batch_fit(net, X_train, y_train, torch.nn.MSELoss(), epochs, lr)
    guess = _predict(net, X_test)
    return (np.sqrt(skm.mean_squared_error(y_test, guess)),
            skm.r2_score(y_test, guess, multioutput="uniform_average"))


def _mean_psd(x, fs):
    return scipy.signal.welch(x, fs=fs, axis=1, nperseg=256)[1].mean(axis=0)


def _channel_acf(x, maxlag):
    return np.array([np.mean([autocorr_seq(w[:, c], maxlag) for w in x]) for c in range(x.shape[-1])])


def _channel_corr(x):
    return np.corrcoef(x.reshape(-1, x.shape[-1]), rowvar=False)


def statistical_similarity(real, fake, fs=128.0):
    gap = lambda a, b: float(np.mean(np.abs(a - b)))
    maxlag = int(0.75 * fs)
    return (gap(_mean_psd(real, fs), _mean_psd(fake, fs)),
            gap(_channel_acf(real, maxlag), _channel_acf(fake, maxlag)),
            gap(_channel_corr(real), _channel_corr(fake)))

"""Loading of tests/golden/steps_*.npz (outputs of the unmodified reference, see oracle/make_golden_steps.py)."""
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"

# the hyper-parameters the fixtures were generated with (oracle/make_golden_steps.py HP)
HP = dict(lr_g=1e-3, lr_d=2e-4, betas=(0.5, 0.9), alpha_sup=5.0, beta_rec=0.2, label_smooth=0.2, inst_noise=0.3,
          clip=0.5, r1_gamma=1.0, target=0.5 * (0.45 + 0.60), band=0.60 - 0.45, gamma_cov=0.05, gamma_acf=0.05,
          acf_max_lag=64)
CASES = ("tiny", "c1", "refdefault")


class StepFixture:
    def __init__(self, name):
        z = np.load(GOLDEN / f"steps_{name}.npz")
        self.x_dim, self.z_dim, self.h_dim, self.layers, self.B, self.T, self.seed = [int(v) for v in z["dims"]]
        self.x = torch.from_numpy(z["x"])
        self.init = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init/")}
        self.final = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("final/")}
        self.grads = {s: {k.split("/", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"grad_{s}/")}
                      for s in ("ae", "sup", "d", "g")}
        self.ae_loss, self.sup_loss = float(z["ae_loss"]), float(z["sup_loss"])
        self.d_out, self.g_out = z["d_out"].tolist(), z["g_out"].tolist()

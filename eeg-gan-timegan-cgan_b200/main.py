"""B200-native drop-in for timeGAN/main.py (mn:13-79): JSON/YAML config -> train_single_npz for every
posture*_*.npz under data_dir.  Same keys and defaults as the reference (timegan_config.json loads
unchanged); optional extra keys: z_dim, hidden_dim, proj_dtype, noise, log_every."""
import argparse
import json
from pathlib import Path
from typing import Any, Dict

from . import dist as _dist
from . import train_timegan as tt


def load_config(path: Path) -> Dict[str, Any]:
    if not path.exists():
        raise SystemExit(f"Config file not found: {path}")
    if path.suffix.lower() in {".yaml", ".yml"}:
        try:
            import yaml
        except Exception as e:  # pragma: no cover
            raise SystemExit("YAML config requested but PyYAML not installed. Install with `pip install pyyaml` "
                             "or use JSON.") from e
        with open(path, "r", encoding="utf-8") as f:
            return yaml.safe_load(f)
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


def kwargs_from_config(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """Keyword arguments of train_single_npz from a config dict (mn:51-76 defaults)."""
    g = cfg.get
    kw = dict(
        batch_size=int(g("batch_size", 64)), ae_epochs=int(g("ae_epochs", 120)), sup_epochs=int(g("sup_epochs", 150)),
        gan_steps=int(g("gan_steps", 8000)), lr_g=float(g("lr_g", 1e-3)), lr_d=float(g("lr_d", 2e-4)),
        betas=(float(g("beta1", 0.5)), float(g("beta2", 0.9))), alpha_sup=float(g("alpha_sup", 5.0)),
        beta_rec=float(g("beta_rec", 0.2)), label_smooth=float(g("label_smooth", 0.2)),
        inst_noise_start=float(g("inst_noise_start", 0.3)), inst_noise_end=float(g("inst_noise_end", 0.1)),
        grad_clip=float(g("grad_clip", 0.5)), layers=int(g("layers", 1)), dropout=float(g("dropout", 0.2)),
        seed=int(g("seed", 42)), r1_gamma=float(g("r1_gamma", 1.0)), d_min_acc=float(g("d_min_acc", 0.45)),
        d_max_acc=float(g("d_max_acc", 0.60)), gamma_cov=float(g("gamma_cov", 0.05)),
        gamma_acf=float(g("gamma_acf", 0.05)), acf_max_lag=int(g("acf_max_lag", 64)))
    for k, cast in (("z_dim", int), ("hidden_dim", int), ("proj_dtype", str), ("noise", str), ("log_every", int),
                    ("graph", bool)):
        if g(k) is not None:
            kw[k] = cast(g(k))
    return kw


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default="timegan_config.json", help="Path to config file (JSON or YAML)")
    args = ap.parse_args(argv)
    cfg = load_config(Path(args.config))
    data_dir = Path(cfg.get("data_dir", "./preprocessed"))
    out_root = Path(cfg.get("out_dir", "./timegan_runs"))
    out_root.mkdir(parents=True, exist_ok=True)
    files = sorted(data_dir.glob("posture*_*.npz"))
    if not files:
        raise SystemExit(f"No NPZs found in {data_dir}. Did you run preprocessing?")
    _dist.init()
    device = tt.device_autoselect()
    print(f"Using device: {device}")
    print(f"Found {len(files)} datasets → training {len(files)} models.")
    kw = kwargs_from_config(cfg)
    for fp in files:
        run_dir = out_root / fp.stem
        print(f"\n=== Training {fp.name} → {run_dir} ===")
        tt.train_single_npz(npz_path=fp, out_dir=run_dir, device=device, **kw)
    _dist.shutdown()
    print("\nAll models trained. Checkpoints, logs, and synthetic data are under:", out_root)


if __name__ == "__main__":
    main()

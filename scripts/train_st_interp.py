"""Spatio-temporal DeepKriging training driver on the B200-native `stnf` (same CLI and yaml schema as upstream
scripts/train_st_interp.py; artefact names `results.json`, `training_history.csv`, `model_best.pt`,
`model_final.pt`, `predictions.npz`, `basis_info.npz` are kept; plotting is left to the offline tools).

    python scripts/train_st_interp.py --config configs/config_st_interp.yaml [--data_file ...] [--n_experiments N]
    torchrun --nproc-per-node 8 scripts/train_st_interp.py --config ...   # independent experiments packed per GPU
    torchrun --nproc-per-node 8 scripts/train_st_interp.py --config ... --data_parallel
                                   # ONE model trained by all GPUs: each global batch split over the ranks, one gradient
                                   # exchange per step over NVLink (st_dadk_b200/peer.py), rank 0 writes the artefacts

The per-batch Python of upstream (list-of-dict dataset, collate, per-step .to(device)/.item()) is replaced by the
device-resident ObservationTable + st_dadk_b200.trainer.fit; observation / split masks are drawn with the same numpy
global-RNG calls as upstream so that a seeded run sees the same samples.
"""
import argparse
import json
import os
import sys
import time
from datetime import datetime
from pathlib import Path

import numpy as np
import torch
import yaml

sys.path.append(str(Path(__file__).resolve().parent.parent))

from stnf.models.st_interp import STInterpMLP, create_model   # noqa: E402
from stnf.dataio.kaust_loader import load_kaust_csv_single, ObservationTable   # noqa: E402
from stnf.utils import set_seed, ModelEMA   # noqa: E402,F401


# ----------------------------------------------------------------------------- losses / penalties (host-side API)
def quantile_loss(y_pred, y_true, quantile):
    """Pinball loss, mean over elements (upstream :37-50)."""
    e = y_true - y_pred
    return torch.maximum((quantile - 1) * e, quantile * e).mean()


def non_crossing_penalty(y_pred_multi_q: torch.Tensor, reduction: str = "mean", power: int = 1):
    """sum_k relu(q_k - q_{k+1})^power per sample (upstream :53-85)."""
    if y_pred_multi_q.dim() != 2 or y_pred_multi_q.shape[1] < 2:
        return torch.tensor(0.0, device=y_pred_multi_q.device)
    if power not in (1, 2):
        raise ValueError(f"Unsupported power={power}; use 1 or 2.")
    v = torch.relu(y_pred_multi_q[:, :-1] - y_pred_multi_q[:, 1:])
    per = (v ** 2 if power == 2 else v).sum(dim=1)
    if reduction == "mean":
        return per.mean()
    if reduction == "sum":
        return per.sum()
    raise ValueError(f"Unsupported reduction='{reduction}'; use 'mean' or 'sum'.")


def compute_p_nc_delta_penalty(delta_params: list) -> torch.Tensor:
    """P_nc(delta) = sum_{k>=2} [delta_k0 - max(delta_k0, sum_j max(0, -delta_kj))]  (Eq. 3.10, upstream :88-150)."""
    if not delta_params or len(delta_params) < 2:
        dev = delta_params[0].device if delta_params else torch.device("cpu")
        return torch.tensor(0.0, device=dev)
    total = torch.tensor(0.0, device=delta_params[0].device)
    for d in delta_params[1:]:
        total = total + d[0] - torch.max(d[0], torch.clamp(-d[1:], min=0.0).sum())
    return total


def check_loss_numpy(y_pred, y_true, quantile):
    e = y_true - y_pred
    return np.mean(np.maximum((quantile - 1) * e, quantile * e))


def compute_crps(predictions_dict, y_true, weights=None):
    """CRPS = 2 * sum_k w_k rho_{tau_k}(y - Q_{tau_k}), uniform weights by default (Eq. 4.6, upstream :169-223)."""
    qs = sorted(predictions_dict.keys())
    if not qs:
        raise ValueError("predictions_dict cannot be empty")
    if len(qs) == 1:
        return 2.0 * check_loss_numpy(predictions_dict[qs[0]], y_true, qs[0])
    if weights is None:
        w = np.full(len(qs), 1.0 / len(qs))
    else:
        w = np.asarray(weights, dtype=float)
        if len(w) != len(qs):
            raise ValueError(f"weights length ({len(w)}) must match number of quantiles ({len(qs)})")
        w = w / w.sum()
    return 2.0 * float(sum(wk * check_loss_numpy(predictions_dict[q], y_true, q) for wk, q in zip(w, qs)))


def compute_crps_multi_quantile(preds, y_true, quantile_levels, weights=None):
    y = y_true.flatten() if y_true.ndim > 1 else y_true
    return compute_crps({q: preds[:, i] for i, q in enumerate(quantile_levels)}, y, weights=weights)


# ----------------------------------------------------------------------------- observation sampling (numpy RNG parity)
def create_spatial_obs_prob_fn(pattern="uniform", intensity=1.0):
    """'corner': p(s) ~ 1 / (1 + intensity |s|^2)^2 (upstream :251-279).  Returns a VECTORISED weight function
    (coords (S,2) float32 -> weights float32) computing the same FP32 values as upstream's per-site calls."""
    if pattern == "uniform" or pattern is None:
        return None
    if pattern == "corner":
        return lambda c: 1.0 / (1.0 + intensity * (c[..., 0] ** 2 + c[..., 1] ** 2)) ** 2
    raise ValueError(f"Unknown pattern: {pattern}")


def sample_observations(z_data, coords, obs_method="site-wise", obs_ratio=0.5, obs_prob_fn=None, seed=None):
    """Observed (t, site) mask (upstream :282-342): same numpy global-RNG calls in the same order."""
    if seed is not None:
        np.random.seed(seed)
    T, S = z_data.shape
    if obs_prob_fn is not None:
        w = np.asarray(obs_prob_fn(coords))
        probs = np.clip(w / w.mean() * obs_ratio, 0, 1)
    else:
        probs = np.ones(S) * obs_ratio
    mask = np.zeros((T, S), dtype=bool)
    if obs_method == "site-wise":
        sites = np.random.choice(S, size=int(S * obs_ratio), replace=False, p=probs / probs.sum())
        mask[:, sites] = True
        return mask, sites
    if obs_method == "random":
        mask = np.random.rand(T, S) < probs[np.newaxis, :].repeat(T, axis=0)
        return mask, np.where(mask.any(axis=0))[0]
    raise ValueError(f"Unknown obs_method: {obs_method}")


def split_train_valid(obs_mask, obs_sites, split_method="site-wise", train_ratio=0.8, seed=None):
    """Train / validation masks from the observed mask (upstream :345-410)."""
    if seed is not None:
        np.random.seed(seed)
    train, valid = np.zeros_like(obs_mask), np.zeros_like(obs_mask)
    if split_method == "site-wise":
        order = obs_sites.copy()
        np.random.shuffle(order)
        k = int(len(obs_sites) * train_ratio)
        train[:, order[:k]] = obs_mask[:, order[:k]]
        valid[:, order[k:]] = obs_mask[:, order[k:]]
        return train, valid
    if split_method == "random":
        pairs = np.argwhere(obs_mask)
        order = np.random.permutation(len(pairs))
        k = int(len(pairs) * train_ratio)
        tr, va = pairs[order[:k]], pairs[order[k:]]
        train[tr[:, 0], tr[:, 1]] = True
        valid[va[:, 0], va[:, 1]] = True
        return train, valid
    raise ValueError(f"Unknown split_method: {split_method}")


def create_dataset_from_mask(z_data, coords, mask, p_covariates=0) -> ObservationTable:
    """Struct-of-arrays replacement of upstream's list of per-sample dicts (:413-450): same samples, same order."""
    return ObservationTable.from_mask(z_data, coords, mask, p_covariates)


def auto_batch_size(batch_size: int, n_train: int, min_batches: int = 10) -> int:
    """Halve until there are at least `min_batches` batches per epoch (upstream :2276-2288)."""
    while n_train / batch_size < min_batches and batch_size > 1:
        batch_size //= 2
    return batch_size


# ----------------------------------------------------------------------------- training / evaluation
def train_model(model, train_data, val_data, config, device, output_dir):
    """Upstream signature (:463); `train_data` / `val_data` are ObservationTables (or anything with the fields
    coords, t, y).  The batch size is config['batch_size'] as already adjusted by the caller."""
    from st_dadk_b200.trainer import fit
    bs = int(config.get("_effective_batch_size", config.get("batch_size", 256)))
    return fit(model, train_data, val_data, config, device, output_dir=output_dir, batch_size=bs,
               use_cuda_graph=bool(config.get("cuda_graph", True)))


@torch.no_grad()
def evaluate_model(model, data, device, config=None):
    """MSE / MAE / RMSE (median quantile for multi-quantile) + check loss / CRPS (upstream :884-961)."""
    from st_dadk_b200.predict import Predictor
    model.eval()
    tab = data.to(device)
    preds, _ = Predictor(model).points(tab.coords, tab.t, tab.X)
    preds, trues = preds.cpu().numpy(), tab.y.cpu().numpy()[:, None]
    rtype = config.get("regression_type", "mean") if config is not None else "mean"
    levels = config.get("quantile_levels", [0.1, 0.5, 0.9]) if config is not None else None
    pm = preds[:, len(levels) // 2:len(levels) // 2 + 1] if rtype == "multi-quantile" else preds
    mse = float(np.mean((pm - trues) ** 2))
    out = {"mse": mse, "mae": float(np.mean(np.abs(pm - trues))), "rmse": float(np.sqrt(mse))}
    if rtype == "quantile" and config is not None and "current_quantile" in config:
        out["check_loss"] = float(check_loss_numpy(preds, trues, config["current_quantile"]))
    if rtype == "multi-quantile":
        out["crps"] = float(compute_crps_multi_quantile(preds, trues, levels))
        cl = [check_loss_numpy(preds[:, i:i + 1], trues, q) for i, q in enumerate(levels)]
        out["mean_check_loss"] = out["check_loss"] = float(np.mean(cl))
    return out


@torch.no_grad()
def predict_field(model, coords, T, device, config=None):
    """(T, S) prediction field of upstream's plot_spatial_mse loop (:1228-1248), in one sharded-by-point pass."""
    from st_dadk_b200.predict import Predictor
    model.eval()
    out, _ = Predictor(model).space_time_field(torch.as_tensor(coords, device=device), T)
    q = out.shape[1]
    return out[:, q // 2].reshape(T, -1).cpu().numpy()


def save_results(results, output_dir):
    def conv(o):
        if isinstance(o, np.ndarray):
            return o.tolist()
        if isinstance(o, (np.floating,)):
            return float(o)
        if isinstance(o, (np.integer,)):
            return int(o)
        if isinstance(o, dict):
            return {k: conv(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [conv(v) for v in o]
        return o
    with open(Path(output_dir) / "results.json", "w") as f:
        json.dump(conv(results), f, indent=2)


def _run_single_quantile_experiment(config, experiment_id, output_dir, device, verbose=True, parallel_mode=False):
    t0 = time.time()
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    seed = config.get("base_seed", 42) + experiment_id - 1
    set_seed(seed)
    z_full, coords, _ = load_kaust_csv_single(config.get("data_file", "data/2b/2b_7.csv"),
                                              normalize=config.get("normalize_target", False))
    fn = create_spatial_obs_prob_fn(config.get("obs_spatial_pattern", "uniform"), config.get("obs_spatial_intensity", 1.0))
    obs_mask, obs_sites = sample_observations(z_full, coords, config.get("obs_method", "site-wise"),
                                              config.get("obs_ratio", 0.5), fn, seed=seed)
    train_mask, valid_mask = split_train_valid(obs_mask, obs_sites, config.get("split_method", "site-wise"),
                                               config.get("train_ratio", 0.8), seed=seed + 10000)
    p = config.get("p_covariates", 0)
    train, valid, test = (create_dataset_from_mask(z_full, coords, m, p) for m in (train_mask, valid_mask, ~obs_mask))
    print(f"Train dataset: {len(train)} samples\nVal dataset: {len(valid)} samples\nTest dataset: {len(test)} samples")
    bs = auto_batch_size(config.get("batch_size", 256), len(train))
    cfg = dict(config, _effective_batch_size=bs)
    train_coords = None
    if config.get("spatial_init_method", "uniform") in ("gmm", "random_site", "kmeans_balanced"):
        train_coords = train.coords.numpy()
    model = create_model(cfg, train_coords=train_coords).to(device)
    print(f"Model parameters: {sum(p.numel() for p in model.parameters()):,}")
    model, history, centers_hist = train_model(model, train, valid, cfg, device, output_dir)
    metrics = {k: evaluate_model(model, d, device, cfg) for k, d in (("train", train), ("valid", valid), ("test", test))}
    total = time.time() - t0
    res = {"experiment_id": experiment_id, "experiment_seed": seed, "regression_type": config.get("regression_type", "mean"),
           "config": dict(config, output_dir=str(output_dir)), "metrics": metrics, "training_history": history,
           "total_time_seconds": total,
           "total_time_formatted": f"{int(total // 3600):02d}:{int((total % 3600) // 60):02d}:{int(total % 60):02d}",
           "model_parameters": sum(p.numel() for p in model.parameters()),
           "timestamp": datetime.now().strftime("%Y-%m-%d %H:%M:%S")}
    for split, key in (("train", "train"), ("valid", "valid"), ("test", "test")):
        for m in ("mse", "mae", "rmse"):
            res[f"{key}_{m}"] = metrics[split][m]
        if "check_loss" in metrics[split]:
            res[f"{key}_check_loss"] = metrics[split]["check_loss"]
        if "crps" in metrics[split]:
            res[f"{key}_crps"] = metrics[split]["crps"]
    if config.get("regression_type") == "multi-quantile":
        res["quantile_levels"] = config.get("quantile_levels")
    if config.get("regression_type") == "quantile":
        res["quantile_level"] = config.get("current_quantile")
    if _dp_rank() != 0:          # data parallel: the replicas are identical, rank 0 writes
        return res
    save_results(res, output_dir)
    torch.save(model.state_dict(), output_dir / "model_final.pt")
    field = predict_field(model, coords, z_full.shape[0], device, cfg)
    np.savez(output_dir / "predictions.npz", predictions=field, coords=coords, z_true=z_full, obs_mask=obs_mask)
    sb = model.spatial_basis
    np.savez(output_dir / "basis_info.npz", centers=sb.centers.detach().cpu().numpy(),
             bandwidths=sb.bandwidths.detach().cpu().numpy())
    return res


def _dp_rank() -> int:
    """Rank inside a data-parallel group (torch.distributed initialised by --data_parallel), else 0."""
    import torch.distributed as dist
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def run_single_experiment(config, experiment_id, output_dir, device, verbose=True, parallel_mode=False,
                          skip_existing=False):
    """One experiment; single-quantile regression with several levels trains one model per level (upstream :1936-2161)."""
    output_dir = Path(output_dir)
    if skip_existing and (output_dir / "results.json").exists():
        return json.load(open(output_dir / "results.json"))
    rtype, levels = config.get("regression_type", "mean"), config.get("quantile_levels", [0.5])
    if rtype == "quantile" and len(levels) > 1:
        out = {}
        for q in levels:
            out[q] = _run_single_quantile_experiment(dict(config, current_quantile=q), experiment_id,
                                                     output_dir / f"quantile_{q}", device, verbose, parallel_mode)
        return out
    if rtype == "quantile" and "current_quantile" not in config:
        config = dict(config, current_quantile=levels[0])
    return _run_single_quantile_experiment(config, experiment_id, output_dir, device, verbose, parallel_mode)


def run_multiple_experiments(config, base_output_dir, device, parallel=False, start_exp_id=None, end_exp_id=None,
                             skip_existing=False):
    """Experiments are independent: with torchrun they are dealt round-robin to the ranks (one GPU each), replacing
    upstream's joblib CPU fan-out (:2914-3026)."""
    n = config.get("n_experiments", 10)
    ids = list(range(start_exp_id or 1, (end_exp_id or n) + 1))
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = 0, 1        # data parallel: every rank takes part in every experiment
    results = []
    for i in ids[rank::world]:
        d = Path(base_output_dir) / f"experiment_{i:03d}"
        try:
            results.append(run_single_experiment(config, i, d, device, skip_existing=skip_existing))
        except Exception as e:   # keep the sweep alive, record the failure like upstream
            d.mkdir(parents=True, exist_ok=True)
            (d / "error.txt").write_text(repr(e))
            print(f"[ERROR] experiment {i}: {e!r}")
    return results


METRIC_KEYS = ("train_mse", "train_mae", "train_rmse", "valid_mse", "valid_mae", "valid_rmse", "test_mse", "test_mae",
               "test_rmse", "total_time_seconds")


def aggregate_results(all_results: list, summary_dir):
    """summary_statistics.json (mean / std / min / max / median / values per metric) and all_experiments.csv with
    upstream's names and columns (upstream :2790-2908; its spatial-MSE plots are left to the offline tools)."""
    import pandas as pd
    summary_dir = Path(summary_dir)
    data = {k: [] for k in METRIC_KEYS}
    for r in all_results:
        for k in METRIC_KEYS[:-1]:
            split, metric = k.split("_", 1)
            data[k].append(r["metrics"][split][metric] if "metrics" in r else r.get(k, 0))
        data["total_time_seconds"].append(r["total_time_seconds"])
    summary = {"n_experiments": len(all_results), "statistics": {}}
    for k, vals in data.items():
        a = np.array(vals, dtype=float)
        summary["statistics"][k] = {"mean": float(np.mean(a)), "std": float(np.std(a)), "min": float(np.min(a)),
                                    "max": float(np.max(a)), "median": float(np.median(a)), "values": [float(v) for v in vals]}
    with open(summary_dir / "summary_statistics.json", "w") as f:
        json.dump(summary, f, indent=2)
    cols = {"experiment_id": [r.get("experiment_id", i + 1) for i, r in enumerate(all_results)]}
    if all_results and "experiment_seed" in all_results[0]:
        cols["experiment_seed"] = [r["experiment_seed"] for r in all_results]
    cols.update(data)
    pd.DataFrame(cols).to_csv(summary_dir / "all_experiments.csv", index=False)
    return summary


def collect_experiment_results(base_output_dir, ids):
    """results.json of every finished experiment under `base_output_dir`, ordered by experiment id (the experiments of
    one launch are dealt to the ranks; whoever sees them all complete aggregates)."""
    out = []
    for i in ids:
        f = Path(base_output_dir) / f"experiment_{i:03d}" / "results.json"
        if not f.exists():
            return None
        out.append(json.load(open(f)))
    return out


def launch_directory(tag: str) -> Path:
    """results/<date>/<time>_<tag>, ONE directory per launch: under torchrun every rank derives it from the start time
    of the common parent (the elastic agent), not from its own clock."""
    t = datetime.now()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        try:
            import psutil
            t = datetime.fromtimestamp(psutil.Process(os.getppid()).create_time())
        except Exception:      # noqa: BLE001 -- no psutil: whole minutes of the local clock
            t = t.replace(second=0, microsecond=0)
    return Path("results") / t.strftime("%Y%m%d") / f"{t.strftime('%H%M%S')}_{tag}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, default="configs/config_st_interp.yaml")
    ap.add_argument("--data_file", type=str, default=None)
    ap.add_argument("--n_experiments", type=int, default=None)
    ap.add_argument("--base_seed", type=int, default=None)
    ap.add_argument("--parallel", action="store_true")
    ap.add_argument("--n_jobs", type=int, default=None)
    ap.add_argument("--start_exp_id", type=int, default=None)
    ap.add_argument("--end_exp_id", type=int, default=None)
    ap.add_argument("--skip-existing", action="store_true")
    ap.add_argument("--data_parallel", action="store_true",
                    help="under torchrun: train ONE model on all GPUs (global batches split over the ranks) instead of "
                         "dealing independent experiments to them")
    args = ap.parse_args()
    with open(args.config) as f:
        config = yaml.safe_load(f)
    for k in ("data_file", "n_experiments", "base_seed"):
        if getattr(args, k) is not None:
            config[k] = getattr(args, k)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = f"cuda:{local}"         # the yaml's `device: cpu` cannot be honoured: this implementation is CUDA-only
    torch.cuda.set_device(local)
    if args.data_parallel and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))
    out = launch_directory(config.get("tag", "default"))
    out.mkdir(parents=True, exist_ok=True)
    rank = int(os.environ.get("RANK", "0"))
    if rank == 0:
        yaml.safe_dump(config, open(out / "config.yaml", "w"))
    run_multiple_experiments(config, out, device, args.parallel, args.start_exp_id, args.end_exp_id, args.skip_existing)
    n = config.get("n_experiments", 10)
    ids = list(range(args.start_exp_id or 1, (args.end_exp_id or n) + 1))
    (out / f".rank{rank}.done").write_text("done")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if _dp_rank() == 0 and all((out / f".rank{r}.done").exists() for r in range(1 if args.data_parallel else world)):
        # the last rank to finish (or any rank that sees every marker) merges: same content whoever writes it
        allr = collect_experiment_results(out, ids)
        flat = [r for r in (allr or []) if isinstance(r, dict) and "total_time_seconds" in r]
        if flat:
            summ = aggregate_results(flat, out)
            print(json.dumps({k: {"mean": v["mean"], "std": v["std"]} for k, v in summ["statistics"].items()}, indent=2))


if __name__ == "__main__":
    main()

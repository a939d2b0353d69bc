"""Host-side evaluation metrics (upstream stnf/utils/metrics.py:9-163).  Not on the accelerated path: numpy
on arrays that were already copied back from the device."""
from typing import Dict, Union

import numpy as np
import torch

Array = Union[np.ndarray, torch.Tensor]


def _np(a: Array) -> np.ndarray:
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def _errors(y_true: np.ndarray, y_pred: np.ndarray):
    yt, yp = y_true.reshape(-1), y_pred.reshape(-1)
    ok = ~(np.isnan(yt) | np.isnan(yp))
    return yt[ok], yp[ok]


def compute_metrics(y_true: Array, y_pred: Array, per_horizon: bool = False) -> Dict[str, float]:
    """RMSE / MAE / R^2 / MSE over all finite pairs; optionally per forecast horizon for (B,H,S,1) inputs."""
    y_true, y_pred = _np(y_true), _np(y_pred)
    yt, yp = _errors(y_true, y_pred)
    res = yt - yp
    mse = float(np.mean(res ** 2))
    out = {"rmse": float(np.sqrt(mse)), "mae": float(np.mean(np.abs(res))),
           "r2": float(1 - np.sum(res ** 2) / (np.sum((yt - yt.mean()) ** 2) + 1e-8)), "mse": mse}
    if per_horizon and y_true.ndim == 4:
        rm, ma = [], []
        for h in range(y_true.shape[1]):
            a, b = _errors(y_true[:, h], y_pred[:, h])
            rm.append(float(np.sqrt(np.mean((a - b) ** 2))))
            ma.append(float(np.mean(np.abs(a - b))))
        out["rmse_per_horizon"], out["mae_per_horizon"] = rm, ma
    return out


def compute_spatial_metrics(y_true: Array, y_pred: Array, coords: np.ndarray, n_bins: int = 5) -> Dict[str, list]:
    """RMSE / MAE per ring of distance from the origin, for (B,H,S,1) arrays."""
    y_true, y_pred = _np(y_true), _np(y_pred)
    dist = np.hypot(coords[:, 0], coords[:, 1])
    edges = np.linspace(0, dist.max(), n_bins + 1)
    centers, rmse, mae = [], [], []
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (dist >= lo) & (dist < hi)
        if not sel.any():
            continue
        a, b = _errors(y_true[:, :, sel, :], y_pred[:, :, sel, :])
        rmse.append(float(np.sqrt(np.mean((a - b) ** 2))) if a.size else float("nan"))
        mae.append(float(np.mean(np.abs(a - b))) if a.size else float("nan"))
        centers.append(float((lo + hi) / 2))
    return {"bin_centers": centers, "rmse_by_distance": rmse, "mae_by_distance": mae}


def print_metrics(metrics: Dict[str, float], prefix: str = ""):
    if prefix:
        print(prefix)
    print(f"  RMSE: {metrics['rmse']:.6f}")
    print(f"  MAE:  {metrics['mae']:.6f}")
    print(f"  R²:   {metrics['r2']:.6f}")
    if "rmse_per_horizon" in metrics:
        print(f"  RMSE per horizon: {metrics['rmse_per_horizon']}")

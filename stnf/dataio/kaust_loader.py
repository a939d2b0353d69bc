"""KAUST CSV loader (upstream stnf/dataio/kaust_loader.py:19-76) and the device-resident sample table that
replaces the upstream list-of-dicts dataset + collate (scripts/train_st_interp.py:413-460).

`load_kaust_csv_single` keeps the upstream contract -- z (T, S) float32 NaN-filled, coords (S, 2) float32, sites
numbered by first appearance of (x, y) in the file, optional global z-normalisation -- but is vectorised
(pandas factorize + one fancy-index store) instead of two `iterrows` passes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import pandas as pd
import torch


def load_kaust_csv_single(data_path: str, normalize: bool = True) -> Tuple[np.ndarray, np.ndarray, Dict]:
    df = pd.read_csv(data_path)
    print(f"[INFO] Loaded data: {len(df)} rows")
    if "t" not in df.columns:
        # purely spatial competition files (data/1a, 1b): one time step; upstream raises KeyError('t') here
        df = df.assign(t=1)
    xy = df[["x", "y"]].to_numpy()
    # site id = order of first appearance of the (x, y) pair (drop_duplicates semantics, kaust_loader.py:40-51)
    keys = pd.MultiIndex.from_arrays([xy[:, 0], xy[:, 1]])
    site, uniq = pd.factorize(keys, sort=False)
    S = len(uniq)
    print(f"[INFO] Total sites: {S}")
    first = np.full(S, -1, dtype=np.int64)
    first[site[::-1]] = np.arange(len(site))[::-1]
    coords = xy[first].astype(np.float32)
    t_idx = df["t"].to_numpy().astype(np.int64) - 1
    T = int(df["t"].to_numpy().max())
    print(f"[INFO] Time range: 1 ~ {T}")
    z = np.full((T, S), np.nan, dtype=np.float32)
    if "z" in df.columns:
        z[t_idx, site] = df["z"].to_numpy()          # later duplicates overwrite earlier ones, as in upstream's loop
    meta: Dict = {}
    if normalize:
        vals = z[~np.isnan(z)]
        mu, sd = vals.mean(), vals.std()
        z = (z - mu) / sd
        meta["z_mean"], meta["z_std"] = mu, sd
        print(f"[INFO] Normalized z: mean={mu:.4f}, std={sd:.4f}")
    return z, coords, meta


@dataclass
class ObservationTable:
    """Struct-of-arrays sample table: one row per observed (t, site) pair, sorted by (t, site) exactly like
    upstream's np.argwhere(mask) loop (train_st_interp.py:429-448); NaN targets are dropped; t is
    t_idx/(T-1) (0 when T == 1).  `to(device)` makes it HBM-resident; batches are index gathers."""
    coords: torch.Tensor   # (N, 2)
    t: torch.Tensor        # (N,)
    y: torch.Tensor        # (N,)
    X: Optional[torch.Tensor] = None

    @staticmethod
    def from_mask(z_data: np.ndarray, coords: np.ndarray, mask: np.ndarray, p_covariates: int = 0) -> "ObservationTable":
        T = z_data.shape[0]
        ts, ss = np.nonzero(mask)
        y = z_data[ts, ss]
        keep = ~np.isnan(y)
        ts, ss, y = ts[keep], ss[keep], y[keep]
        # upstream computes t_idx / (T - 1) in Python floats (float64) and stores it as float32
        tn = (ts / (T - 1)).astype(np.float32) if T > 1 else np.zeros(len(ts), dtype=np.float32)
        X = torch.zeros(len(ts), p_covariates) if p_covariates > 0 else None
        return ObservationTable(torch.from_numpy(coords[ss].astype(np.float32)), torch.from_numpy(tn),
                                torch.from_numpy(y.astype(np.float32)), X)

    def __len__(self):
        return int(self.y.shape[0])

    def to(self, device) -> "ObservationTable":
        mv = lambda a: a.to(device, non_blocking=True) if a is not None else None
        return ObservationTable(mv(self.coords), mv(self.t), mv(self.y), mv(self.X))

    def pin(self) -> "ObservationTable":
        pn = lambda a: a.pin_memory() if a is not None else None
        return ObservationTable(pn(self.coords), pn(self.t), pn(self.y), pn(self.X))

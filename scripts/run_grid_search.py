"""Grid search over hyper-parameter configurations, packed onto GPUs.

Upstream fans configurations out to CPU worker processes with joblib (scripts/run_grid_search.py:329-387).  Here the
configurations are independent jobs dealt to the ranks of a torchrun launch (one process per GPU); with
`--configs_per_gpu C` each rank additionally splits its share over C worker PROCESSES on the same GPU.  A 50-epoch
run of the default model is ~0.1 s of GPU time and ~0.5 s of host work (CSV, knot placement, evaluation, artefact
files), so what packing buys is host parallelism; worker threads inside one interpreter were measured slower than
sequential (GIL, serialised graph captures), separate processes are not.  Outputs keep upstream's file names:
grid_search_summary.csv, grid_search_detail.csv, grid_search_configs.json.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_grid_search.py --config ... --configs_per_gpu 2
"""
import argparse
import copy
import itertools
import json
import os
import sys
import time
from datetime import datetime
from pathlib import Path

import numpy as np
import pandas as pd
import torch
import yaml

sys.path.append(str(Path(__file__).resolve().parent.parent))
from scripts.train_st_interp import run_single_experiment   # noqa: E402

# the sweep of BASELINE config 5 (64 = 4 x 2 x 2 x 2 x 2); upstream hard-codes its own grid at :257-274
DEFAULT_GRID = {
    "lr": [5e-3, 1e-2, 2e-2, 4e-2],
    "dropout": [0.0, 0.1],
    "hidden_dims": [[256, 256, 128], [128, 128]],
    "spatial_basis_function": ["wendland", "triangular"],
    "basis_mode": ["uniform-fixed", "gmm-learnable"],
}


def generate_config_combinations(base_config: dict, param_grid: dict, config_filter=None):
    """Cartesian product of `param_grid` applied to copies of `base_config`, each tagged with a config_id and a
    readable tag (upstream :22-99)."""
    keys = list(param_grid)
    out = []
    for cid, values in enumerate(itertools.product(*[param_grid[k] for k in keys]), 1):
        cfg = copy.deepcopy(base_config)
        parts = []
        for k, v in zip(keys, values):
            if k == "basis_mode":
                init, learn = v.split("-")
                cfg["spatial_init_method"], cfg["spatial_learnable"] = init, learn == "learnable"
            else:
                cfg[k] = v
            parts.append(f"{k}={'x'.join(map(str, v)) if isinstance(v, list) else v}")
        if config_filter is not None and not config_filter(cfg):
            continue
        cfg["config_id"] = cid
        cfg["tag"] = "_".join(parts)
        out.append(cfg)
    return out


def _run_config(cfg, out_dir, device, results):
    t0 = time.time()
    rows = []
    for i in range(1, int(cfg.get("n_experiments", 1)) + 1):
        d = Path(out_dir) / f"config_{cfg['config_id']:03d}" / f"experiment_{i:03d}"
        try:
            r = run_single_experiment(cfg, i, d, device, verbose=False, skip_existing=True)
            rows.append({"config_id": cfg["config_id"], "tag": cfg["tag"], "experiment_id": i,
                         **{k: r[k] for k in ("train_rmse", "valid_rmse", "test_rmse", "test_mae") if k in r},
                         "total_time_seconds": r.get("total_time_seconds")})
        except Exception as e:
            d.mkdir(parents=True, exist_ok=True)
            (d / "error.txt").write_text(repr(e))
    results.extend(rows)
    print(f"[grid] config {cfg['config_id']} ({cfg['tag']}) done in {time.time() - t0:.1f}s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="configs/config_st_interp.yaml")
    ap.add_argument("--grid", default=None, help="json file with a param grid (default: the 64-config sweep)")
    ap.add_argument("--output_dir", default=None)
    ap.add_argument("--configs_per_gpu", type=int, default=1)
    ap.add_argument("--n_experiments", type=int, default=None)
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--worker_slice", default=None, help="internal: 'i:C' = this process is worker i of C of its rank")
    args = ap.parse_args()
    base = yaml.safe_load(open(args.config))
    if args.n_experiments is not None:
        base["n_experiments"] = args.n_experiments
    if args.epochs is not None:
        base["epochs"] = args.epochs
    grid = json.load(open(args.grid)) if args.grid else DEFAULT_GRID
    configs = generate_config_combinations(base, grid)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    out = Path(args.output_dir or Path("results") / f"grid_{datetime.now().strftime('%Y%m%d_%H%M%S')}")
    out.mkdir(parents=True, exist_ok=True)
    mine = configs[rank::world]
    n_workers = max(1, args.configs_per_gpu)
    t0 = time.time()
    if n_workers > 1 and args.worker_slice is None:
        # parent of this rank: C worker processes on the same GPU, each with every C-th configuration of the share
        import subprocess
        base_cmd = [sys.executable, os.path.abspath(__file__), "--config", args.config, "--output_dir", str(out),
                    "--configs_per_gpu", str(n_workers)]
        if args.grid:
            base_cmd += ["--grid", args.grid]
        if args.n_experiments is not None:
            base_cmd += ["--n_experiments", str(args.n_experiments)]
        if args.epochs is not None:
            base_cmd += ["--epochs", str(args.epochs)]
        procs = [subprocess.Popen(base_cmd + ["--worker_slice", f"{i}:{n_workers}"]) for i in range(n_workers)]
        rcs = [p.wait() for p in procs]
        parts = [out / f"grid_search_detail_rank{rank}_w{i}.csv" for i in range(n_workers)]
        frames = []
        for f in parts:
            try:
                frames.append(pd.read_csv(f))
            except (FileNotFoundError, pd.errors.EmptyDataError):
                pass
        results = pd.concat(frames).to_dict("records") if frames else []
        if any(rcs):
            print(f"[grid] worker exit codes: {rcs}", flush=True)
    else:
        results = []
        if args.worker_slice is not None:
            wi, wc = (int(v) for v in args.worker_slice.split(":"))
            mine = mine[wi::wc]
        for cfg in mine:
            _run_config(cfg, out, device, results)
        torch.cuda.synchronize()
        if args.worker_slice is not None:
            pd.DataFrame(results).to_csv(out / f"grid_search_detail_rank{rank}_w{wi}.csv", index=False)
            return
    wall = time.time() - t0
    detail = pd.DataFrame(results)
    detail.to_csv(out / f"grid_search_detail_rank{rank}.csv", index=False)
    if len(detail):
        summary = detail.groupby(["config_id", "tag"]).agg(["mean", "std"]).reset_index()
        summary.columns = ["_".join(c).strip("_") for c in summary.columns]
        summary.to_csv(out / f"grid_search_summary_rank{rank}.csv", index=False)
    if rank == 0:
        json.dump([{k: v for k, v in c.items()} for c in configs], open(out / "grid_search_configs.json", "w"), indent=1,
                  default=str)
    print(json.dumps({"rank": rank, "configs": len(mine), "wall_s": wall, "configs_per_gpu": n_workers}))


if __name__ == "__main__":
    main()

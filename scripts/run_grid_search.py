"""Grid search over hyper-parameter configurations, packed onto GPUs.

Upstream fans configurations out to CPU worker processes with joblib (scripts/run_grid_search.py:329-387).  Here the
configurations are independent jobs dealt to the ranks of a torchrun launch (one process per GPU); with
`--configs_per_gpu C` each rank additionally splits its share over C worker PROCESSES on the same GPU (forked before
the rank touches CUDA, so a worker costs no second interpreter / torch start-up).  A 50-epoch
run of the default model is ~0.1 s of GPU time and ~0.5 s of host work (CSV, knot placement, evaluation, artefact
files), so what packing buys is host parallelism; worker threads inside one interpreter were measured slower than
sequential (GIL, serialised graph captures), separate processes are not.  Outputs are upstream's (:102-237): ONE
grid_search_summary.csv, grid_search_detail.csv, grid_search_configs.json / .csv per launch with the same columns, and
summary_statistics.json + all_experiments.csv inside every config directory -- merged by whichever rank finishes last
from the per-experiment results.json files (the ranks never communicate).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_grid_search.py --config ... --configs_per_gpu 2
"""
import time as _time
_T_PROCESS = _time.time()          # for the start-up timeline printed at the end (imports below are part of it)
import argparse
import copy
import itertools
import json
import os
import sys
import time
from pathlib import Path

import pandas as pd
import torch
import yaml

sys.path.append(str(Path(__file__).resolve().parent.parent))
from scripts.train_st_interp import (run_single_experiment, aggregate_results, collect_experiment_results,   # noqa: E402
                                     launch_directory)

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


STAT_METRICS = ("test_rmse", "test_mae", "test_mse", "valid_rmse", "valid_mae", "valid_mse", "train_rmse", "train_mae",
                "train_mse", "total_time_seconds")
CONFIG_COLUMNS = ("spatial_basis_function", "spatial_init_method", "spatial_learnable", "obs_method", "obs_ratio",
                  "obs_spatial_pattern")


def save_experiment_results(all_results, output_dir):
    """grid_search_summary.csv (one row per config: mean / std / min / max / median of every metric),
    grid_search_detail.csv (one row per config x experiment), grid_search_configs.json / .csv -- upstream's files and
    columns (:102-237)."""
    output_dir = Path(output_dir)
    summary_rows, detail_rows, configs, index = [], [], {}, []
    for res in all_results:
        if res is None or res.get("summary") is None:
            continue
        cfg, summ = res["config"], res["summary"]
        base = {"config_id": cfg["config_id"], "tag": cfg["tag"],
                **{k: cfg.get(k, "wendland" if k == "spatial_basis_function" else None) for k in CONFIG_COLUMNS}}
        row = dict(base, n_experiments=summ["n_experiments"])
        for m in STAT_METRICS:
            st = summ["statistics"].get(m)
            if st is not None:
                for k in ("mean", "std", "min", "max", "median"):
                    row[f"{m}_{k}"] = st[k]
        summary_rows.append(row)
        n_exp = summ["n_experiments"]
        for e in range(n_exp):
            d = {"config_id": cfg["config_id"], "tag": cfg["tag"], "experiment_id": e + 1,
                 **{k: base[k] for k in CONFIG_COLUMNS}}
            for m in STAT_METRICS:
                if m in summ["statistics"]:
                    d[m] = summ["statistics"][m]["values"][e]
            detail_rows.append(d)
        configs[str(cfg["config_id"])] = cfg
        index.append({"config_id": cfg["config_id"], "tag": cfg["tag"]})
    df_summary, df_detail = pd.DataFrame(summary_rows), pd.DataFrame(detail_rows)
    df_summary.to_csv(output_dir / "grid_search_summary.csv", index=False)
    df_detail.to_csv(output_dir / "grid_search_detail.csv", index=False)
    with open(output_dir / "grid_search_configs.json", "w", encoding="utf-8") as f:
        json.dump(configs, f, indent=2, ensure_ascii=False, default=str)
    pd.DataFrame(index).to_csv(output_dir / "grid_search_configs.csv", index=False)
    return df_summary, df_detail


def merge_outputs(configs, out):
    """Per-config aggregation (summary_statistics.json, all_experiments.csv in the config directory, as upstream's
    run_multiple_experiments leaves them) and the launch-level CSVs, from the results.json files on disk."""
    all_results = []
    for cfg in configs:
        cdir = Path(out) / f"config_{cfg['config_id']:03d}"
        ids = list(range(1, int(cfg.get("n_experiments", 1)) + 1))
        rs = collect_experiment_results(cdir, ids) if cdir.exists() else None
        rs = [r for r in (rs or []) if isinstance(r, dict) and "total_time_seconds" in r]
        all_results.append({"config": cfg, "summary": aggregate_results(rs, cdir) if rs else None})
    return save_experiment_results(all_results, out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="configs/config_st_interp.yaml")
    ap.add_argument("--grid", default=None, help="json file with a param grid (default: the 64-config sweep)")
    ap.add_argument("--output_dir", default=None)
    ap.add_argument("--configs_per_gpu", type=int, default=1)
    ap.add_argument("--n_experiments", type=int, default=None)
    ap.add_argument("--epochs", type=int, default=None)
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
    device = f"cuda:{local}"
    t_launch = float(os.environ.get("STDADK_SWEEP_T0", str(_T_PROCESS)))     # when the launcher started this sweep
    out = Path(args.output_dir) if args.output_dir else launch_directory("grid_search")   # one directory per launch
    out.mkdir(parents=True, exist_ok=True)
    mine = configs[rank::world]
    n_workers = max(1, args.configs_per_gpu)
    # Host side of a sweep: the work per configuration is host-bound (driver, scikit-learn knot placement), so the host
    # cores are shared out among the processes of the box (torchrun pins OMP_NUM_THREADS=1, which makes every mixture
    # fit 5-10x slower), and mixture fits are shared between processes through files (knot_init._gmm_fit_shared).
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    threads = max(1, (os.cpu_count() or 1) // max(1, local_world * n_workers))
    os.environ.setdefault("STDADK_GMM_CACHE_DIR", str(out / ".gmm_cache"))
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = str(threads)

    def run_share(share, tag):
        """One worker: its configurations one after the other on this rank's GPU (own CUDA context, own streams)."""
        try:
            torch.set_num_threads(threads)
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=threads)
        except Exception:
            pass
        torch.cuda.set_device(local)
        t_start = time.time()
        results, t_first = [], None
        for cfg in share:
            _run_config(cfg, out, device, results)
            if t_first is None:
                t_first = time.time()
        torch.cuda.synchronize()
        print(f"[grid] {tag}: {len(share)} configurations; since launch: worker started +{t_start - t_launch:.1f} s, first "
              f"configuration done +{(t_first or t_start) - t_launch:.1f} s, all done +{time.time() - t_launch:.1f} s", flush=True)

    t0 = time.time()
    if n_workers > 1:
        # C worker processes on the same GPU, each with every C-th configuration of the share.  They are FORKED from
        # this process before it touches CUDA (the interpreter and the imported torch are shared copy-on-write: a
        # worker costs no second start-up, which at this problem size is longer than its share of the sweep).
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        procs = [ctx.Process(target=run_share, args=(mine[i::n_workers], f"rank {rank} worker {i}")) for i in range(n_workers)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        rcs = [p.exitcode for p in procs]
        if any(rcs):
            print(f"[grid] worker exit codes: {rcs}", flush=True)
    else:
        run_share(mine, f"rank {rank}")
    wall = time.time() - t0
    (out / f".rank{rank}.done").write_text(f"{wall:.3f}")
    merged = False
    if all((out / f".rank{r}.done").exists() for r in range(world)):
        # every rank's share is on disk: whoever sees that merges (identical content if two ranks do)
        merge_outputs(configs, out)
        merged = True
    print(json.dumps({"rank": rank, "configs": len(mine), "wall_s": wall, "configs_per_gpu": n_workers, "merged": merged}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Secondary measurements (not the driver's bench contract): BASELINE config 4 shape on one GPU -- 3-level ~100k-knot
basis (sides 100/174/244, K_s = 99,812; W1 = 256 x 99,882 = 102 MB), uniform/fixed/Wendland, MSE -- training step
throughput at several per-GPU batch sizes, and the packed-configs experiment.  Prints one JSON object per line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def config4(batches=(4096, 65536), steps=20):
    from stnf.models import STInterpMLP
    from stnf.dataio import ObservationTable
    from st_dadk_b200.trainer import Trainer
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = STInterpMLP(k_spatial_centers=[10000, 30276, 59536], k_temporal_centers=[10, 15, 45],
                        hidden_dims=[256, 256, 128], dropout=0.1, layernorm=True, output_dim=1)
    n = 4_000_000
    rng = np.random.default_rng(2025)
    coords = rng.random((n, 2)).astype(np.float32)
    t = (rng.integers(0, 100, n) / 99.0).astype(np.float32)
    y = (np.sin(2 * np.pi * (coords[:, 0] + t)) * np.cos(2 * np.pi * coords[:, 1])).astype(np.float32)
    table = ObservationTable(torch.from_numpy(coords), torch.from_numpy(t), torch.from_numpy(y)).to(dev)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).to(dev)
    cfg = dict(lr=1e-3, weight_decay=5e-4, grad_clip=10.0, regression_type="mean")
    tr = Trainer(model, cfg, dev, batches_per_epoch=100, use_cuda_graph=True)
    assert tr.ex.sparse
    n_params = tr.flat.n
    for B in batches:
        for i in range(4):
            tr.train_step(table, perm, i * B, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            tr.train_step(table, perm, ((i + 4) * B) % (n - B), B)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kt = tr.profile_step(table, perm, B, B, repeats=3)
        print(json.dumps({"workload": "config4: K_s=99812 (100^2+174^2+244^2), W1 102 MB, support-walking block 1",
                          "batch": B, "ms_per_step": ms, "train_samples_per_s": B / (ms * 1e-3), "params": n_params,
                          "loss": tr.pop_loss_sum() / (steps + 4 + 4),
                          "kernel_ms": {k: round(v["ms"], 4) for k, v in kt["kernels"].items()}}), flush=True)


def _write_kaust_csv(path, sites, T, seed=2025, with_t=True):
    """Synthetic stand-in for the upstream CSVs (not shipped): z = sin(2pi(x+t)) cos(2pi y) + 0.5 sin(6pi x y) + 0.1 N(0,1)
    (SURVEY.md section 8d, config 2), KAUST long format x,y,t,z."""
    rng = np.random.default_rng(seed)
    S = sites.shape[0]
    with open(path, "w") as f:
        f.write("x,y,t,z\n")
        for k in range(1, T + 1):
            tt = (k - 1) / max(T - 1, 1)
            z = (np.sin(2 * np.pi * (sites[:, 0] + tt)) * np.cos(2 * np.pi * sites[:, 1])
                 + 0.5 * np.sin(6 * np.pi * sites[:, 0] * sites[:, 1]) + 0.1 * rng.standard_normal(S))
            f.write("\n".join(f"{a:.6f},{b:.6f},{k},{c:.6f}" for (a, b), c in zip(sites, z)) + "\n")


def config1(epochs=30, out_root="/tmp/stdadk_cfg1"):
    """BASELINE config 1 shape: configs/config_st_interp.yaml as shipped (gmm knots, learnable, multi-quantile Q=5)
    and the uniform/fixed/mean variant on a 1a-shaped file (90,000 points, one time step; obs 0.1 / train 0.8 =>
    7,200 training samples, batch auto-halved to 512), whole driver: CSV load, knot init, training, evaluation,
    dense prediction, artefact files.  Wall-clock per run."""
    import yaml
    from scripts.train_st_interp import run_single_experiment
    os.makedirs(out_root, exist_ok=True)
    rng = np.random.default_rng(7)
    csv = os.path.join(out_root, "1a_like.csv")
    _write_kaust_csv(csv, np.round(rng.random((90000, 2)), 6), 1)
    base = yaml.safe_load(open(os.path.join(ROOT, "configs", "config_st_interp.yaml")))
    base.update(data_file=csv, epochs=epochs, patience=epochs, n_experiments=1, obs_method="random")
    variants = {"shipped (gmm, learnable, Q=5)": {},
                "uniform / fixed / mean": dict(spatial_init_method="uniform", spatial_learnable=False,
                                               regression_type="mean")}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):          # warm-up: CUDA context, library load, kernel attributes
        run_single_experiment(dict(base, epochs=2, spatial_init_method="uniform", spatial_learnable=False,
                                   regression_type="mean"), 1, os.path.join(out_root, "warm"), "cuda:0", verbose=False)
    for name, upd in variants.items():
        cfg = dict(base, **upd)
        d = os.path.join(out_root, name.split()[0])
        import contextlib
        import io
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            r = run_single_experiment(cfg, 1, d, "cuda:0", verbose=False)
        torch.cuda.synchronize()
        wall = time.time() - t0
        print(json.dumps({"workload": f"config1: 1a-shaped, {name}", "epochs": epochs, "wall_s": wall,
                          "train_time_s": r.get("training_time_seconds", r.get("total_time_seconds")),
                          "test_rmse": r.get("test_rmse"), "train_samples": 7200, "batch": 512}), flush=True)


def config5(epochs=50, configs_per_gpu=(1, 4), out_root="/tmp/stdadk_cfg5"):
    """BASELINE config 5: the 64-configuration sweep (lr x dropout x hidden x basis function x knot mode) on a
    2a-shaped data set (S=1000 sites x T=100), n_experiments=1, fixed epochs, through scripts/run_grid_search.py on
    ONE GPU: sequentially and with several configurations in flight on separate streams."""
    import subprocess
    import yaml
    os.makedirs(out_root, exist_ok=True)
    rng = np.random.default_rng(11)
    csv = os.path.join(out_root, "2a_like.csv")
    _write_kaust_csv(csv, np.round(rng.random((1000, 2)), 6), 100)
    base = yaml.safe_load(open(os.path.join(ROOT, "configs", "config_st_interp.yaml")))
    base.update(data_file=csv, epochs=epochs, patience=epochs, n_experiments=1, obs_method="random",
                regression_type="mean")
    ypath = os.path.join(out_root, "base.yaml")
    yaml.safe_dump(base, open(ypath, "w"))
    for c in configs_per_gpu:
        out = os.path.join(out_root, f"sweep_c{c}")
        t0 = time.time()
        pr = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_grid_search.py"), "--config", ypath,
                             "--output_dir", out, "--configs_per_gpu", str(c)], capture_output=True, text=True)
        wall = time.time() - t0
        last = [ln for ln in pr.stdout.strip().splitlines() if ln.startswith("{")]
        errs = sum(1 for r, _, fs in os.walk(out) for f in fs if f == "error.txt")
        print(json.dumps({"workload": "config5: 64-config sweep, 2a-shaped (S=1000, T=100), 1 GPU", "epochs": epochs,
                          "configs_per_gpu": c, "wall_s": wall, "rc": pr.returncode, "failed_configs": errs,
                          "rank_line": json.loads(last[-1]) if last else None,
                          "stderr_tail": pr.stderr[-300:] if pr.returncode else ""}), flush=True)


def _dist_setup():
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        from datetime import timedelta
        dist.init_process_group("nccl", device_id=dev, timeout=timedelta(seconds=300))
        dist.barrier()
    return rank, world, dev


def _max_over_ranks(ms, dev, world):
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_config3(args, emit):
    """BASELINE config 3: 10M-point dense grid (1000 x 1000 x 10, generated on the device), point-sharded, Q=1 and Q=5."""
    from stnf.models import STInterpMLP
    from st_dadk_b200.predict import Predictor
    rank, world, dev = _dist_setup()
    res = {}
    for q in (1, 5):
        torch.manual_seed(0)
        model = STInterpMLP(output_dim=q).to(dev).eval()
        pr = Predictor(model, static_weights=True)
        for _ in range(max(args.warmup, 1) if args.warmup < 5 else 3):
            pr.grid(1000, 1000, 10, rank, world)
        torch.cuda.synchronize()
        reps = max(1, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out, _ = pr.grid(1000, 1000, 10, rank, world)
        e1.record()
        torch.cuda.synchronize()
        ms = _max_over_ranks(e0.elapsed_time(e1) / reps, dev, world)
        res[f"q{q}"] = {"ms": ms, "points_per_s": 1e7 / (ms * 1e-3), "kernel": "field" if pr.used_field_kernel else "generic"}
    if rank == 0:
        emit({"metric": "predict_points_per_s", "value": res["q1"]["points_per_s"], "unit": "points/s", "n_gpus": world,
              "steps": reps, "warmup": args.warmup, "ms_per_step": res["q1"]["ms"], "higher_is_better": True,
              "scaling": "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
              "config": {"workload": "BASELINE config 3: 10,000,000-point dense grid 1000x1000x10 generated on device, "
                                     "contiguous block sharding by point, default model", "q5": res["q5"], "q1": res["q1"],
                         "l2": "output 40 MB x Q per pass; inputs generated (0 B)"}})


def run_config4(args, emit):
    """BASELINE config 4: training with the 3-level ~100k-knot basis (K_s = 99,812, W1 102 MB), data parallel over the
    ranks with one all-reduce of the flat gradient (102.8 MB) per step; per-GPU batch 65,536."""
    from stnf.models import STInterpMLP
    from stnf.dataio import ObservationTable
    from st_dadk_b200.trainer import Trainer
    rank, world, dev = _dist_setup()
    torch.manual_seed(0)
    model = STInterpMLP(k_spatial_centers=[10000, 30276, 59536], k_temporal_centers=[10, 15, 45],
                        hidden_dims=[256, 256, 128], dropout=0.1, layernorm=True, output_dim=1)
    n = 50_000_000 // max(world, 8) if world > 1 else 6_250_000      # this rank's share of the 50M observations
    rng = np.random.default_rng(2025 + rank)
    coords = rng.random((n, 2), dtype=np.float32)
    t = (rng.integers(0, 100, n) / 99.0).astype(np.float32)
    y = (np.sin(2 * np.pi * (coords[:, 0] + t)) * np.cos(2 * np.pi * coords[:, 1])).astype(np.float32)
    table = ObservationTable(torch.from_numpy(coords), torch.from_numpy(t), torch.from_numpy(y)).to(dev)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1 + rank)).to(dev)
    tr = Trainer(model, dict(lr=1e-3, weight_decay=5e-4, grad_clip=10.0, regression_type="mean"), dev,
                 batches_per_epoch=100, use_cuda_graph=True)
    B = 65536
    out = {}
    for i in range(max(args.warmup, 3) if args.warmup < 8 else 4):
        tr.train_step(table, perm, i * B, B, B * world)
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, 30))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        tr.train_step(table, perm, ((i + 4) * B) % (n - B), B, B * world)
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, dev, world)
    ar_ms = None
    if world > 1:
        import torch.distributed as dist
        g = tr.flat.g[:tr.flat.n_exchange]
        for _ in range(3):
            dist.all_reduce(g)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            dist.all_reduce(g)
        e1.record()
        torch.cuda.synchronize()
        ar_ms = _max_over_ranks(e0.elapsed_time(e1) / 10, dev, world)
        tr.flat.g.zero_()
    if rank == 0:
        emit({"metric": "train_samples_per_s", "value": B * world / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
              "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
              "config": {"workload": "BASELINE config 4: K_s = 99,812 knots (100^2 + 174^2 + 244^2), W1 102 MB, support-walking "
                                     "block 1, 50M observations over 8 ranks (6.25M resident per GPU)", "global_batch": B * world,
                         "params": tr.flat.n, "allreduce_bytes": 4 * tr.flat.n_exchange, "allreduce_ms_alone": ar_ms,
                         "parallelism": f"dp{world}" if world > 1 else "single"}})


def run_config(which, args, emit):
    {1: run_config1, 3: run_config3, 4: run_config4, 5: run_config5}[which](args, emit)


def run_config1(args, emit):
    raise SystemExit("config 1: use `python bench_extra.py config1`")


def run_config5(args, emit):
    raise SystemExit("config 5: use `python bench_extra.py config5`")


if __name__ == "__main__":
    which = sys.argv[1:] or ["config4"]
    for w in which:
        {"config4": config4, "config1": config1, "config5": config5}[w]()

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
        dist.init_process_group("nccl", device_id=dev, timeout=timedelta(seconds=1800))
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
    reps = max(1, min(args.steps, 10))

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return _max_over_ranks(e0.elapsed_time(e1) / reps, dev, world)

    for q in (1, 5):
        torch.manual_seed(0)
        model = STInterpMLP(output_dim=q).to(dev).eval()
        pr = Predictor(model, static_weights=True)
        ms_sites = timed(lambda: pr.grid_by_sites(1000, 1000, 10, rank, world))
        ms_blocks = timed(lambda: pr.grid(1000, 1000, 10, rank, world))
        res[f"q{q}"] = {"ms": ms_sites, "points_per_s": 1e7 / (ms_sites * 1e-3), "ms_contiguous_point_blocks": ms_blocks,
                        "points_per_s_contiguous_point_blocks": 1e7 / (ms_blocks * 1e-3),
                        "kernel": "field" if pr.used_field_kernel else "generic"}
    if rank == 0:
        emit({"metric": "predict_points_per_s", "value": res["q1"]["points_per_s"], "unit": "points/s", "n_gpus": world,
              "steps": reps, "warmup": 3, "ms_per_step": res["q1"]["ms"], "higher_is_better": True,
              "scaling": "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
              "config": {"workload": "BASELINE config 3: 10,000,000-point dense grid 1000x1000x10 generated on device, default "
                                     "model; sharded by point: every rank takes its 1/N of the SITES at all 10 time steps "
                                     "(contiguous blocks of the point index are timed beside it)", "q5": res["q5"], "q1": res["q1"],
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
                 batches_per_epoch=100, use_cuda_graph=not getattr(args, "no_graph", False))
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


# ---------------------------------------------------------------------------------------------------------------
# Reference legs of configs 1 and 5: the reference's OWN driver (oracle/_ref/scripts/train_st_interp.py, copied
# unmodified by oracle/Makefile) on the box's host cores.  bench.py's CPU-baseline leg is one of the places allowed to
# execute oracle/ (test infrastructure); nothing here is on the product path.
def _ref_job(cfg, out_dir, threads):
    """One reference experiment on the CPU (runs inside a joblib worker or in-process).  matplotlib / seaborn are absent
    from this image: inert mocks; the plotting that follows the numeric work may raise on them and is ignored."""
    import sys as _sys
    import time as _time
    from pathlib import Path
    from unittest import mock
    import torch as _torch
    _torch.set_num_threads(threads)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.patches", "matplotlib.colors",
                 "matplotlib.cm", "seaborn", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        _sys.modules.setdefault(name, mock.MagicMock())
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    for k in [k for k in _sys.modules if k == "stnf" or k.startswith("stnf.") or k == "train_st_interp"]:
        if ref_root not in (getattr(_sys.modules[k], "__file__", "") or ""):
            del _sys.modules[k]
    _sys.path.insert(0, os.path.join(ref_root, "scripts"))
    _sys.path.insert(0, ref_root)
    import contextlib
    import io
    import train_st_interp as ref
    Path(out_dir).mkdir(parents=True, exist_ok=True)
    t0 = _time.time()
    err = None
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        try:
            ref.run_single_experiment(cfg, 1, Path(out_dir), "cpu", verbose=False)
        except Exception as e:       # mocked plotting after results.json is on disk
            err = f"{type(e).__name__}: {e}"[:120]
    wall = _time.time() - t0
    done = any(f == "results.json" for _, _, fs in os.walk(out_dir) for f in fs)
    return {"wall_s": wall, "numeric_work_done": done, "raised_after": err}


def _ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "scripts", "train_st_interp.py"))


def _cfg1_base(out_root):
    import yaml
    os.makedirs(out_root, exist_ok=True)
    rng = np.random.default_rng(7)
    csv = os.path.join(out_root, "1a_like.csv")
    if not os.path.exists(csv):
        _write_kaust_csv(csv, np.round(rng.random((90000, 2)), 6), 1)
    base = yaml.safe_load(open(os.path.join(ROOT, "configs", "config_st_interp.yaml")))
    base.update(data_file=csv, n_experiments=1, obs_method="random")
    return base


def run_config1(args, emit):
    """BASELINE config 1: configs/config_st_interp.yaml as shipped (gmm knots, learnable, multi-quantile Q=5) on a
    1a-shaped file (90,000 points, one time step => 7,200 training samples, batch 512, 15 batches / epoch) through
    scripts/train_st_interp.py's run_single_experiment -- the whole driver: CSV load, knot initialisation, training,
    evaluation, prediction, artefacts.  The reference's own driver is timed beside it on the host cores for a bounded
    number of epochs.  `value` = training samples / s of the epoch loop (epochs * 7,200 / time between the first and
    the last epoch), the driver's fixed costs are reported separately."""
    import contextlib
    import io
    from scripts.train_st_interp import run_single_experiment
    rank, world, dev = _dist_setup()
    if rank != 0:
        return
    out_root = "/tmp/stdadk_cfg1"
    base = _cfg1_base(out_root)
    epochs = max(10, min(args.steps, 100))
    with contextlib.redirect_stdout(io.StringIO()):          # warm-up: CUDA context, library load, kernel attributes
        run_single_experiment(dict(base, epochs=2, patience=2), 1, os.path.join(out_root, "warm"), str(dev), verbose=False)
    walls = {}
    for e in (2, epochs):                                    # two lengths: the slope is the per-epoch cost
        cfg = dict(base, epochs=e, patience=e)
        torch.cuda.synchronize()
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            r = run_single_experiment(cfg, 1, os.path.join(out_root, f"gpu_e{e}"), str(dev), verbose=False)
        torch.cuda.synchronize()
        walls[e] = time.time() - t0
    per_epoch = (walls[epochs] - walls[2]) / (epochs - 2)
    line = {"metric": "train_samples_per_s", "value": 7200 / per_epoch, "unit": "samples/s", "n_gpus": 1, "steps": epochs,
            "warmup": 2, "ms_per_step": per_epoch * 1e3 / 15, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": "BASELINE config 1: shipped yaml (gmm knots, learnable, Q=5 pinball) on a 1a-shaped file, "
                                   "90,000 points / 7,200 training samples, batch 512, through run_single_experiment",
                       "epochs": epochs, "wall_s": walls[epochs], "wall_s_2_epochs": walls[2], "s_per_epoch": per_epoch,
                       "test_rmse": r.get("test_rmse") if isinstance(r, dict) else None}}
    if _ref_available():
        cores = os.cpu_count() or 1
        rw = {}
        for e in (1, 3):
            rw[e] = _ref_job(dict(base, epochs=e, patience=e, device="cpu"), os.path.join(out_root, f"ref_e{e}"), cores)
        ref_epoch = (rw[3]["wall_s"] - rw[1]["wall_s"]) / 2
        line["cpu_baseline"] = {"value": 7200 / ref_epoch, "unit": "samples/s", "cores": cores, "kind": "reference",
                                "sample": "the reference's run_single_experiment on the same file and yaml, 1 and 3 epochs "
                                          "(slope = s/epoch); wall %.1f s / %.1f s" % (rw[1]["wall_s"], rw[3]["wall_s"]),
                                "s_per_epoch": ref_epoch, "fixed_cost_s": rw[1]["wall_s"] - ref_epoch,
                                "numeric_work_done": rw[3]["numeric_work_done"]}
        line["config"]["fixed_cost_s"] = walls[2] - 2 * per_epoch
    emit(line)


def run_config5(args, emit):
    """BASELINE config 5: the 64-configuration sweep (lr x dropout x hidden x basis function x knot mode) on a 2a-shaped
    file (S = 1000 sites x T = 100), n_experiments = 1, fixed epochs, through scripts/run_grid_search.py: the
    configurations are dealt round-robin to the ranks (one per GPU), `configs_per_gpu` of them in flight on each GPU.
    `value` = configurations finished per hour over the whole job.  The reference leg: the reference's own
    run_single_experiment on the host, one configuration per core in parallel (what its --parallel --n_jobs does),
    on the first `cores` configurations of the same sweep at the same epochs."""
    import subprocess
    import yaml
    import torch.distributed as dist
    rank, world, dev = _dist_setup()
    epochs = 50 if args.steps >= 20 else max(2, args.steps)
    per_gpu = int(os.environ.get("STDADK_CONFIGS_PER_GPU", "4"))
    out_root = f"/tmp/stdadk_cfg5_w{world}"
    if rank == 0:
        os.makedirs(out_root, exist_ok=True)
        rng = np.random.default_rng(11)
        _write_kaust_csv(os.path.join(out_root, "2a_like.csv"), np.round(rng.random((1000, 2)), 6), 100)
        base = yaml.safe_load(open(os.path.join(ROOT, "configs", "config_st_interp.yaml")))
        base.update(data_file=os.path.join(out_root, "2a_like.csv"), epochs=epochs, patience=epochs, n_experiments=1,
                    obs_method="random", regression_type="mean")
        yaml.safe_dump(base, open(os.path.join(out_root, "base.yaml"), "w"))
        import shutil
        shutil.rmtree(os.path.join(out_root, "sweep"), ignore_errors=True)
    if world > 1:
        dist.barrier()
    t0 = time.time()
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_grid_search.py"), "--config",
                         os.path.join(out_root, "base.yaml"), "--output_dir", os.path.join(out_root, "sweep"),
                         "--configs_per_gpu", str(per_gpu)], capture_output=True, text=True,
                        env=dict(os.environ, STDADK_SWEEP_T0=str(t0)))
    wall = time.time() - t0
    timeline = [ln for ln in pr.stdout.splitlines() if ln.startswith("[grid] rank") and "since launch" in ln]
    t = torch.tensor([wall, float(pr.returncode != 0)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, failed_rank = float(t[0]), bool(t[1] > 0)
    if rank != 0:
        return
    out = os.path.join(out_root, "sweep")
    errs = sum(1 for _, _, fs in os.walk(out) for f in fs if f == "error.txt")
    n_done = sum(1 for _, _, fs in os.walk(out) for f in fs if f == "results.json")
    line = {"metric": "grid_search_configs_per_hour", "value": 64 / wall * 3600, "unit": "configs/h", "n_gpus": world,
            "steps": epochs, "warmup": 0, "ms_per_step": wall * 1e3 / 64, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: 64-configuration sweep on a 2a-shaped file (S=1000 x T=100), "
                                   f"{epochs} epochs each, run_grid_search.py, configurations packed {per_gpu} per GPU",
                       "wall_s": wall, "configs_per_gpu": per_gpu, "results_json_files": n_done, "failed_configs": errs,
                       "rank_failed": failed_rank, "rank0_timeline": timeline, "merged_summary": os.path.exists(os.path.join(out, "grid_search_summary.csv")),
                       "stderr_tail": pr.stderr[-300:] if pr.returncode else ""}}
    if _ref_available() and not os.environ.get("STDADK_SKIP_REFERENCE_LEG"):
        from joblib import Parallel, delayed
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import run_grid_search as gs
        base = yaml.safe_load(open(os.path.join(out_root, "base.yaml")))
        configs = gs.generate_config_combinations(base, gs.DEFAULT_GRID)
        cores = os.cpu_count() or 1
        sample = [dict(c, device="cpu") for c in configs[:min(cores, 64)]]
        t0 = time.time()
        res = Parallel(n_jobs=len(sample))(delayed(_ref_job)(c, os.path.join(out_root, "ref", c.get("tag", str(i))), 1)
                                           for i, c in enumerate(sample))
        rwall = time.time() - t0
        line["cpu_baseline"] = {"value": len(sample) / rwall * 3600, "unit": "configs/h", "cores": cores, "kind": "reference",
                                "sample": f"first {len(sample)} configurations of the same sweep, {epochs} epochs, one per core "
                                          f"(joblib, as the reference's --parallel --n_jobs {cores}); wall {rwall:.1f} s",
                                "numeric_work_done": sum(1 for r in res if r["numeric_work_done"])}
    emit(line)


if __name__ == "__main__":
    which = sys.argv[1:] or ["config4"]
    for w in which:
        {"config4": config4, "config1": config1, "config5": config5}[w]()

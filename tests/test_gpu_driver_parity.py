"""GPU: this repo's driver (scripts/train_st_interp.py -> st_dadk_b200.trainer.fit) against what the UNMODIFIED reference
driver did on the same file and seed (tests/golden/driver_*.npz, made by oracle/gen_golden_driver.py from
/root/reference: _run_single_quantile_experiment -> train_model -> evaluate_model, train_st_interp.py:2164-2505, :463-961).

Pinned per case: initial weights (same seed => same torch init calls), initial GMM knots, the RandomSampler batch order
(first sample of every training batch), the learning rate of every parameter group at every optimizer step (warm-up
written after the step, progressive unfreezing + ramp-up, chainable cosine -- the bug-compatible trajectory of
SURVEY.md 9.5), per-step and per-epoch losses, EMA validation loss / RMSE, best-checkpoint choice, final metrics
(RMSE / MAE / CRPS) and the final model's predictions.

Tolerances.  precision "tf32x3" (the parity mode): losses, validation metrics and predictions within 1e-3 relative of
the reference's FP32 CPU run (measured: 2.6e-4 worst step loss over 95 steps, 2e-5 metrics, 1.9e-4 predictions).
precision "tf32" (throughput mode, one tensor-core pass on operands rounded to 11 bits) is NOT a trajectory-parity
mode: it is run on the same cases to pin what it does deliver -- the exact items below (batches, learning rates),
per-step losses within 1e-1, final metrics within 5e-2, final predictions within 2e-1 relative L2 of the reference run
(measured 5.5e-2 / 9e-3 / 1.0e-1): operand rounding of ~5e-4 per GEMM moves a 100-step trajectory at lr 2e-2 by that
much, while at matched weights outputs agree to 4e-4 (tests/test_gpu_kernels.py).

config1_shipped (learnable knots) needs one more sentence.  Once the knots are unfrozen the reference's FP32 run is not
reproducible by the reference itself: its centre gradients go through torch.cdist's matmul expansion
(|s|^2 + |c|^2 - 2 s.c, st_interp.py:439-440), whose cancellation error for points close to a knot is ~1e-2 of the
gradient, and the same unmodified driver evaluated in FP64 (fixture `step_loss64`) departs from its FP32 run by 2e-3 at
the first unfrozen step and by up to 2e-1 afterwards.  So: up to the unfreezing epoch this path must match the FP32 run
to 1e-3 (measured 6e-5); after it, it must match the reference's FP64 run (direct-difference distances are accurate to
2.5e-6): every step of the first five epochs within 1e-3 (measured 5e-6), the last epoch within 2e-2
(a pinball residual changing sign moves the trajectory late in the run; measured 1.0e-2), validation loss / RMSE of
every epoch within 1e-3 (measured 1e-5 / 2e-4), final metrics and predictions within 3e-3.
"""
import importlib
import sys

import numpy as np
import pytest
import torch

from helpers import golden, ROOT
import driver_cases as dc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run_ours(name, tmp_path, precision, g):
    sys.path.insert(0, ROOT)
    drv = importlib.import_module("scripts.train_st_interp")
    from st_dadk_b200.trainer import Trainer
    case = dc.CASES[name]
    csv = dc.case_csv(name, tmp_path)
    config = dict(case["config"], data_file=csv, precision=precision)
    rec = {"lr": [], "loss": [], "first": [], "n": [], "state0": None, "centers0": None, "bandwidths0": None}
    orig_step, orig_create = Trainer.train_step, drv.create_model

    def step(self, table, perm, row_begin, n_rows, *a, **k):
        rec["lr"].append([float(g["lr"]) for g in self.opt.param_groups])
        i = int(perm[row_begin])
        rec["first"].append([float(table.coords[i, 0]), float(table.coords[i, 1]), float(table.t[i])])
        rec["n"].append(int(n_rows))
        orig_step(self, table, perm, row_begin, n_rows, *a, **k)
        rec["loss"].append(float(self.loss_last.item()))

    def create(cfg, train_coords=None):
        m = orig_create(cfg, train_coords=train_coords)
        rec["state0"] = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}
        rec["centers0"] = m.spatial_basis.centers.detach().cpu().numpy().copy()
        rec["bandwidths0"] = m.spatial_basis.bandwidths.detach().cpu().numpy().copy()
        sb = m.spatial_basis
        if sb.learnable:
            # The GMM fit (scikit-learn on the host, as upstream) is compared below; its last digits depend on the BLAS
            # thread count of the box, so the run continues from the reference's own initial knots: everything after
            # this point is then a function of the code under test only.
            with torch.no_grad():
                sb.centers.copy_(torch.from_numpy(g["centers0"]))
                sb.centers_init.copy_(torch.from_numpy(g["centers0"]))
                sb.log_bandwidths.copy_(torch.from_numpy(g["bandwidths0"]).log())
        return m

    Trainer.train_step, drv.create_model = step, create
    try:
        res = drv._run_single_quantile_experiment(config, int(g["experiment_id"][0]), tmp_path / "experiment_001", DEV,
                                                  verbose=False)
    finally:
        Trainer.train_step, drv.create_model = orig_step, orig_create
    return res, rec, drv, config, tmp_path / "experiment_001"


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("name", ["config2_default", "config1_shipped"])
def test_driver_matches_reference_run(name, precision, tmp_path):
    g = golden("driver_" + name)
    res, rec, drv, config, out_dir = _run_ours(name, tmp_path, precision, g)
    x3 = precision == "tf32x3"
    # ---- bit-level / exact things: initial weights, knots, batch order, learning rates
    for k, v in rec["state0"].items():
        st = g["state0_sum." + k]
        s1, s2 = float(v.astype(np.float64).sum()), float((v.astype(np.float64) ** 2).sum())
        assert abs(s1 - st[0]) <= 1e-6 * max(1.0, abs(st[0])) and abs(s2 - st[1]) <= 1e-6 * max(1.0, st[1]), k
    print("initial knots: max |dc|", np.abs(rec["centers0"] - g["centers0"]).max(), "max rel dbw",
          _rel(rec["bandwidths0"], g["bandwidths0"]).max())
    assert np.allclose(rec["centers0"], g["centers0"], atol=2e-3) and np.allclose(rec["bandwidths0"], g["bandwidths0"], rtol=2e-2)
    assert rec["n"] == g["batch_n"].tolist()
    assert np.array_equal(np.asarray(rec["first"], dtype=np.float32), g["batch_first"]), "training batch order differs"
    lr = np.asarray(rec["lr"])
    assert lr.shape == g["step_lr"].shape and np.allclose(lr, g["step_lr"], rtol=1e-9, atol=1e-15), "lr trajectory differs"
    # ---- losses per step and per epoch, EMA validation, checkpoints
    h = res["training_history"]
    step_rel = _rel(rec["loss"], g["step_loss"])
    ep_rel = {k: _rel(h[k], g["hist_" + k]) for k in ("train_loss", "val_loss", "val_rmse", "lr")}
    print(name, precision, "step loss rel: max", step_rel.max(), "first epoch max", step_rel[:len(step_rel) // len(h["lr"])].max())
    print({k: np.array2string(v, precision=2) for k, v in ep_rel.items()})
    assert len(h["train_loss"]) == len(g["hist_train_loss"])
    assert ep_rel["lr"].max() < 1e-9
    tol_step, tol_epoch = (1e-3, 1e-3) if x3 else (1e-1, 5e-2)
    ref_metric_values, ref_yhat, ref_centers = g["metric_values"], g["yhat_final32"], g["centers_final"]
    if "step_loss64" in g.files:        # learnable knots: FP32 run until the knots move, the FP64 run of the driver after
        bpe = len(step_rel) // len(h["lr"])
        frozen = int(config["basis_unfreeze_epoch"]) * bpe
        rel64 = _rel(rec["loss"], g["step_loss64"])
        band = _rel(g["step_loss"], g["step_loss64"])
        ep64 = {k: _rel(h[k], g["hist64_" + k]) for k in ("train_loss", "val_loss", "val_rmse")}
        print("vs FP32 run while frozen: max", step_rel[:frozen].max(), "| vs FP64 run, all steps: max", rel64.max(),
              "| reference FP32-vs-FP64 band: max", band.max())
        print("vs FP64 run per epoch", {k: np.array2string(v, precision=2) for k, v in ep64.items()})
        assert step_rel[:frozen].max() < tol_step, step_rel[:frozen]
        if x3:
            assert rel64[:5 * bpe].max() < 1e-3 and rel64.max() < 2e-2, rel64
        # (TF32 mode: once the knots move, single steps of this trajectory differ by tens of per cent between ANY two
        #  evaluations -- the reference's own FP32 and FP64 runs differ by 2e-1 -- so only epoch-level numbers are pinned)
        n_frozen_epochs = int(config["basis_unfreeze_epoch"])
        if x3:
            assert ep64["train_loss"].max() < 3e-3, ep64["train_loss"]
            for k in ("val_loss", "val_rmse"):
                assert ep64[k].max() < 1e-3, (k, ep64[k])
        else:       # TF32 mode: epochs with frozen knots within 5e-2; afterwards the run must stay a sane training run
            for k in ("train_loss", "val_loss", "val_rmse"):
                assert ep64[k][:n_frozen_epochs].max() < 5e-2, (k, ep64[k])
                assert ep64[k].max() < 3e-1, (k, ep64[k])
        ref_metric_values, ref_yhat, ref_centers = g["run64_metric_values"], g["run64_yhat_final"], g["run64_centers_final"]
    else:
        assert step_rel.max() < tol_step, step_rel
        for k in ("train_loss", "val_loss", "val_rmse"):
            assert ep_rel[k].max() < tol_epoch, (k, ep_rel[k])
    # ---- final metrics and predictions (model_final.pt == best EMA checkpoint, as in the reference run)
    assert bool(g["best_equals_final"][0])
    ref_metrics = dict(zip([str(s) for s in g["metric_names"]], ref_metric_values))
    worst = 0.0
    for key, val in ref_metrics.items():
        split, metric = key.split("_", 1)
        worst = max(worst, float(_rel(res["metrics"][split][metric], val)))
    print("final metrics worst rel", worst)
    learn = "step_loss64" in g.files
    assert worst < ((3e-3 if learn else 1e-3) if x3 else (2e-1 if learn else 5e-2))
    from stnf.models.st_interp import create_model
    model = create_model(dict(config, spatial_init_method="uniform"))     # shapes do not depend on the init method
    model.load_state_dict(torch.load(out_dir / "model_final.pt"))
    model = model.to(DEV).eval()
    with torch.no_grad():
        yh = model(torch.zeros(len(g["eval_coords"]), 0, device=DEV), torch.from_numpy(g["eval_coords"]).to(DEV),
                   torch.from_numpy(g["eval_t"]).to(DEV)).cpu().numpy()
    rl2 = float(np.linalg.norm(yh - ref_yhat) / np.linalg.norm(ref_yhat))
    cen = float(np.abs(model.spatial_basis.centers.detach().cpu().numpy() - ref_centers).max())
    print("final predictions rel L2", rl2, "final knots max abs diff", cen)
    assert rl2 < ((3e-3 if learn else 1e-3) if x3 else (4e-1 if learn else 2e-1))
    assert cen < (2e-4 if x3 else 2e-2)

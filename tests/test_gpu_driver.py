"""GPU: the training driver end to end on a small synthetic KAUST-format file (fixed and learnable/multi-quantile)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_csv(path, S=300, T=12, seed=0):
    rng = np.random.default_rng(seed)
    xy = np.round(rng.random((S, 2)), 6)
    rows = ["x,y,t,z"]
    for t in range(1, T + 1):
        z = np.sin(2 * np.pi * (xy[:, 0] + t / T)) * np.cos(2 * np.pi * xy[:, 1]) + 0.05 * rng.standard_normal(S)
        rows += [f"{a},{b},{t},{c:.6f}" for (a, b), c in zip(xy, z)]
    path.write_text("\n".join(rows) + "\n")


@pytest.mark.parametrize("variant", ["fixed_mean", "learnable_mq", "delta_mq"])
def test_run_single_experiment(tmp_path, variant):
    from scripts import train_st_interp as drv
    csv = tmp_path / "toy.csv"
    _write_csv(csv)
    cfg = dict(data_file=str(csv), k_spatial_centers=[9, 25], k_temporal_centers=[4, 6], hidden_dims=[64, 32],
               dropout=0.1, layernorm=True, obs_method="random", obs_ratio=0.5, split_method="random", train_ratio=0.8,
               epochs=6, lr=1e-2, weight_decay=5e-4, batch_size=256, patience=50, grad_clip=10.0, scheduler="cosine",
               warmup_epochs=2, base_seed=2025, regression_type="mean")
    if variant == "learnable_mq":
        cfg.update(spatial_init_method="gmm", spatial_learnable=True, gradient_damping=True, damping_threshold=0.0,
                   damping_strength=5.0, domain_penalty_weight=0.01, basis_unfreeze_epoch=2, basis_lr_rampup_epochs=2,
                   regression_type="multi-quantile", quantile_levels=[0.05, 0.5, 0.95], sparsity_penalty_type="sparse_group",
                   sparsity_lambda_l1=1e-4, sparsity_lambda_group=1e-4, non_crossing_weight=0.1)
    if variant == "delta_mq":
        cfg.update(regression_type="multi-quantile", quantile_levels=[0.1, 0.5, 0.9], use_delta_reparameterization=True,
                   non_crossing_lambda=1.0)
    out = tmp_path / "exp"
    res = drv.run_single_experiment(cfg, 1, out, "cuda:0", verbose=False)
    assert np.isfinite(res["test_rmse"]) and res["test_rmse"] < 5.0
    hist = res["training_history"]
    assert len(hist["train_loss"]) == 6 and all(np.isfinite(hist["train_loss"]))
    assert hist["train_loss"][-1] < hist["train_loss"][0]
    for f in ("results.json", "training_history.csv", "model_best.pt", "model_final.pt", "predictions.npz", "basis_info.npz"):
        assert (out / f).exists(), f
    pred = np.load(out / "predictions.npz")["predictions"]
    assert pred.shape == (12, 300) and np.isfinite(pred).all()
    sd = torch.load(out / "model_final.pt")
    key = "mlp_trunk.0.weight" if variant == "delta_mq" else "mlp.0.weight"
    assert sd[key].shape == (64, 34 + 10)          # upstream (out, in) shape, k_s + k_t = 34 + 10
    if variant == "learnable_mq":
        assert "spatial_basis.log_bandwidths" in sd and res["test_crps"] > 0
    json.load(open(out / "results.json"))

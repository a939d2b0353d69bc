"""Driver-level parity cases shared by oracle/gen_golden_driver.py (runs the UNMODIFIED reference driver in the build
container and commits what it did) and tests/test_gpu_driver_parity.py (runs this repo's driver on the same file and
compares).  Shapes follow BASELINE.json configs 1 and 2 at a size the CPU reference finishes in seconds:

  config1_shipped : configs/config_st_interp.yaml as shipped upstream -- gmm knots, learnable centres / bandwidths with
                    gradient damping and the domain penalty, multi-quantile Q=5, cosine + warm-up + progressive
                    unfreezing with ramp-up -- on a purely spatial file (T=1, like data/1a through the t=1 adapter).
  config2_default : uniform lattice, fixed basis, mean regression on a spatio-temporal file (like data/2b).

Only the epoch counts are shortened and dropout is 0 (dropout masks cannot match a CPU mt19937 stream).
"""
import os

import numpy as np


def field(x, y, t, rng):
    return (np.sin(2 * np.pi * (x + t)) * np.cos(2 * np.pi * y) + 0.5 * np.sin(6 * np.pi * x * y)
            + 0.1 * rng.standard_normal(x.shape))


def write_csv(path, S, T, seed):
    """x,y,t,z file in the layout load_kaust_csv_single reads (kaust_loader.py:19-76): t = 1..T, sites in a fixed order."""
    rng = np.random.default_rng(seed)
    xy = rng.random((S, 2))
    rows = ["x,y,t,z"]
    for ti in range(T):
        tn = ti / (T - 1) if T > 1 else 0.0
        z = field(xy[:, 0], xy[:, 1], tn, rng)
        rows += [f"{xy[s, 0]:.6f},{xy[s, 1]:.6f},{ti + 1},{z[s]:.6f}" for s in range(S)]
    with open(path, "w") as f:
        f.write("\n".join(rows) + "\n")
    return path


SHIPPED_YAML = dict(     # configs/config_st_interp.yaml of the reference, key for key
    tag="integrated", k_spatial_centers=[25, 81, 121], k_temporal_centers=[10, 15, 45],
    spatial_basis_function="wendland", spatial_init_method="gmm", spatial_learnable=True, gradient_damping=True,
    damping_threshold=0.0, damping_strength=5.0, domain_penalty_weight=0.01, movement_penalty_weight=0.0,
    sparsity_penalty_type="sparse_group", sparsity_lambda_l1=0.0, sparsity_lambda_group=0.0,
    sparsity_apply_to_spatial=True, sparsity_apply_to_temporal=False, sparsity_threshold_ratio=0.01,
    hidden_dims=[256, 256, 128], dropout=0.1, layernorm=True, p_covariates=0, obs_method="site-wise", obs_ratio=0.1,
    obs_spatial_pattern="corner", obs_spatial_intensity=10.0, split_method="random", train_ratio=0.8,
    normalize_target=False, epochs=500, lr=2e-2, basis_lr_ratio=0.05, weight_decay=5e-4, batch_size=4096, patience=50,
    grad_clip=10.0, scheduler="cosine", warmup_epochs=10, basis_unfreeze_epoch=10, basis_lr_rampup_epochs=10,
    n_experiments=50, base_seed=2025, num_workers=0, regression_type="multi-quantile",
    quantile_levels=[0.05, 0.25, 0.5, 0.75, 0.95], use_delta_reparameterization=False, device="cpu")

CASES = {
    "config1_shipped": dict(
        data=dict(S=4000, T=1, seed=11),
        config=dict(SHIPPED_YAML, dropout=0.0, obs_ratio=0.5, epochs=6, warmup_epochs=2, basis_unfreeze_epoch=2,
                    basis_lr_rampup_epochs=2)),
    "config2_default": dict(
        data=dict(S=300, T=20, seed=12),
        config=dict(SHIPPED_YAML, dropout=0.0, obs_ratio=0.5, obs_spatial_pattern="uniform", split_method="site-wise",
                    spatial_init_method="uniform", spatial_learnable=False, gradient_damping=False,
                    domain_penalty_weight=0.0, sparsity_penalty_type="none", regression_type="mean", epochs=5,
                    warmup_epochs=1, basis_unfreeze_epoch=0, basis_lr_rampup_epochs=0)),
}


def case_csv(name, directory):
    d = CASES[name]["data"]
    return write_csv(os.path.join(str(directory), f"{name}.csv"), d["S"], d["T"], d["seed"])

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


if __name__ == "__main__":
    config4()

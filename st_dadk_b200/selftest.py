"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle.

This is the one place in the package that imports `oracle/` (the checker); nothing on the product path does."""
import os
import sys

import numpy as np
import torch


def smoke():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device: st_dadk_b200 has no CPU path")
    from oracle import stdadk_oracle as orc          # checker only
    from stnf.models import STInterpMLP
    from stnf.dataio import ObservationTable
    from st_dadk_b200.trainer import Trainer
    from st_dadk_b200.predict import Predictor

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = STInterpMLP(hidden_dims=[256, 256, 128], dropout=0.1, output_dim=1)
    state = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    rng = np.random.default_rng(0)
    n = 1000
    coords = rng.random((n, 2)).astype(np.float32)
    t = (rng.integers(0, 100, n) / 99.0).astype(np.float32)
    y = (np.sin(6 * coords[:, 0]) * np.cos(4 * coords[:, 1]) + t).astype(np.float32)
    m = orc.OracleModel(
        centers=state["spatial_basis.centers"], bandwidths=state["spatial_basis._bandwidths"],
        t_centers=state["temporal_basis.centers"], t_bandwidths=state["temporal_basis.bandwidths"],
        weights=[state[f"mlp.{i}.weight"] for i in (0, 4, 8, 12)], biases=[state[f"mlp.{i}.bias"] for i in (0, 4, 8, 12)],
        ln_gamma=[state[f"mlp.{i}.weight"] for i in (1, 5, 9)], ln_beta=[state[f"mlp.{i}.bias"] for i in (1, 5, 9)],
        dropout=0.1)
    cfg = dict(lr=2e-2, weight_decay=5e-4, grad_clip=10.0, regression_type="mean")
    tr = Trainer(model, cfg, dev, batches_per_epoch=4)
    table = ObservationTable(torch.from_numpy(coords), torch.from_numpy(t), torch.from_numpy(y)).to(dev)
    perm = torch.arange(n, device=dev)
    # oracle: same dropout masks (Philox keyed on seed/step/layer/row), one forward + loss
    masks = [orc.dropout_keep_mask(n, w.shape[0], 0.1, tr.seed, 0, l) for l, w in enumerate(m.weights[:-1])]
    yref, cache = orc.forward(m, None, coords, t[:, None], train=True, keep_masks=masks, return_cache=True)
    lref, dy = orc.loss_and_grad(yref, y, "mse")
    gref = orc.backward(m, cache, dy)
    # one optimisation step in its three phases, so that the gradient can be read before the fused AdamW kernel
    # consumes and zeroes it
    tr._push_hyper()
    tr._step_compute(table, perm, 0, n, n)
    gw = tr.flat.gviews[id(model.mlp[0].weight)].cpu().numpy()
    tr._step_exchange()
    tr._step_update()
    loss = tr.pop_loss_sum()
    assert float(tr.flat.g.abs().max()) == 0.0, "smoke: the update must leave the gradient buffer zeroed"
    rel = abs(loss - lref) / abs(lref)
    gerr = np.abs(gw - gref["weights"][0]).max() / np.abs(gref["weights"][0]).max()
    assert rel < 1e-3, f"smoke: training loss {loss} vs oracle {lref}"
    assert gerr < 3e-2, f"smoke: dW1 error {gerr}"
    # prediction on a small dense grid, two shards == one shard, vs oracle at the updated weights
    state2 = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    m.weights = [state2[f"mlp.{i}.weight"] for i in (0, 4, 8, 12)]
    m.biases = [state2[f"mlp.{i}.bias"] for i in (0, 4, 8, 12)]
    m.ln_gamma = [state2[f"mlp.{i}.weight"] for i in (1, 5, 9)]
    m.ln_beta = [state2[f"mlp.{i}.bias"] for i in (1, 5, 9)]
    model.eval()
    pr = Predictor(model)
    full, _ = pr.grid(40, 30, 3)
    parts = [pr.grid(40, 30, 3, r, 2)[0] for r in range(2)]
    assert torch.equal(torch.cat(parts), full), "smoke: sharded prediction differs from the single-shard run"
    gc, gt = orc.grid_points(40, 30, 3, 0, 3600)
    yo = orc.forward(m, None, gc, gt)
    perr = np.linalg.norm(full.cpu().numpy() - yo) / np.linalg.norm(yo)
    assert perr < 1e-3, f"smoke: prediction error {perr}"
    torch.cuda.synchronize()
    print(f"[smoke] ok: train loss {loss:.6f} (oracle {lref:.6f}, rel {rel:.1e}), dW1 err {gerr:.1e}, "
          f"grid prediction rel-L2 {perr:.1e}, sharding bit-exact")

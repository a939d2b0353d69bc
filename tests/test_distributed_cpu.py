"""CPU (gloo, world_size 2): host-side logic of the multi-GPU paths -- block sharding, shared permutation,
and the data-parallel gradient identity the trainer relies on (per-rank partial gradients scaled by the GLOBAL
batch, summed with one all-reduce of the flat buffer == full-batch gradient)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import golden, oracle_from_state, state_of, orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from st_dadk_b200.trainer import epoch_permutation, shard_rows, draw_loader_base_seed
    from st_dadk_b200.predict import shard_range
    try:
        # 1. every rank draws the same permutation from the same seed (global batch = slices of ONE permutation)
        torch.manual_seed(2025)
        draw_loader_base_seed()
        perm = epoch_permutation(1000, "cpu")
        gathered = [torch.zeros_like(perm) for _ in range(world)]
        dist.all_gather(gathered, perm)
        assert all(torch.equal(g, perm) for g in gathered)
        # 2. block partition of a global batch: contiguous, disjoint, covering
        n, bs = 1000, 96
        lo, hi = shard_rows(bs, rank, world)
        edges = torch.tensor([lo, hi])
        all_edges = [torch.zeros_like(edges) for _ in range(world)]
        dist.all_gather(all_edges, edges)
        assert all_edges[0][0] == 0 and all_edges[-1][1] == bs
        assert all(int(all_edges[i][1]) == int(all_edges[i + 1][0]) for i in range(world - 1))
        assert shard_range(10_000_000, rank, world) == orc.shard_range(10_000_000, rank, world)
        # 3. data-parallel gradient identity on the oracle (what Trainer._step_body computes per rank)
        g = golden("small_mse")
        m = oracle_from_state(state_of(g))
        coords, t, y = g["coords"], g["t"], g["y"]
        N = coords.shape[0]
        b, e = shard_rows(N, rank, world)
        yh, cache = orc.forward(m, None, coords[b:e], t[b:e], return_cache=True)
        dy = 2.0 * (yh - y[b:e]) / N                       # inv_count uses the GLOBAL number of rows
        part = orc.backward(m, cache, dy)
        flat = torch.from_numpy(np.concatenate([w.reshape(-1) for w in part["weights"]] +
                                               [w.reshape(-1) for w in part["biases"]]))
        loss = torch.tensor([float(((yh - y[b:e]) ** 2).sum() / N)])
        dist.all_reduce(flat)
        dist.all_reduce(loss)
        yh_f, cache_f = orc.forward(m, None, coords, t, return_cache=True)
        lf, dyf = orc.loss_and_grad(yh_f, y, "mse")
        full = orc.backward(m, cache_f, dyf)
        ref = np.concatenate([w.reshape(-1) for w in full["weights"]] + [w.reshape(-1) for w in full["biases"]])
        np.testing.assert_allclose(flat.numpy(), ref, rtol=1e-10, atol=1e-14)
        assert abs(loss.item() - lf) < 1e-6 * abs(lf)   # float32 tensor
        # 4. dropout masks keyed by the global row: shards of the mask == the global mask
        full_mask = orc.dropout_keep_mask(N, 32, 0.1, 7, 3, 0)
        assert np.array_equal(orc.dropout_keep_mask(e - b, 32, 0.1, 7, 3, 0, row_offset=b), full_mask[b:e])
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_loader_matches_reference_semantics(tmp_path):
    """Vectorised load_kaust_csv_single: site order = first appearance, NaN fill, (T, S) layout, and the
    ObservationTable sample order = np.argwhere(mask) row-major with NaN targets dropped."""
    from stnf.dataio import load_kaust_csv_single, ObservationTable
    p = tmp_path / "d.csv"
    p.write_text("x,y,t,z\n0.5,0.5,1,1.0\n0.1,0.2,1,2.0\n0.5,0.5,2,3.0\n0.9,0.9,3,4.0\n0.1,0.2,3,5.0\n")
    z, coords, meta = load_kaust_csv_single(str(p), normalize=False)
    assert z.shape == (3, 3) and coords.dtype == np.float32
    np.testing.assert_allclose(coords, [[0.5, 0.5], [0.1, 0.2], [0.9, 0.9]])
    exp = np.array([[1, 2, np.nan], [3, np.nan, np.nan], [np.nan, 5, 4]], dtype=np.float32)
    assert np.array_equal(np.isnan(z), np.isnan(exp)) and np.allclose(np.nan_to_num(z), np.nan_to_num(exp))
    zn, _, meta = load_kaust_csv_single(str(p), normalize=True)
    vals = exp[~np.isnan(exp)]
    assert abs(meta["z_mean"] - vals.mean()) < 1e-6 and abs(meta["z_std"] - vals.std()) < 1e-6
    mask = np.ones_like(z, dtype=bool)
    tab = ObservationTable.from_mask(z, coords, mask)
    assert len(tab) == 5
    np.testing.assert_allclose(tab.y.numpy(), [1, 2, 3, 5, 4])
    np.testing.assert_allclose(tab.t.numpy(), [0, 0, 0.5, 1, 1])
    np.testing.assert_allclose(tab.coords.numpy()[3], [0.1, 0.2])
    # purely spatial file (data/1a shape): t == 1 for all rows
    p2 = tmp_path / "s.csv"
    p2.write_text('"id_train","x","y","z"\n1,0.1,0.2,1.5\n2,0.3,0.4,2.5\n')
    z2, c2, _ = load_kaust_csv_single(str(p2), normalize=False)
    assert z2.shape == (1, 2)
    assert ObservationTable.from_mask(z2, c2, np.ones_like(z2, dtype=bool)).t.tolist() == [0.0, 0.0]


def test_model_construction_matches_reference_init_on_cpu():
    """Building the module consumes the torch RNG exactly like upstream: same seed => same parameters
    (checked against moments of the reference's state_dict); state_dict keys/shapes are upstream's."""
    from stnf.models import STInterpMLP
    g = golden("default_mse")
    torch.manual_seed(0)
    model = STInterpMLP(k_spatial_centers=[25, 81, 121], k_temporal_centers=[10, 15, 45], hidden_dims=[256, 256, 128],
                        dropout=0.0, layernorm=True, output_dim=1)
    gen = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, torch.nn.LayerNorm):
                mod.weight.add_(0.2 * torch.randn(mod.weight.shape, generator=gen))
                mod.bias.add_(0.1 * torch.randn(mod.bias.shape, generator=gen))
    keys = {k[5:] for k in g.files if k.startswith("stat.")}
    assert set(model.state_dict().keys()) == keys
    for k, v in model.state_dict().items():
        st = g["stat." + k]
        assert abs(float(v.double().sum()) - st[0]) <= 1e-6 * max(1.0, abs(st[0])), k
        assert abs(float((v.double() ** 2).sum()) - st[1]) <= 1e-6 * max(1.0, abs(st[1])), k
    learn = STInterpMLP(spatial_learnable=True, output_dim=5)
    assert {"spatial_basis.centers", "spatial_basis.centers_init", "spatial_basis.log_bandwidths"} <= set(learn.state_dict())
    assert sum(p.numel() for p in learn.parameters()) == 177582
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(3, 0), torch.rand(3, 2), torch.rand(3, 1))

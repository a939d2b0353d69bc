"""CPU: pin the oracle (numpy + C restatement) against the reference's own outputs (tests/golden)."""
import numpy as np
import pytest

from helpers import golden, kat, oracle_from_state, state_of, c_spatial_basis, orc


def test_kat_basis_functions():
    k = kat()
    r = np.array(k["r"])
    np.testing.assert_allclose(orc.wendland(r), k["wendland"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(orc.gaussian(r), k["gaussian"], rtol=1e-15)
    np.testing.assert_allclose(orc.triangular(r), k["triangular"], rtol=0, atol=0)
    # the values written in SURVEY.md 8(c)
    assert orc.wendland(np.array([0.25]))[0] == pytest.approx(0.5747222900390625, abs=1e-15)
    assert orc.wendland(np.array([0.5]))[0] == pytest.approx(0.10807291666666667, abs=1e-15)


def test_knot_lattice():
    # lattice order / bandwidths exact; coordinates within 1 ulp of the reference's torch.linspace buffers
    g = golden("knots")
    ulp = 6e-8
    c, b = orc.uniform_spatial_knots([25, 81, 121])
    assert np.max(np.abs(c - g["centers"])) <= ulp and np.array_equal(b, g["bandwidths"])
    tc, tb = orc.temporal_knots([10, 15, 45])
    assert np.max(np.abs(tc - g["t_centers"])) <= ulp and np.array_equal(tb, g["t_bandwidths"])
    c2, b2 = orc.uniform_spatial_knots([16, 10000])
    assert np.max(np.abs(c2 - g["centers_16_10000"])) <= ulp and np.array_equal(b2, g["bandwidths_16_10000"])
    k9 = np.array(kat()["knots_level9"], dtype=np.float32)
    assert np.array_equal(orc.uniform_spatial_knots([9])[0], k9)   # x-major order


def test_kat_points():
    k = kat()
    kn = golden("knots")
    c, b, tc, tb = kn["centers"], kn["bandwidths"], kn["t_centers"], kn["t_bandwidths"]
    for p in k["points"]:
        phi = orc.spatial_basis(np.array([[p["x"], p["y"]]]), c, b)[0]
        psi = orc.temporal_basis(np.array([[p["t"]]]), tc, tb)[0]
        assert int((phi > 0).sum()) == p["nnz"]
        assert np.nonzero(phi > 0)[0].tolist() == p["support"]
        assert phi.sum() == pytest.approx(p["sum_phi"], rel=1e-12)
        assert int(phi.argmax()) == p["argmax"]
        assert psi.sum() == pytest.approx(p["sum_psi"], rel=1e-12)
        mask = orc.support_mask_f32(np.array([[p["x"], p["y"]]], dtype=np.float32), c, b)[0]
        assert np.nonzero(mask)[0].tolist() == p["support"]
    for fn in ("gaussian", "triangular"):
        phi = orc.spatial_basis(np.array([[0.3, 0.7]]), c, b, fn)[0]
        assert int((phi > 0).sum()) == k[fn + "_0.3_0.7"]["nnz"]
        # (0.3,0.7) sits ~1e-8 from a knot: the reference's FP64 cdist (matmul expansion) is itself
        # only ~1e-8 accurate in d there, which the non-smooth triangular basis exposes
        assert phi.sum() == pytest.approx(k[fn + "_0.3_0.7"]["sum"], rel=1e-8)


@pytest.mark.parametrize("fn", ["wendland", "gaussian", "triangular"])
def test_basis_values_vs_reference(fn):
    g = golden("basis_values")
    kn = golden("knots")
    c, b = kn["centers"], kn["bandwidths"]
    phi64 = orc.spatial_basis(g["coords"], c, b, fn)
    np.testing.assert_allclose(phi64, g[f"phi64_{fn}"], rtol=1e-9, atol=1e-7 if fn == "triangular" else 1e-12)
    # the reference's own FP32 path (cdist matmul expansion) is only ~1e-4 accurate
    np.testing.assert_allclose(phi64, g[f"phi32_{fn}"], rtol=0, atol=2e-4)
    # C restatement in FP32 direct-difference: 1e-5 relative to FP64 reference
    thetap = (b * np.float32(orc.CALIBRATION[fn])).astype(np.float32)
    phic, maskc = c_spatial_basis(g["coords"], c, thetap, fn)
    ref = g[f"phi64_{fn}"]
    # 1e-5 relative where phi is not dominated by the 1-r cancellation at the support edge
    # (FP32 ulp(r)=6e-8 bounds the absolute error), i.e. |err| <= 1e-5 * max(phi, 2e-2)
    assert np.max(np.abs(phic - ref) / np.maximum(ref, 2e-2)) < 1e-5
    # index sets: C restatement == numpy FP32 predicate == reference FP64 non-zeros
    assert np.array_equal(maskc, orc.support_mask_f32(g["coords"], c, b, fn))
    if fn != "gaussian":
        assert np.array_equal(maskc, ref > 0)
    tc, tb = kn["t_centers"], kn["t_bandwidths"]
    np.testing.assert_allclose(orc.temporal_basis(g["t"], tc, tb), g["psi64"], rtol=1e-12)


@pytest.mark.parametrize("name,fn,loss,taus", [
    ("small_mse", "wendland", "mse", None),
    ("small_mq", "wendland", "mq", [0.1, 0.5, 0.9]),
    ("small_noln_tri", "triangular", "mse", None),
    ("small_gauss", "gaussian", "mse", None),
    ("small_learnable", "wendland", "mq", [0.1, 0.5, 0.9]),
    ("small_delta", "wendland", "mq", [0.1, 0.5, 0.9]),
])
def test_forward_backward_vs_reference(name, fn, loss, taus):
    g = golden(name)
    m = oracle_from_state(state_of(g), basis_fn=fn)
    X = np.zeros((g["coords"].shape[0], 0))
    yhat, cache = orc.forward(m, X, g["coords"], g["t"], return_cache=True)
    np.testing.assert_allclose(yhat, g["yhat64"], rtol=1e-6, atol=1e-7)   # weights are FP32-stored
    np.testing.assert_allclose(yhat, g["yhat32"], rtol=0, atol=2e-3)
    lval, dy = orc.loss_and_grad(yhat, g["y"], "mse" if loss == "mse" else "pinball", taus)
    assert lval == pytest.approx(float(g["loss64"]), rel=1e-6)
    learn = name == "small_learnable"
    grads = orc.backward(m, cache, dy, coords=g["coords"], want_knot_grads=learn)
    ref = {k[5:]: g[k] for k in g.files if k.startswith("grad.")}
    pre = "mlp_trunk." if m.delta is not None else "mlp."
    lin = sorted({int(k.split(".")[1]) for k in ref if k.startswith(pre) and ref[k].ndim == 2})
    for li, k in enumerate(lin):
        np.testing.assert_allclose(grads["weights"][li], ref[f"{pre}{k}.weight"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(grads["biases"][li], ref[f"{pre}{k}.bias"], rtol=1e-5, atol=1e-9)
    lns = sorted({int(k.split(".")[1]) for k in ref if k.startswith(pre) and k.endswith("weight") and ref[k].ndim == 1})
    for li, k in enumerate(lns):
        np.testing.assert_allclose(grads["ln_gamma"][li], ref[f"{pre}{k}.weight"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(grads["ln_beta"][li], ref[f"{pre}{k}.bias"], rtol=1e-5, atol=1e-9)
    if m.delta is not None:
        for j, d in enumerate(grads["delta"]):
            np.testing.assert_allclose(d, ref[f"delta_params.{j}"], rtol=1e-5, atol=1e-9)
    if learn:
        np.testing.assert_allclose(grads["centers"], ref["spatial_basis.centers"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(grads["log_bandwidths"], ref["spatial_basis.log_bandwidths"], rtol=1e-5, atol=1e-9)


def test_philox_known_answer():
    # Random123 known-answer vectors for Philox4x32-10
    out = orc.philox4x32(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = orc.philox4x32(np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff),
                         0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    keep = orc.dropout_keep_mask(4096, 256, 0.1, seed=7, step=3, layer=1)
    assert abs(keep.mean() - 0.9) < 5e-3
    # sharding invariance: rows [100,200) of the global mask == mask drawn with row_offset=100
    part = orc.dropout_keep_mask(100, 256, 0.1, seed=7, step=3, layer=1, row_offset=100)
    assert np.array_equal(keep[100:200], part)


def test_adamw_ema_vs_torch():
    import torch
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(1000)
    p = torch.nn.Parameter(torch.tensor(p0, dtype=torch.float64))
    opt = torch.optim.AdamW([p], lr=2e-2, weight_decay=5e-4)
    pn, m, v, sh = p0.copy(), np.zeros(1000), np.zeros(1000), p0.copy()
    sh_t = p.data.clone()
    for step in range(1, 6):
        gnp = rng.standard_normal(1000)
        p.grad = torch.tensor(gnp, dtype=torch.float64)
        opt.step()
        sh_t = 0.99 * sh_t + 0.01 * p.data
        orc.adamw_ema_step(pn, gnp, m, v, sh, step, 2e-2, 5e-4, 0.99)
    np.testing.assert_allclose(pn, p.detach().numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(sh, sh_t.numpy(), rtol=1e-12, atol=1e-14)


def test_shard_and_grid():
    n = 10_000_000
    for world in (1, 2, 4, 8, 3):
        edges = [orc.shard_range(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    c, t = orc.grid_points(1000, 1000, 10, 999_998, 1_000_003)
    assert c[0].tolist() == [np.float32(999 / 999), np.float32(998 / 999)] and t[0, 0] == 0
    assert c[2].tolist() == [0.0, 0.0] and t[2, 0] == np.float32(1 / 9)
    assert orc.shard_range(0, 0, 4) == (0, 0)

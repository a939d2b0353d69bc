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


# ---------------------------------------------------------------------------------------------------------------------
# Size-independent properties (the ones the GPU tests rely on at full size, where no oracle run is affordable)
# ---------------------------------------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as hs   # noqa: E402


@settings(max_examples=60, deadline=None)
@given(n=hs.integers(0, 10**9), world=hs.integers(1, 64))
def test_property_shard_ranges_partition_in_order(n, world):
    """Contiguous block sharding: shards tile [0, n) in rank order and their sizes differ by at most one."""
    edges = [orc.shard_range(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    sizes = []
    for (b, e), nxt in zip(edges, edges[1:] + [(n, n)]):
        assert b <= e == nxt[0]
        sizes.append(e - b)
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=25, deadline=None)
@given(rows=hs.integers(1, 300), cols=hs.integers(1, 70), lo=hs.integers(0, 2**33), step=hs.integers(0, 2**31 - 1),
       layer=hs.integers(0, 3), seed=hs.integers(0, 2**64 - 1), p=hs.sampled_from([0.05, 0.1, 0.5]))
def test_property_dropout_mask_is_keyed_by_global_row(rows, cols, lo, step, layer, seed, p):
    """A shard that starts at global row `lo` draws exactly the rows [lo, lo+rows) of the unsharded mask (rows beyond
    2^32 included), and a different layer / step gives a different stream."""
    full = orc.dropout_keep_mask(rows + 7, cols, p, seed, step, layer, row_offset=lo)
    part = orc.dropout_keep_mask(rows, cols, p, seed, step, layer, row_offset=lo + 7)
    assert np.array_equal(full[7:], part)
    if rows * cols >= 256:
        assert not np.array_equal(part, orc.dropout_keep_mask(rows, cols, p, seed, step + 1, layer, row_offset=lo + 7))
        assert not np.array_equal(part, orc.dropout_keep_mask(rows, cols, p, seed, step, layer + 1, row_offset=lo + 7))


@settings(max_examples=30, deadline=None)
@given(nx=hs.integers(1, 60), ny=hs.integers(1, 60), nt=hs.integers(1, 12), a=hs.integers(0, 10**4), b=hs.integers(0, 10**4))
def test_property_grid_generator_is_a_window_of_the_full_grid(nx, ny, nt, a, b):
    """grid_points(begin, end) is the [begin, end) window of the full grid in x-major order (any sharding of a dense
    grid concatenates to the whole), coordinates stay in [0, 1] and the end points are exact."""
    n = nx * ny * nt
    begin, end = sorted((a % (n + 1), b % (n + 1)))
    cf, tf = orc.grid_points(nx, ny, nt, 0, n)
    c, t = orc.grid_points(nx, ny, nt, begin, end)
    assert np.array_equal(c, cf[begin:end]) and np.array_equal(t, tf[begin:end])
    assert cf.min() >= 0.0 and cf.max() <= 1.0 and tf.min() >= 0.0 and tf.max() <= 1.0
    assert tuple(cf[0]) == (0.0, 0.0) and tf[0, 0] == 0.0
    assert cf[-1, 0] == (1.0 if nx > 1 else 0.0) and cf[-1, 1] == (1.0 if ny > 1 else 0.0)
    assert tf[-1, 0] == (1.0 if nt > 1 else 0.0)


@settings(max_examples=50, deadline=None)
@given(hs.lists(hs.floats(float(np.float32(-1e30)), float(np.float32(1e30)), allow_nan=False, width=32), min_size=1,
                max_size=64))
def test_property_tf32_rounding(xs):
    """tf32_round: idempotent, monotone, sign-symmetric, within half a TF32 ulp (2^-11 relative) of its argument."""
    x = np.asarray(xs, dtype=np.float32)
    r = orc.tf32_round(x)
    assert np.array_equal(orc.tf32_round(r), r)
    assert np.array_equal(orc.tf32_round(-x), -r)
    normal = np.abs(x) > 1e-30
    assert np.all(np.abs(r[normal].astype(np.float64) - x[normal]) <= np.abs(x[normal].astype(np.float64)) * 2.0 ** -11)
    order = np.argsort(x, kind="stable")
    assert np.all(np.diff(r[order]) >= 0)
    assert np.all((r.view(np.uint32) & np.uint32(0x1FFF)) == 0)


@settings(max_examples=20, deadline=None)
@given(seed=hs.integers(0, 2**32 - 1), fn=hs.sampled_from(["wendland", "triangular"]))
def test_property_support_sets_and_basis_values_agree(seed, fn):
    """phi > 0 only inside the FP32 support predicate the kernels use, phi <= 1, phi == 1 exactly at a knot, and the
    predicate is symmetric under reflection of the lattice (x -> 1 - x)."""
    rng = np.random.default_rng(seed)
    c, bw = orc.uniform_spatial_knots([9, 25])
    pts = rng.random((64, 2)).astype(np.float32)
    phi = orc.spatial_basis(pts, c, bw, fn, np.float64)
    mask = orc.support_mask_f32(pts, c, bw, fn)
    assert np.all(phi[~mask] <= 1e-12) and np.all(phi <= 1.0 + 1e-12) and np.all(phi >= 0.0)
    at_knots = orc.spatial_basis(c, c, bw, fn, np.float64)
    assert np.allclose(np.diag(at_knots), 1.0, atol=1e-12)
    assert orc.support_mask_f32(c, c, bw, fn).diagonal().all()

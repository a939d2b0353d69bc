"""GPU: the space-time FIELD kernel (stdadk_predict_field: basis + block 1 once per tile of sites, loop over time)
against the generic per-point kernel, the oracle and itself under sharding."""
import numpy as np
import pytest
import torch

from helpers import orc
from test_gpu_kernels import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(q=1, hidden=(256, 256, 128), fn="wendland", ln=True, seed=3):
    from stnf.models import STInterpMLP
    from test_gpu_model import _perturb_ln
    torch.manual_seed(seed)
    m = STInterpMLP(hidden_dims=list(hidden), dropout=0.0, layernorm=ln, output_dim=q, spatial_basis_function=fn)
    if ln:
        _perturb_ln(m, seed)
    return m.to(DEV).eval()


def _oracle(model, coords, t):
    st = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    from helpers import oracle_from_state
    m = oracle_from_state(st, basis_fn=model.spatial_basis_function)
    return orc.forward(m, None, coords, t)


@pytest.mark.parametrize("q,hidden,fn,ln,S,T", [(1, (256, 256, 128), "wendland", True, 1000, 7),
                                               (5, (256, 256, 128), "wendland", True, 333, 3),
                                               (3, (64, 32), "triangular", False, 130, 12),
                                               (1, (128,), "gaussian", True, 257, 1),
                                               (2, (256, 160, 96, 32), "wendland", True, 500, 5)])
def test_field_kernel_matches_generic_kernel_and_oracle(q, hidden, fn, ln, S, T):
    from st_dadk_b200.predict import Predictor
    model = _model(q, hidden, fn, ln)
    g = torch.Generator().manual_seed(1)
    sites = torch.rand(S, 2, generator=g).to(DEV)
    pr = Predictor(model)
    field, (b, e) = pr.space_time_field(sites, T)
    assert pr.used_field_kernel and (b, e) == (0, S * T)
    gen = Predictor(model)
    gen.use_field_kernel = False
    ref, _ = gen.space_time_field(sites, T)
    assert not gen.used_field_kernel
    # same function; the two kernels differ in the order of FP32 additions and in zt being FP32 instead of a TF32 GEMM
    # (each is one TF32 pass: ~4e-4 from the FP64 oracle, so up to ~1e-3 from each other)
    assert rel_l2(field.cpu().numpy(), ref.cpu().numpy()) < 1e-3
    coords = sites.repeat(T, 1).cpu().numpy()
    t = (torch.arange(T).repeat_interleave(S).float() / max(T - 1, 1)).numpy()[:, None]
    want = _oracle(model, coords, t)
    print("field vs oracle", rel_l2(field.cpu().numpy(), want), "generic vs oracle", rel_l2(ref.cpu().numpy(), want))
    assert rel_l2(field.cpu().numpy(), want) < 1e-3
    # sharding: any cut of the (t, s) row-major field gives the same bits (ragged shards = partial time steps)
    for world in (2, 3, 7):
        parts = [pr.space_time_field(sites, T, r, world)[0] for r in range(world)]
        assert torch.equal(torch.cat(parts), field), world
    # sharded by site: the same bits as a (T, S, Q) field
    for world in (1, 3):
        parts = [pr.space_time_field_by_sites(sites, T, r, world)[0] for r in range(world)]
        assert pr.used_field_kernel
        assert torch.equal(torch.cat(parts, dim=1), field.view(T, S, q)), world
    parts = [gen.space_time_field_by_sites(sites, T, r, 2)[0] for r in range(2)]
    assert torch.equal(torch.cat(parts, dim=1), ref.view(T, S, q))


def test_field_kernel_grid_equals_explicit_sites_and_shards():
    from st_dadk_b200.predict import Predictor
    model = _model(q=2)
    pr = Predictor(model)
    nx, ny, nt = 37, 23, 4
    out, _ = pr.grid(nx, ny, nt)
    assert pr.used_field_kernel
    ii, jj = torch.meshgrid(torch.arange(nx), torch.arange(ny), indexing="ij")
    sites = torch.stack([ii.reshape(-1).float() / (nx - 1), jj.reshape(-1).float() / (ny - 1)], dim=1).to(DEV)
    pr2 = Predictor(model)
    f2, _ = pr2.space_time_field(sites, nt)
    assert rel_l2(out.cpu().numpy(), f2.cpu().numpy()) < 1e-6
    gen = Predictor(model)
    gen.use_field_kernel = False
    ref, _ = gen.grid(nx, ny, nt)
    assert rel_l2(out.cpu().numpy(), ref.cpu().numpy()) < 1e-3
    for world in (2, 5, 8):
        parts = [pr.grid(nx, ny, nt, r, world)[0] for r in range(world)]
        assert torch.equal(torch.cat(parts), out), world
    # sharded by SITE (every rank: its sites at all time steps): the same bits, as an (nt, S, Q) field
    field = out.view(nt, nx * ny, 2)
    for world in (1, 3, 8):
        parts = [pr.grid_by_sites(nx, ny, nt, r, world) for r in range(world)]
        assert parts[0][1][0] == 0 and parts[-1][1][1] == nx * ny
        assert all(parts[i][1][1] == parts[i + 1][1][0] for i in range(world - 1))
        assert torch.equal(torch.cat([p[0] for p in parts], dim=1), field), world
    # ... and through the generic kernels (field kernel off)
    parts = [gen.grid_by_sites(nx, ny, nt, r, 3)[0] for r in range(3)]
    assert torch.equal(torch.cat(parts, dim=1), ref.view(nt, nx * ny, 2))


def test_full_size_10M_point_grid_sample_vs_oracle_and_world8_shards():
    """BASELINE config 3 at its real size: the 1000 x 1000 x 10 grid generated on the device (10,000,000 points).  A
    4,096-row sample (plus corner / boundary rows) is checked against the oracle; the 8-way shardings -- contiguous
    blocks of the row index and by site -- reproduce the single-launch field bit for bit."""
    from st_dadk_b200.predict import Predictor
    model = _model(q=1)
    pr = Predictor(model, static_weights=True)
    nx, ny, nt = 1000, 1000, 10
    S, n = nx * ny, nx * ny * nt
    out, (b, e) = pr.grid(nx, ny, nt)
    assert pr.used_field_kernel and (b, e) == (0, n) and bool(torch.isfinite(out).all())
    rng = np.random.default_rng(5)
    rows = np.unique(np.concatenate([rng.integers(0, n, 4096), [0, ny - 1, S - 1, S, n - S, n - 1, 1_250_000, 1_249_999]]))
    k, site = rows // S, rows % S
    i, j = site // ny, site % ny
    f32 = np.float32          # the grid generator's arithmetic: FP32 division of the FP32 indices
    coords = np.stack([i.astype(f32) / f32(nx - 1), j.astype(f32) / f32(ny - 1)], axis=1)
    t = (k.astype(f32) / f32(nt - 1))[:, None]
    assert coords.dtype == np.float32 and t.dtype == np.float32
    want = _oracle(model, coords, t)
    got = out[torch.from_numpy(rows).to(DEV)].cpu().numpy()
    assert rel_l2(got, want) < 1e-3
    parts = [pr.grid(nx, ny, nt, r, 8) for r in range(8)]
    assert parts[0][1][0] == 0 and parts[-1][1][1] == n and all(parts[r][1][1] == parts[r + 1][1][0] for r in range(7))
    assert torch.equal(torch.cat([p[0] for p in parts]), out)
    del parts
    parts = [pr.grid_by_sites(nx, ny, nt, r, 8) for r in range(8)]
    assert torch.equal(torch.cat([p[0] for p in parts], dim=1), out.view(nt, S, 1))

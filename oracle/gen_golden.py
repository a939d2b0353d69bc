"""Generate tests/golden/* by running the REAL reference (imported from /root/reference).

Run in the build container only:  python oracle/gen_golden.py
The GPU box has no /root/reference; tests read the committed fixtures.  Every fixture
is produced by the unmodified upstream `stnf` package (FP64 via `.double()` and FP32).
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("STDADK_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from stnf.models.st_interp import STInterpMLP, SpatialBasisEmbedding, TemporalBasisEmbedding  # noqa: E402
from stnf.utils.ema import ModelEMA  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(4)


def pinball_multi(y_pred, y, taus):
    # scripts/train_st_interp.py:37-50, 625-631
    losses = []
    for qi, q in enumerate(taus):
        e = y - y_pred[:, qi:qi + 1]
        losses.append(torch.mean(torch.max((q - 1) * e, q * e)))
    return torch.mean(torch.stack(losses))


def kat():
    """Known-answer values of SURVEY.md 8(c), regenerated from the reference in FP64."""
    sb = SpatialBasisEmbedding([25, 81, 121]).double()
    tb = TemporalBasisEmbedding([10, 15, 45]).double()
    r = torch.tensor([0.0, 0.25, 0.5, 0.75, 1.0, 1.5], dtype=torch.float64)
    out = {"r": r.tolist(), "wendland": sb._wendland(r).tolist(), "gaussian": sb._gaussian(r).tolist(),
           "triangular": sb._triangular(r).tolist(), "points": []}
    for (x, y, t) in [(0.3, 0.7, 0.37), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (0.5, 0.5, 0.5),
                      (0.123456, 0.987654, 0.25)]:
        c = torch.tensor([[x, y]], dtype=torch.float64)
        phi = sb(c)[0]
        psi = tb(torch.tensor([[t]], dtype=torch.float64))[0]
        out["points"].append({"x": x, "y": y, "t": t, "nnz": int((phi > 0).sum()), "sum_phi": float(phi.sum()),
                              "argmax": int(phi.argmax()), "max_phi": float(phi.max()),
                              "sum_psi": float(psi.sum()), "support": torch.nonzero(phi > 0).flatten().tolist()})
    for fn in ("gaussian", "triangular"):
        s = SpatialBasisEmbedding([25, 81, 121], basis_function=fn).double()
        phi = s(torch.tensor([[0.3, 0.7]], dtype=torch.float64))[0]
        out[fn + "_0.3_0.7"] = {"nnz": int((phi > 0).sum()), "sum": float(phi.sum())}
    s9 = SpatialBasisEmbedding([9])
    out["knots_level9"] = s9.centers.tolist()
    out["param_counts"] = {
        "q1": sum(p.numel() for p in STInterpMLP().parameters()),
        "q5": sum(p.numel() for p in STInterpMLP(output_dim=5).parameters()),
        "q5_learnable": sum(p.numel() for p in STInterpMLP(output_dim=5, spatial_learnable=True).parameters()),
    }
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(out, f, indent=1)


def knots():
    """Knot buffers exactly as torch.linspace/meshgrid build them (bit-exact check of the lattice)."""
    sb = SpatialBasisEmbedding([25, 81, 121])
    tb = TemporalBasisEmbedding([10, 15, 45])
    sb2 = SpatialBasisEmbedding([16, 10000])
    np.savez_compressed(os.path.join(OUT, "knots.npz"), centers=sb.centers.numpy(), bandwidths=sb._bandwidths.numpy(),
                        t_centers=tb.centers.numpy(), t_bandwidths=tb.bandwidths.numpy(),
                        centers_16_10000=sb2.centers.numpy(), bandwidths_16_10000=sb2._bandwidths.numpy())


def basis_values():
    rng = np.random.default_rng(7)
    coords = rng.random((97, 2)).astype(np.float32)
    coords[:5] = [[0, 0], [1, 1], [0.5, 0.5], [0.25, 0.75], [1, 0]]
    t = rng.random((97, 1)).astype(np.float32)
    out = {"coords": coords, "t": t}
    for fn in ("wendland", "gaussian", "triangular"):
        sb = SpatialBasisEmbedding([25, 81, 121], basis_function=fn)
        out[f"phi32_{fn}"] = sb(torch.from_numpy(coords)).numpy()
        out[f"phi64_{fn}"] = sb.double()(torch.from_numpy(coords).double()).numpy()
    tb = TemporalBasisEmbedding([10, 15, 45])
    out["psi32"] = tb(torch.from_numpy(t)).numpy()
    out["psi64"] = tb.double()(torch.from_numpy(t).double()).numpy()
    np.savez_compressed(os.path.join(OUT, "basis_values.npz"), **out)


def state_to_np(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def model_case(name, ctor_kwargs, n, loss_type, taus, seed, store_state=True, learnable_grads=False):
    """Forward, loss and every parameter gradient of the reference module in FP64 (+ FP32 forward)."""
    torch.manual_seed(seed)
    m32 = STInterpMLP(**ctor_kwargs)
    # LayerNorm affine and biases away from their trivial init so that the test sees them
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mod in m32.modules():
            if isinstance(mod, torch.nn.LayerNorm):
                mod.weight.add_(0.2 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.add_(0.1 * torch.randn(mod.bias.shape, generator=g))
    rng = np.random.default_rng(seed)
    coords = rng.random((n, 2)).astype(np.float32)
    t = (rng.integers(0, 100, size=(n, 1)) / 99.0).astype(np.float32)
    y = rng.standard_normal((n, 1)).astype(np.float32)
    X = np.zeros((n, 0), dtype=np.float32)
    m32.eval()
    with torch.no_grad():
        y32 = m32(torch.from_numpy(X), torch.from_numpy(coords), torch.from_numpy(t)).numpy()
    import copy
    m64 = copy.deepcopy(m32).double()
    m64.eval()  # dropout off; gradients still flow
    yp = m64(torch.from_numpy(X).double(), torch.from_numpy(coords).double(), torch.from_numpy(t).double())
    yt = torch.from_numpy(y).double()
    if loss_type == "mse":
        loss = torch.nn.functional.mse_loss(yp, yt)
    else:
        loss = pinball_multi(yp, yt, taus)
    loss.backward()
    out = {"coords": coords, "t": t, "y": y, "yhat32": y32, "yhat64": yp.detach().numpy(),
           "loss64": np.float64(loss.item())}
    for k, p in m64.named_parameters():
        gnp = p.grad.numpy()
        if store_state:
            out["grad." + k] = gnp
        else:  # big model: keep fixtures small -- moments and a strided sample of each gradient
            out["gstat." + k] = np.array([gnp.sum(), (gnp ** 2).sum()])
            out["gsample." + k] = gnp.reshape(-1)[::97].copy()
    if store_state:
        for k, v in state_to_np(m32).items():
            out["state." + k] = v
    else:
        for k, v in state_to_np(m32).items():
            out["stat." + k] = np.array([v.astype(np.float64).sum(), (v.astype(np.float64) ** 2).sum()])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def train_curve():
    """20 reference training steps (MSE, dropout 0, AdamW + clip + EMA, lr warmup written after the
    step as in train_st_interp.py:709-718) on seeded synthetic data; FP32 CPU reference."""
    torch.manual_seed(123)
    kw = dict(k_spatial_centers=[25, 81, 121], k_temporal_centers=[10, 15, 45], hidden_dims=[256, 256, 128],
              dropout=0.0, layernorm=True)
    model = STInterpMLP(**kw)
    rng = np.random.default_rng(2025)
    n, bs, steps = 2048, 512, 20
    coords = rng.random((n, 2)).astype(np.float32)
    t = (rng.integers(0, 100, size=(n, 1)) / 99.0).astype(np.float32)
    y = (np.sin(2 * np.pi * (coords[:, :1] + t)) * np.cos(2 * np.pi * coords[:, 1:2])
         + 0.5 * np.sin(6 * np.pi * coords[:, :1] * coords[:, 1:2])).astype(np.float32)
    lr, wd, clip = 2e-2, 5e-4, 10.0
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd)
    bpe = n // bs
    ema = ModelEMA(model, decay=1.0 - 1.0 / (10.0 * bpe))
    warmup_steps = 8
    losses, norms = [], []
    ct, tt, yt = torch.from_numpy(coords), torch.from_numpy(t), torch.from_numpy(y)
    X = torch.zeros(bs, 0)
    model.train()
    for s in range(steps):
        lo = (s % bpe) * bs
        opt.zero_grad()
        pred = model(X, ct[lo:lo + bs], tt[lo:lo + bs])
        loss = torch.nn.functional.mse_loss(pred, yt[lo:lo + bs])
        loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), clip)))
        opt.step()
        ema.update(model)
        if s < warmup_steps:
            for gq in opt.param_groups:
                gq["lr"] = lr * (s + 1) / warmup_steps
        losses.append(loss.item())
    ema.apply_shadow()
    model.eval()
    with torch.no_grad():
        yhat_ema = model(torch.zeros(256, 0), ct[:256], tt[:256]).numpy()
    ema.restore()
    with torch.no_grad():
        yhat_raw = model(torch.zeros(256, 0), ct[:256], tt[:256]).numpy()
    np.savez_compressed(os.path.join(OUT, "train_curve.npz"), coords=coords, t=t, y=y,
                        losses=np.array(losses), grad_norms=np.array(norms), yhat_ema=yhat_ema, yhat_raw=yhat_raw,
                        meta=np.array([n, bs, steps, warmup_steps]), hyper=np.array([lr, wd, clip]))


if __name__ == "__main__":
    kat()
    knots()
    basis_values()
    small = dict(k_spatial_centers=[9, 25], k_temporal_centers=[5, 7], hidden_dims=[32, 16], dropout=0.0,
                 layernorm=True)
    model_case("small_mse", dict(small, output_dim=1), 37, "mse", None, 11)
    model_case("small_mq", dict(small, output_dim=3), 37, "mq", [0.1, 0.5, 0.9], 12)
    model_case("small_noln_tri", dict(small, layernorm=False, output_dim=1, spatial_basis_function="triangular"),
               37, "mse", None, 13)
    model_case("small_gauss", dict(small, output_dim=1, spatial_basis_function="gaussian"), 37, "mse", None, 14)
    model_case("small_learnable", dict(small, output_dim=3, spatial_learnable=True), 61, "mq", [0.1, 0.5, 0.9], 15)
    model_case("small_delta", dict(small, output_dim=3, use_delta_reparameterization=True), 37, "mq",
               [0.1, 0.5, 0.9], 16)
    default = dict(k_spatial_centers=[25, 81, 121], k_temporal_centers=[10, 15, 45], hidden_dims=[256, 256, 128],
                   dropout=0.0, layernorm=True)
    model_case("default_mse", dict(default, output_dim=1), 300, "mse", None, 0, store_state=False)
    model_case("default_mq5", dict(default, output_dim=5), 300, "mq", [0.05, 0.25, 0.5, 0.75, 0.95], 1,
               store_state=False)
    train_curve()
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def masks_fixture():
    """Observation / split masks and loss known-answers from the reference's scripts/train_st_interp.py
    (matplotlib is absent in the container: a two-attribute stub satisfies its module-level imports)."""
    import types
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.use = lambda *a, **k: None
            sys.modules[name] = mod
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_train", os.path.join(REF, "scripts", "train_st_interp.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(5)
    T, S = 7, 60
    coords = rng.random((S, 2)).astype(np.float32)
    z = rng.standard_normal((T, S)).astype(np.float32)
    z[2, 5] = np.nan
    out = {"coords": coords, "z": z}
    for tag, (om, pat, sm) in {"a": ("site-wise", "corner", "random"), "b": ("random", "uniform", "site-wise"),
                               "c": ("site-wise", "uniform", "site-wise"), "d": ("random", "corner", "random")}.items():
        fn = ref.create_spatial_obs_prob_fn(pat, 10.0)
        obs, sites = ref.sample_observations(z, coords, om, 0.3, fn, seed=2025)
        tr, va = ref.split_train_valid(obs, sites, sm, 0.8, seed=12025)
        ds = ref.create_dataset_from_mask(z, coords, tr, 0)
        out[f"{tag}_obs"], out[f"{tag}_sites"], out[f"{tag}_train"], out[f"{tag}_valid"] = obs, np.asarray(sites), tr, va
        out[f"{tag}_ds_y"] = np.array([float(s["y"]) for s in ds], dtype=np.float32)
        out[f"{tag}_ds_t"] = np.array([float(s["t"]) for s in ds], dtype=np.float32)
        out[f"{tag}_ds_c"] = np.stack([s["coords"].numpy() for s in ds]) if ds else np.zeros((0, 2), np.float32)
    yp = torch.tensor(rng.standard_normal((11, 3)))
    yt = torch.tensor(rng.standard_normal((11, 1)))
    out["loss_yp"], out["loss_yt"] = yp.numpy(), yt.numpy()
    out["pinball_03"] = ref.quantile_loss(yp[:, :1], yt, 0.3).item()
    out["nc_p1"] = ref.non_crossing_penalty(yp, "mean", 1).item()
    out["nc_p2"] = ref.non_crossing_penalty(yp, "sum", 2).item()
    deltas = [torch.tensor(rng.standard_normal(6)) for _ in range(4)]
    out["deltas"] = np.stack([d.numpy() for d in deltas])
    out["pnc"] = ref.compute_p_nc_delta_penalty(deltas).item()
    out["crps"] = ref.compute_crps_multi_quantile(yp.numpy(), yt.numpy(), [0.1, 0.5, 0.9])
    out["auto_bs"] = np.array([4096, 8000, 512, 4096, 80000, 4096])   # (config bs, n_train, expected) pairs
    np.savez_compressed(os.path.join(OUT, "masks_losses.npz"), **out)


if __name__ == "__main__":
    masks_fixture()
    print("masks_losses.npz", os.path.getsize(os.path.join(OUT, "masks_losses.npz")))

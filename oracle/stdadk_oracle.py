"""CPU oracle for the ST-DADK hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference algorithm for the path
`(x, y, t) -> multi-resolution basis -> MLP -> loss -> gradients -> AdamW/EMA`.
It is the checker for the CUDA kernels.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; the
product path (`st_dadk_b200`, `stnf`) never does and has no CPU fallback.

Parity status: PINNED.  `oracle/gen_golden.py` imports the real reference module
from /root/reference in the build container, runs it in FP64 and FP32 on seeded
inputs and commits the outputs under `tests/golden/`; `tests/test_oracle.py`
checks every function below against those vectors and against the known-answer
values listed in SURVEY.md section 8(c).

Reference citations are `path:line` under the upstream repo root.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# stnf/models/st_interp.py:56-60
CALIBRATION = {"wendland": 1.0, "gaussian": 0.223477, "triangular": 0.654714}
BASIS_CODE = {"wendland": 0, "gaussian": 1, "triangular": 2}


# ----------------------------------------------------------------------------
# Knot lattices
# ----------------------------------------------------------------------------
def linspace01_f32(n: int) -> np.ndarray:
    """linspace(0, 1, n) in float32 by ATen's scalar rule: step=(end-start)/(n-1) in float,
    first half start+step*i, second half end-step*(n-1-i) (st_interp.py:162-163, 567).

    ATen's vectorised CPU fill (SIMD width and FMA contraction depend on the host: AVX2 vs
    AVX-512) differs from this by at most 1 ulp on a few entries, so knot coordinates are treated
    as DATA: the product module builds them with the same torch.linspace call the reference
    makes, the kernels read them from the buffers, and this helper is pinned to the reference's
    buffers to 1 ulp in tests/test_oracle.py.
    """
    if n == 1:
        return np.zeros(1, dtype=np.float32)
    f = np.float32
    step = f(f(1.0) / f(n - 1))
    out = np.empty(n, dtype=np.float32)
    half = n // 2
    for i in range(n):
        out[i] = f(step * f(i)) if i < half else f(f(1.0) - f(step * f(n - 1 - i)))
    return out


def uniform_spatial_knots(n_centers: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """Regular lattices, x-major within a level (st_interp.py:152-185).

    knot j of a level with `side` points per axis sits at (lin[j // side], lin[j % side]);
    bandwidth is 2.5 grid spacings; levels are concatenated in list order.
    """
    cs, bs = [], []
    for k in n_centers:
        side = int(math.sqrt(k))
        if side * side != k:
            raise ValueError(f"n_centers must be perfect squares, got {k}")
        lin = linspace01_f32(side)
        ix, iy = np.divmod(np.arange(k), side)
        cs.append(np.stack([lin[ix], lin[iy]], axis=1))
        spacing = 1.0 / (side - 1) if side > 1 else 1.0
        bs.append(np.full(k, np.float32(2.5 * spacing), dtype=np.float32))
    return np.concatenate(cs).astype(np.float32), np.concatenate(bs).astype(np.float32)


def temporal_knots(n_centers: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """1-D knots linspace(0,1,n) per level, bandwidth 2.5/(n-1) (st_interp.py:557-581)."""
    cs, bs = [], []
    for n in n_centers:
        cs.append(linspace01_f32(n))
        spacing = 1.0 / (n - 1) if n > 1 else 1.0
        bs.append(np.full(n, np.float32(2.5 * spacing), dtype=np.float32))
    return np.concatenate(cs).astype(np.float32), np.concatenate(bs).astype(np.float32)


# ----------------------------------------------------------------------------
# Basis functions
# ----------------------------------------------------------------------------
def wendland(r):
    """(1-r)^6 (35 r^2 + 18 r + 3)/3 on r<1, else 0 (st_interp.py:462-471)."""
    r = np.minimum(r, 1.0)
    return (1.0 - r) ** 6 * (35.0 * r * r + 18.0 * r + 3.0) / 3.0


def wendland_dr(r):
    """d/dr of `wendland`: -(56/3) r (5r+1) (1-r)^5 for r<1 else 0."""
    inside = r < 1.0
    rr = np.where(inside, r, 0.0)
    return np.where(inside, -(56.0 / 3.0) * rr * (5.0 * rr + 1.0) * (1.0 - rr) ** 5, 0.0)


def gaussian(r):
    """exp(-r^2/2) (st_interp.py:473-481)."""
    return np.exp(-0.5 * r * r)


def gaussian_dr(r):
    return -r * np.exp(-0.5 * r * r)


def triangular(r):
    """(1-r)_+ (st_interp.py:483-491)."""
    return np.maximum(1.0 - r, 0.0)


def triangular_dr(r):
    return np.where(r < 1.0, -1.0, 0.0)


_PHI = {"wendland": (wendland, wendland_dr), "gaussian": (gaussian, gaussian_dr),
        "triangular": (triangular, triangular_dr)}


def spatial_basis(coords, centers, bandwidths, fn: str = "wendland", dtype=np.float64):
    """phi (N, K_s): direct-difference distance, r = d / (theta * calib) (st_interp.py:433-460).

    The reference takes torch.cdist's matmul-expansion path in FP32 (error ~1e-4 in phi);
    the oracle evaluates the same formula by direct differences, which in FP64 is the
    reference module run with `.double()` to ~1e-15 (pinned in tests/test_oracle.py).
    """
    coords = np.asarray(coords, dtype=dtype)
    centers = np.asarray(centers, dtype=dtype)
    thetap = np.asarray(bandwidths, dtype=dtype) * dtype(CALIBRATION[fn])
    dx = coords[:, None, 0] - centers[None, :, 0]
    dy = coords[:, None, 1] - centers[None, :, 1]
    d = np.sqrt(dx * dx + dy * dy)
    r = d / thetap[None, :]
    return _PHI[fn][0](r).astype(dtype)


def support_mask_f32(coords, centers, bandwidths, fn: str = "wendland") -> np.ndarray:
    """The support predicate the kernels use: d^2 < theta'^2 in FP32, no FMA contraction.

    d2 = fl(fl(dx*dx) + fl(dy*dy)), theta'2 = fl(theta' * theta'), theta' = fl(bw * fl(calib)).
    Bit-exact between this function, oracle/basis_ref.c and the CUDA kernels.
    Gaussian is not compact: every knot is in the support.
    """
    coords = np.asarray(coords, dtype=np.float32)
    centers = np.asarray(centers, dtype=np.float32)
    if fn == "gaussian":
        return np.ones((coords.shape[0], centers.shape[0]), dtype=bool)
    thetap = (np.asarray(bandwidths, dtype=np.float32) * np.float32(CALIBRATION[fn])).astype(np.float32)
    th2 = (thetap * thetap).astype(np.float32)
    dx = (coords[:, None, 0] - centers[None, :, 0]).astype(np.float32)
    dy = (coords[:, None, 1] - centers[None, :, 1]).astype(np.float32)
    d2 = ((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)).astype(np.float32)
    return d2 < th2[None, :]


def temporal_basis(t, centers, bandwidths, dtype=np.float64):
    """psi (N, K_t) = exp(-0.5 ((t - c)/bw)^2) (st_interp.py:583-596)."""
    t = np.asarray(t, dtype=dtype).reshape(-1, 1)
    c = np.asarray(centers, dtype=dtype).reshape(1, -1)
    b = np.asarray(bandwidths, dtype=dtype).reshape(1, -1)
    s = (t - c) / b
    return np.exp(-0.5 * s * s).astype(dtype)


# ----------------------------------------------------------------------------
# Counter-based dropout RNG shared with the kernels (Philox4x32-10)
# ----------------------------------------------------------------------------
_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs uint32 arrays (broadcastable)."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _PH_M0
            p1 = c2.astype(np.uint64) * _PH_M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def dropout_keep_mask(n_rows: int, n_cols: int, p: float, seed: int, step: int, layer: int,
                      row_offset: int = 0) -> np.ndarray:
    """Keep mask (n_rows, n_cols) the kernels draw for hidden block `layer` at optimizer `step`.

    One Philox4x32-10 call per (row, block of 8 columns): counter = (row_lo, block, layer | row_hi<<8,
    step), key = (seed_lo, seed_hi); column e of the block takes 16 bits of word e>>1 (low half first)
    and is kept iff u16 >= floor(p * 65536).  Keyed on the GLOBAL row index so the mask does not
    depend on how rows are sharded over ranks (st_dadk_b200/csrc/common.cuh: dropout_keep8).
    """
    if p <= 0.0:
        return np.ones((n_rows, n_cols), dtype=bool)
    rows64 = np.arange(n_rows, dtype=np.uint64) + np.uint64(row_offset)
    rows_lo = (rows64 & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    rows_hi = (rows64 >> np.uint64(32)).astype(np.uint32)[:, None]
    blocks = np.arange((n_cols + 7) // 8, dtype=np.uint32)[None, :]
    c2 = (np.uint32(layer) | (rows_hi << np.uint32(8))).astype(np.uint32)
    outs = philox4x32(rows_lo, blocks, c2, np.uint32(step & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    halves = []
    for w in outs:
        halves.append(w & np.uint32(0xFFFF))
        halves.append(w >> np.uint32(16))
    u = np.stack(halves, axis=-1).reshape(n_rows, -1)[:, :n_cols]
    thresh = np.uint32(min(max(int(np.float32(p) * np.float32(65536.0)), 0), 65535))
    return u >= thresh


# ----------------------------------------------------------------------------
# Model parameters and forward / backward
# ----------------------------------------------------------------------------
@dataclass
class OracleModel:
    """Plain-array image of STInterpMLP (st_interp.py:609-692); weights are (out, in)."""
    centers: np.ndarray
    bandwidths: np.ndarray
    t_centers: np.ndarray
    t_bandwidths: np.ndarray
    weights: List[np.ndarray]            # hidden linears then head
    biases: List[np.ndarray]
    ln_gamma: List[Optional[np.ndarray]]  # per hidden layer, None when layernorm=False
    ln_beta: List[Optional[np.ndarray]]
    basis_fn: str = "wendland"
    p: int = 0
    dropout: float = 0.0
    ln_eps: float = 1e-5
    delta: Optional[List[np.ndarray]] = None  # delta-reparameterised head (st_interp.py:671-686)

    def head(self, dtype):
        """Effective head (W (Q,d), b (Q,)): plain Linear, or beta_k = cumsum(delta) (st_interp.py:859-873)."""
        if self.delta is not None:
            beta = np.cumsum(np.stack([d.astype(dtype) for d in self.delta]), axis=0)
            return beta[:, 1:], beta[:, 0]
        return self.weights[-1].astype(dtype), self.biases[-1].astype(dtype)


def features(m: OracleModel, X, coords, t, dtype=np.float64):
    """cat[X, phi, psi] (st_interp.py:836-846)."""
    phi = spatial_basis(coords, m.centers, m.bandwidths, m.basis_fn, dtype)
    psi = temporal_basis(t, m.t_centers, m.t_bandwidths, dtype)
    parts = [phi, psi]
    if m.p > 0 and X is not None and np.size(X) > 0:
        parts.insert(0, np.asarray(X, dtype=dtype))
    return np.concatenate(parts, axis=1)


def _ident(x):
    return x


def forward(m: OracleModel, X, coords, t, dtype=np.float64, train: bool = False,
            keep_masks: Optional[List[np.ndarray]] = None, return_cache: bool = False, rnd=None):
    """y_hat (N, Q) (st_interp.py:827-882).  `keep_masks[l]` is the dropout keep mask of hidden layer l.

    `rnd` (e.g. tf32_round) emulates the kernels' operand rounding: it is applied to every matrix that
    enters a tensor-core product (features, hidden activations, weights) while sums stay in `dtype`;
    with rnd=None this is the reference's exact arithmetic."""
    rnd_ = (lambda a: rnd(np.asarray(a, dtype=np.float32)).astype(dtype)) if rnd is not None else _ident
    h = rnd_(features(m, X, coords, t, dtype))
    cache = {"feat": h, "layers": [], "rnd": rnd_}
    n_hidden = len(m.weights) - (0 if m.delta is not None else 1)
    for l in range(n_hidden):
        W, b = rnd_(m.weights[l].astype(dtype)), m.biases[l].astype(dtype)
        z = h @ W.T + b
        rec = {"h_in": h, "z": z}
        if m.ln_gamma[l] is not None:
            mu = z.mean(axis=1, keepdims=True)
            var = ((z - mu) ** 2).mean(axis=1, keepdims=True)
            rstd = 1.0 / np.sqrt(var + dtype(m.ln_eps))
            xh = (z - mu) * rstd
            y = xh * m.ln_gamma[l].astype(dtype) + m.ln_beta[l].astype(dtype)
            rec.update(xh=xh, rstd=rstd)
        else:
            y = z
        a = np.maximum(y, 0.0)
        rec["relu"] = y > 0
        if train and m.dropout > 0.0:
            keep = keep_masks[l]
            a = np.where(keep, a / dtype(1.0 - m.dropout), 0.0)
            rec["keep"] = keep
        a = a.astype(dtype)
        h = rnd_(a)               # what the next block's tensor-core product consumes
        cache["layers"].append(rec)
    Wh, bh = m.head(dtype)
    y_hat = a @ Wh.T + bh        # the head is evaluated on the unrounded FP32 activation
    cache["h_last"] = a
    return (y_hat, cache) if return_cache else y_hat


def loss_and_grad(y_hat, y, loss_type: str = "mse", taus: Optional[Sequence[float]] = None,
                  nc_weight: float = 0.0, nc_power: int = 1):
    """(loss, dL/dy_hat).  mse: nn.MSELoss; pinball: train_st_interp.py:37-50; multi-quantile
    mean over quantiles of the per-quantile means (:625-631); optional prediction-level
    non-crossing penalty (:53-85, :654-658)."""
    y_hat = np.asarray(y_hat)
    dt = y_hat.dtype.type
    y = np.asarray(y, dtype=y_hat.dtype).reshape(-1, 1)
    n, q = y_hat.shape
    if loss_type == "mse":
        e = y_hat - y
        return float((e * e).mean()), (2.0 * e / (n * q)).astype(y_hat.dtype)
    tau = np.asarray(taus, dtype=y_hat.dtype).reshape(1, -1)
    err = y - y_hat
    per = np.maximum((tau - 1.0) * err, tau * err)
    loss = float(per.mean(axis=0).mean())
    # d/dy_hat of max((tau-1)e, tau e), e = y - y_hat; torch.max sends the gradient to the
    # first argument on ties only when strictly greater..., ties (e == 0) split evenly in autograd.
    g = np.where(err > 0, -tau, np.where(err < 0, 1.0 - tau, 0.5 * (1.0 - 2.0 * tau)))
    grad = (g / (n * q)).astype(y_hat.dtype)
    if nc_weight > 0.0 and q > 1:
        diffs = y_hat[:, :-1] - y_hat[:, 1:]
        viol = np.maximum(diffs, 0.0)
        if nc_power == 2:
            loss += nc_weight * float((viol ** 2).sum(axis=1).mean())
            gd = 2.0 * viol
        else:
            loss += nc_weight * float(viol.sum(axis=1).mean())
            gd = (diffs > 0).astype(y_hat.dtype)
        gd = gd * dt(nc_weight / n)
        grad[:, :-1] += gd
        grad[:, 1:] -= gd
    return loss, grad


def backward(m: OracleModel, cache, d_yhat, coords=None, want_knot_grads: bool = False):
    """Gradients of every parameter given dL/dy_hat (manual reverse of `forward`).

    Returns dict with 'weights', 'biases', 'ln_gamma', 'ln_beta' lists (head last), optional
    'delta', and when `want_knot_grads`: 'centers' (K,2) and 'log_bandwidths' (K,) following
    SURVEY.md section 9.1 (phi'(r), dr/dc = -(s-c)/(d theta'), dr/dlog(theta) = -r).
    """
    dtype = d_yhat.dtype.type
    rnd_ = cache.get("rnd", _ident)
    n_hidden = len(cache["layers"])
    Wh, bh = m.head(dtype)
    h = cache["h_last"]
    g = {"weights": [None] * len(m.weights), "biases": [None] * len(m.biases),
         "ln_gamma": [None] * n_hidden, "ln_beta": [None] * n_hidden}
    dWh = d_yhat.T @ h
    dbh = d_yhat.sum(axis=0)
    if m.delta is not None:
        dbeta = np.concatenate([dbh[:, None], dWh], axis=1)       # (Q, d+1)
        g["delta"] = list(np.cumsum(dbeta[::-1], axis=0)[::-1])   # d delta_l = sum_{k>=l} d beta_k
    else:
        g["weights"][-1], g["biases"][-1] = dWh, dbh
    dh = d_yhat @ Wh
    for l in reversed(range(n_hidden)):
        rec = cache["layers"][l]
        if "keep" in rec:
            dh = np.where(rec["keep"], dh / dtype(1.0 - m.dropout), 0.0)
        dy = np.where(rec["relu"], dh, 0.0)
        if m.ln_gamma[l] is not None:
            gam = m.ln_gamma[l].astype(dtype)
            xh, rstd = rec["xh"], rec["rstd"]
            g["ln_gamma"][l] = (dy * xh).sum(axis=0)
            g["ln_beta"][l] = dy.sum(axis=0)
            gy = dy * gam
            dz = rstd * (gy - gy.mean(axis=1, keepdims=True) - xh * (gy * xh).mean(axis=1, keepdims=True))
        else:
            dz = dy
        dz_r = rnd_(dz)
        g["weights"][l] = dz_r.T @ rec["h_in"]
        g["biases"][l] = dz.sum(axis=0)
        dh = dz_r @ rnd_(m.weights[l].astype(dtype))
    if want_knot_grads:
        coords = np.asarray(coords, dtype=dtype)
        G = dh[:, m.p:m.p + m.centers.shape[0]]                   # dL/dphi (N, K_s)
        c = m.centers.astype(dtype)
        thetap = m.bandwidths.astype(dtype) * dtype(CALIBRATION[m.basis_fn])
        dx = coords[:, None, 0] - c[None, :, 0]
        dy_ = coords[:, None, 1] - c[None, :, 1]
        d = np.sqrt(dx * dx + dy_ * dy_)
        r = d / thetap[None, :]
        dphi = _PHI[m.basis_fn][1](r)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.where(d > 0, 1.0 / (d * thetap[None, :]), 0.0)
        coef = G * dphi
        g["centers"] = np.stack([(coef * (-dx * inv)).sum(axis=0), (coef * (-dy_ * inv)).sum(axis=0)], axis=1)
        g["log_bandwidths"] = (coef * (-r)).sum(axis=0)
    return g


# ----------------------------------------------------------------------------
# Step tail: clip, AdamW, EMA
# ----------------------------------------------------------------------------
def clip_coef(grads: Sequence[np.ndarray], max_norm: float) -> Tuple[float, float]:
    """torch.nn.utils.clip_grad_norm_: total L2 norm, coef = min(1, max_norm/(norm+1e-6))
    (train_st_interp.py:696-707)."""
    total = math.sqrt(sum(float((np.asarray(g, dtype=np.float64) ** 2).sum()) for g in grads))
    return total, min(1.0, max_norm / (total + 1e-6))


def adamw_ema_step(p, g, m, v, shadow, step: int, lr: float, wd: float, ema_decay: Optional[float],
                   beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, clip: float = 1.0):
    """One torch.optim.AdamW step (decoupled decay, bias-corrected) followed by
    ModelEMA.update (ema.py:52-66).  `step` is 1-based.  Arrays are updated in place."""
    g = g * clip
    p *= (1.0 - lr * wd)
    m *= beta1
    m += (1.0 - beta1) * g
    v *= beta2
    v += (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p -= (lr / bc1) * m / denom
    if ema_decay is not None and shadow is not None:
        shadow *= ema_decay
        shadow += (1.0 - ema_decay) * p
    return p, m, v, shadow


# ----------------------------------------------------------------------------
# Sharding and grid generation (bit-exact index arithmetic)
# ----------------------------------------------------------------------------
def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition [floor(r n / R), floor((r+1) n / R))  (SURVEY.md 8e)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def grid_points(nx: int, ny: int, nt: int, begin: int, end: int):
    """Points n in [begin, end) of the dense space-time grid, n = (k*nx + i)*ny + j,
    x = i/(nx-1), y = j/(ny-1), t = k/(nt-1) in FP32 (division correctly rounded)."""
    n = np.arange(begin, end, dtype=np.int64)
    j = n % ny
    i = (n // ny) % nx
    k = n // (ny * nx)
    def axis(idx, cnt):
        if cnt <= 1:
            return np.zeros(idx.shape, dtype=np.float32)
        return (idx.astype(np.float32) / np.float32(cnt - 1)).astype(np.float32)
    return np.stack([axis(i, nx), axis(j, ny)], axis=1), axis(k, nt).reshape(-1, 1)


def tf32_round(x):
    """Round-to-nearest (ties away) to 10 mantissa bits: the cvt.rna.tf32.f32 the kernels apply
    to operands before the tensor-core product."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x1000)) & np.uint64(0xFFFFE000)
    return u.astype(np.uint32).view(np.float32)

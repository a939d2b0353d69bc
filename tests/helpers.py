"""Shared helpers for the tests: golden loading, oracle model construction (test-only)."""
import ctypes
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import stdadk_oracle as orc  # noqa: E402


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def kat():
    with open(os.path.join(GOLD, "kat.json")) as f:
        return json.load(f)


def oracle_from_state(state, basis_fn="wendland", dropout=0.0):
    """Build an OracleModel from reference state_dict arrays (keys as in SURVEY.md section 5)."""
    g = lambda k: np.asarray(state[k])
    learn = "spatial_basis.log_bandwidths" in state
    centers = g("spatial_basis.centers")
    bw = np.exp(g("spatial_basis.log_bandwidths").astype(np.float64)).astype(np.float32) if learn \
        else g("spatial_basis._bandwidths")
    prefix = "mlp_trunk." if any(k.startswith("mlp_trunk.") for k in state) else "mlp."
    idx = sorted({int(k.split(".")[1]) for k in state if k.startswith(prefix)})
    weights, biases, gam, bet = [], [], [], []
    i = 0
    while i < len(idx):
        k = idx[i]
        W = g(f"{prefix}{k}.weight")
        if W.ndim == 2:
            weights.append(W)
            biases.append(g(f"{prefix}{k}.bias"))
            nxt = idx[i + 1] if i + 1 < len(idx) else None
            if nxt is not None and g(f"{prefix}{nxt}.weight").ndim == 1:
                gam.append(g(f"{prefix}{nxt}.weight"))
                bet.append(g(f"{prefix}{nxt}.bias"))
                i += 1
            else:
                gam.append(None)
                bet.append(None)
        i += 1
    delta = None
    if prefix == "mlp_trunk.":
        q = len([k for k in state if k.startswith("delta_params.")])
        delta = [g(f"delta_params.{j}") for j in range(q)]
    else:  # last linear is the head: it has no LN slot
        gam, bet = gam[:-1], bet[:-1]
    return orc.OracleModel(centers=centers, bandwidths=bw, t_centers=g("temporal_basis.centers"),
                           t_bandwidths=g("temporal_basis.bandwidths"), weights=weights, biases=biases,
                           ln_gamma=gam, ln_beta=bet, basis_fn=basis_fn, dropout=dropout, delta=delta)


def state_of(npz, prefix="state."):
    return {k[len(prefix):]: npz[k] for k in npz.files if k.startswith(prefix)}


_cref = None


def c_ref():
    """ctypes handle to oracle/_build/libbasis_ref.so (built on demand with gcc)."""
    global _cref
    if _cref is None:
        so = os.path.join(ROOT, "oracle", "_build", "libbasis_ref.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _cref = ctypes.CDLL(so)
    return _cref


def c_spatial_basis(coords, centers, thetap, fn):
    lib = c_ref()
    coords = np.ascontiguousarray(coords, dtype=np.float32)
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    thetap = np.ascontiguousarray(thetap, dtype=np.float32)
    n, k = coords.shape[0], centers.shape[0]
    phi = np.empty((n, k), dtype=np.float32)
    mask = np.empty((n, k), dtype=np.uint8)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.ref_spatial_basis_f32(P(coords), ctypes.c_int64(n), P(centers), P(thetap), ctypes.c_int(k),
                              ctypes.c_int(orc.BASIS_CODE[fn]), P(phi))
    lib.ref_support_mask(P(coords), ctypes.c_int64(n), P(centers), P(thetap), ctypes.c_int(k),
                         ctypes.c_int(orc.BASIS_CODE[fn]), P(mask))
    return phi, mask.astype(bool)

import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from helpers import orc
from st_dadk_b200 import _lib as L, ops
from test_gpu_kernels import _run_dense, T
np.set_printoptions(linewidth=200, precision=4, suppress=True)
for rows, n_in, n_out in [(128, 32, 32), (128,64,128), (200,96,48), (77,128,16), (384,256,256)]:
    rng = np.random.default_rng(rows + n_in + n_out)
    A = orc.tf32_round(rng.standard_normal((rows, n_in)).astype(np.float32))
    W = orc.tf32_round((rng.standard_normal((n_out, n_in)) / np.sqrt(n_in)).astype(np.float32))
    bias = rng.standard_normal(n_out).astype(np.float32)
    got, _ = _run_dense(ops, L, A, W, bias)
    ref = np.maximum(A.astype(np.float64) @ W.astype(np.float64).T + bias, 0.0)
    refr = orc.tf32_round(ref.astype(np.float32))
    err = np.abs(got - refr)
    print("case", rows, n_in, n_out, "maxerr", err.max(), "nan", np.isnan(got).sum(), "frac bad", (err > 1e-4).mean())
    bad = np.argwhere(err > 1e-4)
    if len(bad):
        print(" bad rows uniq", np.unique(bad[:,0])[:20], "cols uniq", np.unique(bad[:,1])[:40])
        r, c = bad[0]
        print(" first bad", r, c, got[r, c], refr[r, c])
        print(" got row0[:8]", got[0,:8], "ref", refr[0,:8])

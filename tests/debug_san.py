import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from helpers import orc
from st_dadk_b200 import _lib as L, ops
from st_dadk_b200.executor import Executor, LossSpec
from test_gpu_kernels import _run_dense, _default_oracle_model, spec_from_oracle, T
rng = np.random.default_rng(0)
A = rng.standard_normal((256, 256)).astype(np.float32); W = rng.standard_normal((256, 256)).astype(np.float32) / 16
got, _ = _run_dense(ops, L, A, W, np.zeros(256, np.float32))
print("dense ok", np.isfinite(got).all())
m = _default_oracle_model(1, q=1); m.dropout = 0.1
n = 300
coords = rng.random((n, 2)).astype(np.float32); t = rng.random((n, 1)).astype(np.float32); y = rng.standard_normal(n).astype(np.float32)
ex = Executor(spec_from_oracle(m, dropout=0.1)); ex.loss_acc.zero_()
pts = ops.make_points(T(coords), T(t))
yh = ex.forward(pts, train=True, step=1, seed=5, y=T(y), loss=LossSpec("mse"), inv_count=1.0 / n, save=True)
g = ex.backward(); torch.cuda.synchronize()
print("chain ok", float(ex.loss_acc))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from helpers import orc
from st_dadk_b200 import _lib as L, ops
from st_dadk_b200.executor import Executor, LossSpec
from test_gpu_kernels import _default_oracle_model, spec_from_oracle, T, rel_err, rel_l2
for q, loss, taus, p in [(1, "mse", None, 0.0), (5, "pinball", [0.05, 0.25, 0.5, 0.75, 0.95], 0.1)]:
    m = _default_oracle_model(42 + q, q=q); m.dropout = p
    n = 1000
    rng = np.random.default_rng(9)
    coords = rng.random((n, 2)).astype(np.float32)
    t = (rng.integers(0, 100, (n, 1)) / 99.0).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    seed, step = 0x1234567890ABCDEF, 17
    masks = [orc.dropout_keep_mask(n, w.shape[0], p, seed, step, l, 0) for l, w in enumerate(m.weights[:-1])]
    res = {}
    for name, rnd in (("fp64", None), ("tf32emu", orc.tf32_round)):
        yref, cache = orc.forward(m, None, coords, t, train=True, keep_masks=masks, return_cache=True, rnd=rnd)
        lref, dy = orc.loss_and_grad(yref, y, loss, taus)
        res[name] = (yref, lref, orc.backward(m, cache, dy))
    ex = Executor(spec_from_oracle(m, dropout=p)); ex.loss_acc.zero_()
    pts = ops.make_points(T(coords), T(t))
    yhat = ex.forward(pts, train=True, step=step, seed=seed, y=T(y), loss=LossSpec(loss, taus or ()), inv_count=1.0/(n*q), save=True)
    grads = ex.backward(); torch.cuda.synchronize()
    for name in res:
        yref, lref, gref = res[name]
        print(f"[{loss} vs {name}] yhat maxrel {rel_err(yhat.cpu().numpy(), yref):.2e} l2 {rel_l2(yhat.cpu().numpy(), yref):.2e} loss {abs(ex.loss_acc.item()-lref)/abs(lref):.2e}")
        for l in range(3):
            print(f"   dW{l} max {rel_err(grads['weights'][l].cpu().numpy(), gref['weights'][l]):.2e} l2 {rel_l2(grads['weights'][l].cpu().numpy(), gref['weights'][l]):.2e}  db {rel_err(grads['biases'][l].cpu().numpy(), gref['biases'][l]):.2e} dgam {rel_err(grads['gammas'][l].cpu().numpy(), gref['ln_gamma'][l]):.2e} dbet {rel_err(grads['betas'][l].cpu().numpy(), gref['ln_beta'][l]):.2e}")
        print(f"   dWh {rel_err(grads['head_w'].cpu().numpy(), gref['weights'][3]):.2e} dbh {rel_err(grads['head_b'].cpu().numpy(), gref['biases'][3]):.2e}")

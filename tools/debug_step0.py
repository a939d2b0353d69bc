"""Debug aid: first optimisation step of a driver parity case vs the reference's (tests/golden/_debug_step0_*.npz)."""
import importlib, os, sys, tempfile
from pathlib import Path
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import driver_cases as dc
name = sys.argv[1] if len(sys.argv) > 1 else "config1_shipped"
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
g = np.load(os.path.join(ROOT, "tests/golden", f"driver_{name}.npz"))
d = np.load(os.path.join(ROOT, "tests/golden", f"_debug_step0_{name}.npz"))
drv = importlib.import_module("scripts.train_st_interp")
from st_dadk_b200.trainer import Trainer
class Stop(Exception): pass
orig_step, orig_create = Trainer.train_step, drv.create_model
cap = {}
def create(cfg, train_coords=None):
    m = orig_create(cfg, train_coords=train_coords)
    sb = m.spatial_basis
    if sb.learnable:
        with torch.no_grad():
            sb.centers.copy_(torch.from_numpy(g["centers0"])); sb.centers_init.copy_(torch.from_numpy(g["centers0"]))
            sb.log_bandwidths.copy_(torch.from_numpy(g["bandwidths0"]).log())
    return m
def step(self, table, perm, row_begin, n_rows, *a, **k):
    self.use_cuda_graph = False
    # run compute only, then capture grads, then the update
    self._push_hyper()
    self._step_compute(table, perm, row_begin, n_rows, n_rows, 0)
    torch.cuda.synchronize()
    cap["g"] = {n: self.flat.gviews[id(p)].detach().cpu().numpy().copy() for n, p in self.model.named_parameters()}
    self._step_update()
    torch.cuda.synchronize()
    cap["p"] = {n: p.detach().cpu().numpy().copy() for n, p in self.model.named_parameters()}
    cap["sq"] = self.sqnorms.cpu().numpy()
    cap["hyper"] = self.hyper.cpu().numpy()
    raise Stop()
Trainer.train_step, drv.create_model = step, create
with tempfile.TemporaryDirectory() as tmp:
    csv = dc.case_csv(name, tmp)
    config = dict(dc.CASES[name]["config"], data_file=csv, precision=prec)
    try:
        drv._run_single_quantile_experiment(config, int(g["experiment_id"][0]), Path(tmp) / "e", "cuda", verbose=False)
    except Stop:
        pass
print("sqnorms", cap["sq"], "hyper", cap["hyper"])
tot = {}
for n in cap["g"]:
    ours, ref = cap["g"][n], d["grad0." + n]
    # reference grads are post-clip; compare up to a scalar
    sc = float((ours * ref).sum() / max((ours * ours).sum(), 1e-30))
    err = np.abs(ours * sc - ref).max() / max(np.abs(ref).max(), 1e-30)
    p1, r1 = cap["p"][n], d["state1." + n]
    perr = np.abs(p1 - r1).max()
    print(f"{n:32s} |g|ref {np.abs(ref).max():.3e} scale {sc:.5f} grad relerr {err:.2e}   param maxabs diff {perr:.3e}  (n bad>1e-3: {(np.abs(p1-r1)>1e-3).sum()}/{p1.size})")

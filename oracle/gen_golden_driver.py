"""Driver-level golden fixtures: run the UNMODIFIED reference driver (scripts/train_st_interp.py of /root/reference:
_run_single_quantile_experiment -> train_model -> evaluate_model, :2164-2505, :463-961) on the cases of
tests/driver_cases.py and commit what it did -- per-step losses, the learning rate of every parameter group at every
optimizer step, the first sample of every training batch (the RandomSampler order), per-epoch history, the initial
(GMM) and final knots, final metrics and the final model's predictions on a fixed point set.

Build container only:  python oracle/gen_golden_driver.py     (test infrastructure; never imported by the product)

The reference is not modified: observation happens through wrappers installed from outside (torch.optim.AdamW.step,
torch.Tensor.backward, a forward pre-hook on the model the reference creates).  matplotlib / seaborn are absent from
this image; they are replaced by inert mocks, and the plotting calls that follow the numeric work (after results.json
and model_final.pt are written, :2502-2505) are allowed to fail.
"""
import json
import os
import sys
import tempfile
from pathlib import Path
from unittest import mock

import numpy as np
import torch

REF = os.environ.get("STDADK_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import driver_cases as dc  # noqa: E402

for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.patches", "matplotlib.colors",
             "matplotlib.cm", "seaborn", "mpl_toolkits", "mpl_toolkits.axes_grid1"):
    sys.modules.setdefault(name, mock.MagicMock())
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "scripts"))
import train_st_interp as ref  # noqa: E402  (the reference's own script)
from stnf.models.st_interp import create_model as ref_create_model  # noqa: E402

torch.set_num_threads(4)
CASES_WITH_FP64 = {"config1_shipped": True}     # pinball loss only: nn.MSELoss refuses mixed FP64 / FP32 arguments
EVAL_POINTS = 512


def eval_points(T):
    rng = np.random.default_rng(99)
    c = rng.random((EVAL_POINTS, 2)).astype(np.float32)
    t = (rng.integers(0, T, size=(EVAL_POINTS, 1)) / max(T - 1, 1)).astype(np.float32)
    return c, t


def run_case(name, experiment_id=1, fp64=False):
    """Returns False (and writes nothing) when the run starts from a DEGENERATE first step: a gradient entry that is
    mathematically zero (pinball gradients take a few discrete values, so the head-bias gradient of a quantile cancels
    exactly when the signs of the residuals balance) is pure rounding residue, and Adam's first update
    lr*g/(|g|+1e-8) turns that residue -- whose value depends on the summation order of the machine, not on the
    algorithm -- into a parameter change of up to ~lr.  Such a run does not pin an implementation; the next experiment
    id (= next seed, as upstream's experiment loop would use) is taken instead."""
    case = dc.CASES[name]
    log = {"step_loss": [], "step_lr": [], "batch_first": [], "batch_n": []}
    created = {}

    orig_step = torch.optim.AdamW.step
    orig_backward = torch.Tensor.backward

    def step_wrapper(self, *a, **k):
        log["step_lr"].append([float(g["lr"]) for g in self.param_groups])
        first = len(log["step_lr"]) == 1
        if first:
            gs = [p.grad.detach().abs() for grp in self.param_groups for p in grp["params"] if p.grad is not None and p.numel() <= 8]
            log["min_small_grad"] = min(float(g.min()) for g in gs) if gs else 1.0
        if first and os.environ.get("STDADK_DEBUG_STEP0"):
            model = created["model"]
            dbg = {"grad0." + n: p.grad.detach().clone().numpy() for n, p in model.named_parameters() if p.grad is not None}
        r = orig_step(self, *a, **k)
        if os.environ.get("STDADK_DEBUG_STEP0"):
            mdl = created["model"]
            log.setdefault("dbg_centers", []).append(mdl.spatial_basis.centers.detach().clone().numpy())
            log.setdefault("dbg_bw", []).append(mdl.spatial_basis.bandwidths.detach().clone().numpy())
            log.setdefault("dbg_gc", []).append(mdl.spatial_basis.centers.grad.detach().clone().numpy() if getattr(mdl.spatial_basis.centers, "grad", None) is not None else np.zeros((1, 2)))
        if first and os.environ.get("STDADK_DEBUG_STEP0"):
            dbg.update({"state1." + n: p.detach().clone().numpy() for n, p in model.named_parameters()})
            np.savez_compressed(os.path.join(OUT, f"_debug_step0_{name}.npz"), **dbg)
        return r

    def backward_wrapper(self, *a, **k):
        log["step_loss"].append(float(self.detach()))
        return orig_backward(self, *a, **k)

    def create_wrapper(config, train_coords=None):
        model = ref_create_model(config, train_coords=train_coords)
        created["centers0"] = model.spatial_basis.centers.detach().clone().numpy()
        created["bandwidths0"] = model.spatial_basis.bandwidths.detach().clone().numpy()
        created["state0"] = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}

        if fp64:        # the same driver on the same module evaluated in FP64: the reference's own rounding-noise floor
            model = model.double()

        def pre_hook(mod, args):
            if mod.training:
                X, coords, t = args
                log["batch_first"].append([float(coords[0, 0]), float(coords[0, 1]), float(t[0, 0])])
                log["batch_n"].append(int(coords.shape[0]))
            if fp64:
                return tuple(a.double() for a in args)
        model.register_forward_pre_hook(pre_hook)
        created["model"] = model
        return model

    with tempfile.TemporaryDirectory() as tmp:
        csv = dc.case_csv(name, tmp)
        config = dict(case["config"], data_file=csv)
        out_dir = Path(tmp) / "experiment_001"
        out_dir.mkdir()
        with mock.patch.object(torch.optim.AdamW, "step", step_wrapper), \
                mock.patch.object(torch.Tensor, "backward", backward_wrapper), \
                mock.patch.object(ref, "create_model", create_wrapper):
            try:
                ref._run_single_quantile_experiment(config, experiment_id, out_dir, "cpu", verbose=False)
            except Exception as e:   # plotting on mocked matplotlib; the numeric artefacts are on disk by then
                print(f"[{name}] reference raised after the numeric work: {type(e).__name__}: {e}")
        if log.get("min_small_grad", 1.0) < 1e-6:
            print(f"[{name}] experiment {experiment_id}: degenerate first step (a head gradient entry is "
                  f"{log['min_small_grad']:.1e}); trying the next experiment id")
            return False
        results = json.load(open(out_dir / "results.json"))
        hist = {k: np.asarray(v, dtype=np.float64) for k, v in results["training_history"].items()}
        if fp64:
            fs = torch.load(out_dir / "model_final.pt")
            model = created["model"]
            model.load_state_dict(fs)
            model.eval()
            c, t = eval_points(case["data"]["T"])
            with torch.no_grad():
                y64 = model(torch.zeros(EVAL_POINTS, 0), torch.from_numpy(c), torch.from_numpy(t)).numpy()
            fm = {f"{split}_{k}": float(v) for split, d in results["metrics"].items() for k, v in d.items()
                  if isinstance(v, (int, float))}
            return {"step_loss64": np.asarray(log["step_loss"]), **{"hist64_" + k: v for k, v in hist.items()},
                    "run64_yhat_final": y64, "run64_centers_final": fs["spatial_basis.centers"].numpy(),
                    "run64_metric_values": np.asarray([fm[k] for k in sorted(fm)])}
        final_state = torch.load(out_dir / "model_final.pt")
        best_state = torch.load(out_dir / "model_best.pt")
        model = created["model"]
        model.load_state_dict(final_state)
        model.eval()
        T = case["data"]["T"]
        c, t = eval_points(T)
        with torch.no_grad():
            yhat = model(torch.zeros(EVAL_POINTS, 0), torch.from_numpy(c), torch.from_numpy(t)).numpy()
        m64 = model.double()
        with torch.no_grad():
            yhat64 = m64(torch.zeros(EVAL_POINTS, 0, dtype=torch.float64), torch.from_numpy(c).double(),
                         torch.from_numpy(t).double()).numpy()
    metrics = results["metrics"]
    flat_metrics = {f"{split}_{k}": float(v) for split, d in metrics.items() for k, v in d.items()
                    if isinstance(v, (int, float))}
    out = {
        "step_loss": np.asarray(log["step_loss"]), "step_lr": np.asarray(log["step_lr"]),
        "batch_first": np.asarray(log["batch_first"], dtype=np.float32), "batch_n": np.asarray(log["batch_n"]),
        "centers0": created["centers0"], "bandwidths0": created["bandwidths0"],
        "centers_final": final_state["spatial_basis.centers"].numpy(),
        "eval_coords": c, "eval_t": t, "yhat_final32": yhat, "yhat_final64": yhat64,
        "metric_names": np.asarray(sorted(flat_metrics)), "metric_values": np.asarray([flat_metrics[k] for k in sorted(flat_metrics)]),
        "best_equals_final": np.asarray([all(torch.equal(best_state[k], final_state[k]) for k in final_state)]),
        "experiment_id": np.asarray([experiment_id]),
    }
    for k, v in hist.items():
        out["hist_" + k] = v
    if CASES_WITH_FP64.get(name):
        out.update(run_case(name, experiment_id, fp64=True))
        print(f"[{name}] reference FP32 vs FP64 run, per-step loss rel diff: max "
              f"{np.max(np.abs(out['step_loss64'] - out['step_loss']) / np.abs(out['step_loss64'])):.2e}")
    for k, v in created["state0"].items():     # initial weights: same seed => same init calls => must match bit for bit
        out["state0_sum." + k] = np.asarray([v.astype(np.float64).sum(), (v.astype(np.float64) ** 2).sum()])
    if os.environ.get("STDADK_DEBUG_STEP0") and "dbg_centers" in log:
        np.savez_compressed(os.path.join(OUT, f"_debug_traj_{name}.npz"), centers=np.asarray(log["dbg_centers"]),
                            bw=np.asarray(log["dbg_bw"]), gc=np.asarray(log["dbg_gc"]))
    np.savez_compressed(os.path.join(OUT, f"driver_{name}.npz"), **out)
    print(f"[{name}] steps={len(log['step_loss'])} epochs={len(hist['train_loss'])} "
          f"train_loss={hist['train_loss']} val_loss={hist['val_loss']} lr={hist['lr']}")
    print(f"[{name}] metrics: {flat_metrics}")
    return True


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(dc.CASES)):
        eid = 1
        while not run_case(n, eid):
            eid += 1

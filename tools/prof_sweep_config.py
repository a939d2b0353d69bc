"""Where the HOST time of one sweep configuration goes (BASELINE config 5 shape: 2a-like file S=1000 x T=100, 50 epochs):
cProfile of scripts/train_st_interp.py: run_single_experiment, second run of the process (the first pays CUDA context,
library load and graph captures)."""
import cProfile, io, os, pstats, sys, time, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, yaml
import bench_extra as be
from scripts.train_st_interp import run_single_experiment
out = "/tmp/stdadk_prof_sweep"
os.makedirs(out, exist_ok=True)
csv = os.path.join(out, "2a_like.csv")
be._write_kaust_csv(csv, np.round(np.random.default_rng(11).random((1000, 2)), 6), 100)
base = yaml.safe_load(open(os.path.join(ROOT, "configs", "config_st_interp.yaml")))
base.update(data_file=csv, epochs=50, patience=50, n_experiments=1, obs_method="random", regression_type="mean")
for mode in ("uniform-fixed", "gmm-learnable"):
    init, learn = mode.split("-")
    cfg = dict(base, spatial_init_method=init, spatial_learnable=learn == "learnable")
    with contextlib.redirect_stdout(io.StringIO()):
        run_single_experiment(cfg, 1, os.path.join(out, mode + "_warm"), "cuda:0", verbose=False)
    pr = cProfile.Profile()
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        pr.enable()
        run_single_experiment(dict(cfg, lr=0.011), 1, os.path.join(out, mode), "cuda:0", verbose=False)
        torch.cuda.synchronize()
        pr.disable()
    print(f"== {mode}: {time.time() - t0:.3f} s")
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print("\n".join(ln[:150] for ln in s.getvalue().splitlines()[4:]))

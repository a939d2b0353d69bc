"""fit() on a seeded synthetic problem: `python tools/check_fit_ddp.py out.json` (one GPU) or under torchrun (data
parallel: every global batch split over the ranks, one gradient exchange per step).  Rank 0 writes the per-epoch
history and the final parameters' checksum; tests/test_gpu_ddp.py compares the two runs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
world, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
from stnf.models import STInterpMLP
from stnf.dataio import ObservationTable
from stnf.utils import set_seed
from st_dadk_b200.trainer import fit
rng = np.random.default_rng(0)
n, nv = 6000, 1500
c, t = rng.random((n + nv, 2)).astype(np.float32), (rng.integers(0, 20, n + nv) / 19.0).astype(np.float32)
y = (np.sin(2 * np.pi * (c[:, 0] + t)) * np.cos(2 * np.pi * c[:, 1]) + 0.1 * rng.standard_normal(n + nv)).astype(np.float32)
T = lambda a, lo, hi: torch.from_numpy(a[lo:hi])
train = ObservationTable(T(c, 0, n), T(t, 0, n), T(y, 0, n))
val = ObservationTable(T(c, n, n + nv), T(t, n, n + nv), T(y, n, n + nv))
cfg = dict(lr=1e-2, weight_decay=5e-4, grad_clip=5.0, regression_type="multi-quantile", quantile_levels=[0.1, 0.5, 0.9],
           epochs=4, warmup_epochs=1, scheduler="cosine", patience=10, precision=os.environ.get("PRECISION", "tf32"),
           spatial_learnable=os.environ.get("LEARNABLE", "0") == "1", basis_lr_ratio=0.05)
set_seed(7)
model = STInterpMLP(hidden_dims=[128, 64], dropout=0.1, output_dim=3, spatial_learnable=cfg["spatial_learnable"])
model, hist, _ = fit(model, train, val, cfg, dev, output_dir=None, batch_size=500, verbose=False)
tr = model._trainer
p = tr.flat.p.double()
if world > 1:
    parts = [torch.empty_like(tr.flat.p) for _ in range(world)]
    dist.all_gather(parts, tr.flat.p)
    same = all(torch.equal(parts[0], q) for q in parts)
else:
    same = True
if int(os.environ.get("RANK", "0")) == 0:
    json.dump({"history": hist, "p_sum": float(p.sum()), "p_sq": float((p * p).sum()), "replicas_identical": same,
               "world": world, "peer_exchange": tr._peer is not None}, open(sys.argv[1], "w"))
    print("FIT OK", world, hist["train_loss"], flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

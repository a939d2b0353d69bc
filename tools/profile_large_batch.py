"""Per-kernel times of a training step of the default model at large batch sizes (throughput regime)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stnf.models import STInterpMLP
from stnf.dataio import ObservationTable
from st_dadk_b200.trainer import Trainer
dev = torch.device("cuda")
n = 2_000_000
rng = np.random.default_rng(0)
coords = rng.random((n, 2)).astype(np.float32); t = rng.random(n).astype(np.float32)
y = np.sin(6 * coords[:, 0] + t).astype(np.float32)
table = ObservationTable(torch.from_numpy(coords), torch.from_numpy(t), torch.from_numpy(y)).to(dev)
perm = torch.randperm(n).to(dev)
for B in (int(a) for a in (sys.argv[1:] or ["65536"])):
    torch.manual_seed(0)
    tr = Trainer(STInterpMLP(dropout=0.1).to(dev), dict(lr=1e-3, weight_decay=5e-4, grad_clip=10.0, regression_type="mean"),
                 dev, batches_per_epoch=100, use_cuda_graph=True)
    for i in range(4):
        tr.train_step(table, perm, i * B, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        tr.train_step(table, perm, ((i + 4) * B) % (n - B), B)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    kt = tr.profile_step(table, perm, B, B, repeats=3)
    print(json.dumps({"batch": B, "ms_per_step": round(ms, 4), "M_samples_per_s": round(B / ms / 1e3, 2),
                      "kernel_ms": {k: round(v["ms"], 4) for k, v in kt["kernels"].items()}}), flush=True)

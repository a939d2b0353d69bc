import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stnf.models import STInterpMLP
from stnf.dataio import ObservationTable
from st_dadk_b200.trainer import Trainer
B = int(os.environ.get("DBG_B", "4096")); drop = float(os.environ.get("DBG_DROP", "0.1"))
torch.manual_seed(0)
model = STInterpMLP(dropout=drop)
rng = np.random.default_rng(0)
n = 2 * B
tab = ObservationTable(torch.from_numpy(rng.random((n, 2)).astype(np.float32)), torch.from_numpy(rng.random(n).astype(np.float32)),
                       torch.from_numpy(rng.standard_normal(n).astype(np.float32))).to("cuda")
tr = Trainer(model, dict(lr=1e-3, grad_clip=10.0, regression_type="mean"), "cuda", batches_per_epoch=2, use_cuda_graph=False)
perm = torch.arange(n, device="cuda")
for s in range(2):
    tr.train_step(tab, perm, s * B, B)
torch.cuda.synchronize()
print("ok", tr.pop_loss_sum())

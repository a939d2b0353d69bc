"""Profiling driver: dense prediction of 1M space-time points (3 layer_fwd launches), eager launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stnf.models import STInterpMLP
from st_dadk_b200.predict import Predictor
torch.manual_seed(0)
model = STInterpMLP(dropout=0.1).to("cuda").eval()
pr = Predictor(model)
n = int(os.environ.get("PRED_N", "1000"))
for _ in range(2):
    out, _ = pr.grid(n, 1000, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out, _ = pr.grid(n, 1000, 1); e1.record(); torch.cuda.synchronize()
print("points", n * 1000, "ms", e0.elapsed_time(e1), "Mpts/s", n * 1000 / e0.elapsed_time(e1) / 1e3)

"""Profiling driver: dense prediction of 1M grid points.  Default: the whole-network kernel (stdadk_predict);
PRED_LAYERED=1: one layer_fwd launch per block; PRED_DBG=1 (library built with -DSTDADK_PF_DEBUG): wait/phase cycles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stnf.models import STInterpMLP
from st_dadk_b200.predict import Predictor
torch.manual_seed(0)
model = STInterpMLP(dropout=0.1).to("cuda").eval()
pr = Predictor(model)
pr._prepare()
if os.environ.get("PRED_LAYERED"):
    pr.ex.fused_predict = False
n = int(os.environ.get("PRED_N", "1000"))
for _ in range(2):
    out, _ = pr.grid(n, 1000, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out, _ = pr.grid(n, 1000, 1); e1.record(); torch.cuda.synchronize()
print("points", n * 1000, "ms", e0.elapsed_time(e1), "Mpts/s", n * 1000 / e0.elapsed_time(e1) / 1e3)

if os.environ.get("PRED_DBG"):
    from st_dadk_b200 import _lib as L
    import ctypes
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.lib().stdadk_debug_counters(ctypes.c_void_p(cnt.data_ptr()))
    out, _ = pr.grid(n, 1000, 1)
    torch.cuda.synchronize()
    L.lib().stdadk_debug_counters(ctypes.c_void_p(0))
    c = cnt.cpu().numpy().astype(float)
    ctas = min(148, (n * 1000 + 127) // 128)
    names = ["mma wait A", "mma wait W", "mma total", "worker wait acc", "worker wait aempty", "worker total",
             "worker LN barrier", "producer wait wempty", "phase gen (incl. aempty wait)", "phase tmem-load+stats",
             "phase normalize+store", "phase last block + head"]
    for nm, v in zip(names, c):
        print(f"{nm:22s} {v / ctas:12.0f} cycles/CTA")

"""Profiling driver: space-time field kernel on the 10M-point grid; with STDADK_LIB pointing at a -DSTDADK_PF_DEBUG build
(tools/build_dbg.sh) it prints where each role's cycles go."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stnf.models import STInterpMLP
from st_dadk_b200.predict import Predictor
from st_dadk_b200 import _lib as L
torch.manual_seed(0)
q = int(os.environ.get("PRED_Q", "1"))
model = STInterpMLP(dropout=0.1, output_dim=q).to("cuda").eval()
pr = Predictor(model, static_weights=True)
nx, ny, nt = 1000, 1000, int(os.environ.get("PRED_T", "10"))
for _ in range(2):
    out, _ = pr.grid(nx, ny, nt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out, _ = pr.grid(nx, ny, nt); e1.record(); torch.cuda.synchronize()
n = nx * ny * nt
print("field kernel" if pr.used_field_kernel else "generic kernel", "points", n, "ms", e0.elapsed_time(e1), "Gpts/s", n / e0.elapsed_time(e1) / 1e6)
if os.environ.get("STDADK_LIB"):
    cnt = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.lib().stdadk_debug_counters(ctypes.c_void_p(cnt.data_ptr()))
    out, _ = pr.grid(nx, ny, nt)
    torch.cuda.synchronize()
    L.lib().stdadk_debug_counters(ctypes.c_void_p(0))
    c = cnt.cpu().numpy().astype(float)
    steps = (nx * ny + 127) // 128 * nt / 148.0
    names = ["mma wait H", "mma wait W", "mma total", "worker wait acc", "worker wait zt", "worker total", "worker barrier",
             "-", "phase gen", "phase ld+stats", "phase normalize+store", "phase last+head"]
    for nm, v in zip(names, c):
        print(f"{nm:24s} {v / 148:12.0f} cycles/CTA   {v / 148 / steps:9.0f} per time step")

import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stnf.models import STInterpMLP
from st_dadk_b200.predict import Predictor
torch.manual_seed(0)
pr = Predictor(STInterpMLP(dropout=0.1).to("cuda").eval())
sites = torch.rand(10000, 2, device="cuda")
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, _ = pr.space_time_field(sites, 100); e1.record(); torch.cuda.synchronize()
    print("field", i, e0.elapsed_time(e1), "ms")
for i in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, _ = pr.grid(1000, 1000, 1); e1.record(); torch.cuda.synchronize()
    print("grid", i, e0.elapsed_time(e1), "ms")
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for i in range(5):
    flush.fill_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, _ = pr.space_time_field(sites, 100); e1.record(); torch.cuda.synchronize()
    print("field after flush", i, e0.elapsed_time(e1), "ms")
for i in range(3):
    flush.fill_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out, _ = pr.grid(1000, 1000, 1); e1.record(); torch.cuda.synchronize()
    print("grid after flush", i, e0.elapsed_time(e1), "ms")
import time
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, _ = pr.space_time_field(sites, 100)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("field host enqueue", (t1 - t0) * 1e3, "ms, total", (t2 - t0) * 1e3)

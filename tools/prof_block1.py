"""Profiling driver: the fused basis + Linear1 + LayerNorm/ReLU forward kernel (layer_fwd, block 1) alone on 1M explicit
random points (the throughput regime of bench.py's `roofline` entry)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stnf.models import STInterpMLP
from st_dadk_b200.predict import Predictor
torch.manual_seed(0)
model = STInterpMLP(dropout=0.1).to("cuda").eval()
pr = Predictor(model, static_weights=True)
n = int(os.environ.get("ROWS", str(1 << 20)))
g = torch.Generator().manual_seed(5)
c, t = torch.rand(n, 2, generator=g).cuda(), torch.rand(n, generator=g).cuda()
if os.environ.get("SORTED"):
    o = torch.argsort(torch.floor(c[:, 0] * 32.0) * 2.0 + c[:, 1]); c, t = c[o].contiguous(), t[o].contiguous()
r = pr.profile_block1(c, t, repeats=3)
print("block-1 forward", r["rows"], "rows", r["ms"], "ms", r["bytes"] / r["ms"] / 1e6, "GB/s", "frac of 6531.9:", r["bytes"] / r["ms"] / 1e6 / 6531.9)

import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from helpers import golden, orc
from stnf.models import STInterpMLP
from test_gpu_model import DEFAULT, _perturb_ln
g = golden("default_mse")
torch.manual_seed(0)
model = STInterpMLP(**DEFAULT, output_dim=1); _perturb_ln(model, 0)
st = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
model = model.to("cuda").eval()
coords, t, y = [torch.tensor(g[k], device="cuda") for k in ("coords", "t", "y")]
X = torch.zeros(coords.shape[0], 0, device="cuda")
runs = []
for r in range(3):
    model.zero_grad()
    loss = torch.nn.functional.mse_loss(model(X, coords, t), y); loss.backward()
    runs.append({k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters()})
for k in runs[0]:
    d = max(np.abs(runs[0][k] - runs[i][k]).max() for i in (1, 2)) / np.abs(runs[0][k]).max()
    print(f"run-to-run {k:14s} {d:.2e}")
m = orc.OracleModel(centers=st["spatial_basis.centers"], bandwidths=st["spatial_basis._bandwidths"], t_centers=st["temporal_basis.centers"],
    t_bandwidths=st["temporal_basis.bandwidths"], weights=[st[f"mlp.{i}.weight"] for i in (0, 3, 6, 9)], biases=[st[f"mlp.{i}.bias"] for i in (0, 3, 6, 9)],
    ln_gamma=[st[f"mlp.{i}.weight"] for i in (1, 4, 7)], ln_beta=[st[f"mlp.{i}.bias"] for i in (1, 4, 7)])
for name, rnd in (("fp64", None), ("tf32emu", orc.tf32_round)):
    yh, cache = orc.forward(m, None, g["coords"], g["t"], return_cache=True, rnd=rnd)
    l, dy = orc.loss_and_grad(yh, g["y"], "mse"); gr = orc.backward(m, cache, dy)
    ref = {"mlp.0.weight": gr["weights"][0], "mlp.3.weight": gr["weights"][1], "mlp.6.weight": gr["weights"][2], "mlp.9.weight": gr["weights"][3],
           "mlp.1.weight": gr["ln_gamma"][0], "mlp.4.weight": gr["ln_gamma"][1], "mlp.7.weight": gr["ln_gamma"][2],
           "mlp.1.bias": gr["ln_beta"][0], "mlp.4.bias": gr["ln_beta"][1], "mlp.7.bias": gr["ln_beta"][2],
           "mlp.0.bias": gr["biases"][0], "mlp.3.bias": gr["biases"][1], "mlp.6.bias": gr["biases"][2], "mlp.9.bias": gr["biases"][3]}
    print(name, {k: f"{np.abs(runs[0][k] - v).max() / np.abs(v).max():.1e}" for k, v in ref.items()})

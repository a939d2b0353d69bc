import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from helpers import golden
from stnf.models import STInterpMLP
from stnf.dataio import ObservationTable
from st_dadk_b200.trainer import Trainer
g = golden("train_curve")
n, bs, steps, warm = [int(v) for v in g["meta"]]
lr, wd, clip = [float(v) for v in g["hyper"]]
DEFAULT = dict(k_spatial_centers=[25, 81, 121], k_temporal_centers=[10, 15, 45], hidden_dims=[256, 256, 128], dropout=0.0, layernorm=True)
torch.manual_seed(123)
model = STInterpMLP(**DEFAULT)
table = ObservationTable(torch.from_numpy(g["coords"]), torch.from_numpy(g["t"].reshape(-1)), torch.from_numpy(g["y"].reshape(-1))).to("cuda")
bpe = n // bs
cfg = dict(lr=lr, weight_decay=wd, grad_clip=clip, warmup_epochs=warm // bpe, epochs=100, regression_type="mean")
tr = Trainer(model, cfg, "cuda", batches_per_epoch=bpe, use_cuda_graph=False)
print("decay", tr.ema_decay, "init shadow==p", torch.equal(tr.flat.shadow, tr.flat.p))
perm = torch.arange(n, device="cuda")
sh = tr.flat.p.clone()
for s in range(steps):
    tr.train_step(table, perm, (s % bpe) * bs, bs)
    sh = tr.ema_decay * sh + (1 - tr.ema_decay) * tr.flat.p
print("shadow vs host ema", float((tr.flat.shadow - sh).abs().max()), float(sh.abs().max()))
model.eval()
X = torch.zeros(256, 0, device="cuda")
with torch.no_grad():
    raw = model(X, table.coords[:256], table.t[:256, None]).cpu().numpy()
    tr.flat.apply_shadow()
    ema = model(X, table.coords[:256], table.t[:256, None]).cpu().numpy()
    tr.flat.restore()
print("raw", raw[:4, 0], g["yhat_raw"][:4, 0]); print("ema", ema[:4, 0], g["yhat_ema"][:4, 0])
for k, v in model.named_parameters():
    gv = tr.flat.views[id(v)]
    # shadow view with the same geometry
    off = gv.storage_offset()
    shv = tr.flat.shadow.as_strided(gv.shape, gv.stride(), off)
    print(k, 'raw %.6f %.6f' % (float(v.sum()), float(v.abs().sum())), 'ema %.6f %.6f' % (float(shv.sum()), float(shv.abs().sum())))
from helpers import orc
tr.flat.apply_shadow()
st = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
with torch.no_grad():
    ema2 = model(X, table.coords[:256], table.t[:256, None]).cpu().numpy()
tr.flat.restore()
m = orc.OracleModel(centers=st['spatial_basis.centers'], bandwidths=st['spatial_basis._bandwidths'], t_centers=st['temporal_basis.centers'],
    t_bandwidths=st['temporal_basis.bandwidths'], weights=[st[f'mlp.{i}.weight'] for i in (0, 3, 6, 9)], biases=[st[f'mlp.{i}.bias'] for i in (0, 3, 6, 9)],
    ln_gamma=[st[f'mlp.{i}.weight'] for i in (1, 4, 7)], ln_beta=[st[f'mlp.{i}.bias'] for i in (1, 4, 7)])
yo = orc.forward(m, None, g["coords"][:256], g["t"][:256])
print("oracle on MY ema weights", yo[:4, 0], "gpu", ema2[:4, 0], "rel l2", np.linalg.norm(yo - ema2) / np.linalg.norm(yo))

"""GPU: data-parallel training step, two ranks sharing cuda:0 over gloo (the collective is the only thing NCCL would
change): both ranks end bit-identical, and equal the single-process run on the full global batch -- including dropout,
whose masks are keyed by the global row."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
CFG = dict(lr=1e-2, weight_decay=5e-4, grad_clip=1.0, regression_type="multi-quantile", quantile_levels=[0.1, 0.5, 0.9],
           warmup_epochs=1, scheduler="cosine", epochs=10)
N, B, STEPS = 2048, 512, 4


def _data():
    rng = np.random.default_rng(3)
    c = rng.random((N, 2)).astype(np.float32)
    t = rng.random(N).astype(np.float32)
    y = (np.sin(5 * c[:, 0]) + t).astype(np.float32)
    return c, t, y


def _run(rank, world, graph, out):
    from stnf.models import STInterpMLP
    from stnf.dataio import ObservationTable
    from st_dadk_b200.trainer import Trainer, shard_rows
    torch.manual_seed(11)
    model = STInterpMLP(hidden_dims=[64, 32], dropout=0.1, output_dim=3, spatial_learnable=True)
    c, t, y = _data()
    table = ObservationTable(torch.from_numpy(c), torch.from_numpy(t), torch.from_numpy(y)).to("cuda:0")
    tr = Trainer(model, CFG, "cuda:0", batches_per_epoch=N // B, use_cuda_graph=graph)
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(5)).to("cuda:0")
    losses = []
    for s in range(STEPS):
        lo, hi = shard_rows(B, rank, world)
        tr.train_step(table, perm, s * B + lo, hi - lo, B, key_offset=lo)
        losses.append(tr.pop_loss_sum())
    torch.cuda.synchronize()
    out[(world, rank, graph)] = (tr.flat.p.cpu().numpy(), tr.flat.shadow.cpu().numpy(), losses)


def _worker(rank, world, port, graph, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _run(rank, world, graph, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("graph", [False, True])
def test_two_rank_step_equals_single_rank(graph):
    mgr = mp.Manager()
    out = mgr.dict()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, graph, out), nprocs=2, join=True)
    _run(0, 1, graph, out)
    p0, s0, l0 = out[(2, 0, graph)]
    p1, s1, l1 = out[(2, 1, graph)]
    ps, ss, ls = out[(1, 0, graph)]
    assert np.array_equal(p0, p1) and np.array_equal(s0, s1) and l0 == l1        # replicas stay bit-identical
    assert np.max(np.abs(p0 - ps)) < 2e-5 * max(1.0, np.abs(ps).max())           # == single rank on the full batch
    assert np.allclose(l0, ls, rtol=1e-5)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box (NVLink peer memory)")
def test_peer_memory_exchange_matches_nccl_two_gpus():
    """tools/check_peer_allreduce.py under torchrun on 2 GPUs: gradient exchange through symmetric memory
    (stdadk_peer_allreduce, inside the step graph) vs the NCCL all-reduce -- replicas bit-identical, same weights."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tools", "check_peer_allreduce.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
@pytest.mark.parametrize("learnable", ["0", "1"])
def test_fit_data_parallel_equals_single_gpu(tmp_path, learnable):
    """trainer.fit under torchrun on 2 GPUs (global batches split over the ranks, peer-memory gradient exchange inside
    the step graph) follows the single-GPU run: same per-epoch training / validation losses and learning rates,
    replicas bit-identical.  (Sums of the two half-batch gradients differ from the full-batch sum in rounding only;
    dropout masks are keyed by the global row.)"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tools", "check_fit_ddp.py")
    env = dict(os.environ, LEARNABLE=learnable)
    one = subprocess.run([sys.executable, script, str(tmp_path / "one.json")], capture_output=True, text=True, timeout=300,
                         env={k: v for k, v in env.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
    assert one.returncode == 0 and "FIT OK" in one.stdout, one.stdout[-1500:] + one.stderr[-1500:]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), script, str(tmp_path / "two.json")],
                         capture_output=True, text=True, timeout=300, env=env)
    assert two.returncode == 0 and "FIT OK" in two.stdout, two.stdout[-1500:] + two.stderr[-1500:]
    a, b = json.load(open(tmp_path / "one.json")), json.load(open(tmp_path / "two.json"))
    assert b["world"] == 2 and b["replicas_identical"] and b["peer_exchange"]
    assert a["history"]["lr"] == b["history"]["lr"]
    for k in ("train_loss", "val_loss", "val_rmse"):
        assert np.allclose(a["history"][k], b["history"][k], rtol=2e-3), (k, a["history"][k], b["history"][k])
    assert abs(a["p_sq"] - b["p_sq"]) < 2e-3 * a["p_sq"]

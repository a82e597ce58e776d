"""Experiment-2 learning loop, data-parallel over NCCL (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/exp2_nccl.py
Every rank solves its shard of each mini-batch on its own B200 (no collective inside the layer), the Linear(5, n)
gradients are summed with ONE NCCL all-reduce per epoch (lqp_py_b200/sharding.py).  Rank 0 then repeats the run
alone (world of one) and checks that the sharded run traced the same loss curve and ended at the same weights
(the stop test is shard-local, reference :312, hence a tight tolerance in the solver and 1e-6 here)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from lqp_py_b200 import sharding
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dt = torch.float64
torch.set_default_dtype(dt)
dz, nB, nf, epochs, mini = 200, 64, 5, 12, 32
Q, _, A, b, lb, ub = [t.to(dev) for t in create_qp_data(dz, nB, 2 * dz, seed=0, requires_grad=False, dtype=dt)[:6]]
gen = torch.Generator().manual_seed(0)
feats = torch.randn(nB, nf, generator=gen, dtype=dt).to(dev)
p_true = (feats @ torch.randn(nf, dz, generator=gen, dtype=dt).to(dev)).unsqueeze(2)
QP = SolveBoxQP(control=box_qp_control(eps_rel=1e-9, eps_abs=1e-9))
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
model, hist = sharding.train_learn_p(QP, Q, p_true, A, b, lb, ub, feats, n_epochs=epochs, n_mini_batch=mini, lr=5e-4, seed=0)
torch.cuda.synchronize()
dt_run = time.perf_counter() - t0
flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    model1, hist1 = sharding.train_learn_p(QP, Q, p_true, A, b, lb, ub, feats, n_epochs=epochs, n_mini_batch=mini, lr=5e-4, seed=0)
    flat1 = torch.cat([p.detach().reshape(-1) for p in model1.parameters()])
    eh = max(abs(a - r) / max(abs(r), 1e-300) for a, r in zip(hist, hist1))
    ew = float((flat - flat1).abs().max() / flat1.abs().max())
    out = {"world": world, "epochs": epochs, "mini_batch": mini, "dz": dz, "ranks_hold_identical_weights": bool(same),
           "loss_curve_rel_err_vs_single_process": eh, "weights_rel_err_vs_single_process": ew,
           "ms_per_epoch": dt_run / epochs * 1e3, "loss_first": hist[0], "loss_last": hist[-1]}
    print(json.dumps(out), flush=True)
    assert same and eh < 1e-6 and ew < 1e-6, out

"""Developer tool (GPU box): iterations saved by a warm start in an Experiment-2-like sequence -- the same
batch of problems solved again after p moved by a relative step `eps` (a learning step changes p_hat slightly).
One JSON line per (dz, eps).   python tools/warm_start.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp
dev = torch.device("cuda:0")
for dz in (100, 500):
    Q, p, A, b, lb, ub = [t.to(dev) for t in create_qp_data(dz, 32, 2 * dz, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
    ctl = box_qp_control(eps_rel=1e-5, eps_abs=1e-5)
    base = torch_solve_box_qp(Q, p, A, b, lb, ub, ctl)
    for eps in (0.0, 1e-3, 1e-2, 1e-1):
        p2 = p + eps * torch.randn(p.shape, generator=torch.Generator().manual_seed(1)).to(dev)
        cold = torch_solve_box_qp(Q, p2, A, b, lb, ub, ctl)
        warm = torch_solve_box_qp(Q, p2, A, b, lb, ub, ctl, z0=base["z"], u0=base["u"])
        gap = float((warm["x"] - cold["x"]).abs().max() / cold["x"].abs().max())
        print(json.dumps({"dz": dz, "batch": 32, "p_step": eps, "iter_cold": cold["iter"], "iter_warm": warm["iter"],
                          "x_gap_rel": gap, "status_warm": warm["status"]}), flush=True)

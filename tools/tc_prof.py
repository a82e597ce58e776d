"""Developer tool (GPU box): forward + backward at dz, B (fp32) a few times -- run under ncu for a per-kernel launch list.
Usage: python tools/tc_prof.py [dz] [B] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lqp_py_b200.control import box_qp_control  # noqa: E402
from lqp_py_b200.datasets import create_qp_data  # noqa: E402
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
data = [t.to(dev) for t in create_qp_data(n, B, 2 * n, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
g = torch.ones(B, n, 1, device=dev)
for _ in range(reps):
    sol = torch_solve_box_qp(*data, control)
    torch_solve_box_qp_grad(g, sol["x"], sol["u"], sol["lams"], sol["nus"], data[0], data[2], data[4], data[5], sol["rho"])
torch.cuda.synchronize()
print("iter", sol["iter"])

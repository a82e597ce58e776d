"""Developer tool (GPU box): factor / backward-factor phase times for the pivot kernel variants (LQPB_TC_PIVOT set by caller)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lqp_py_b200 import _abi
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad
dev = torch.device("cuda:0")
data = [t.to(dev) for t in create_qp_data(500, 128, 1000, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
g = torch.ones(128, 500, 1, device=dev)
_abi.profile_enable(True)
ff = bf = 0.0
for rep in range(13):
    sol = torch_solve_box_qp(*data, control)
    prf = _abi.profile_get()
    torch_solve_box_qp_grad(g, sol["x"], sol["u"], sol["lams"], sol["nus"], data[0], data[2], data[4], data[5], sol["rho"])
    torch.cuda.synchronize()
    prb = _abi.profile_get()
    if rep >= 3:
        ff += prf["factor_ms"]; bf += prb["bwd_factor_ms"]
print(f"LQPB_TC_PIVOT={os.environ.get('LQPB_TC_PIVOT')}: factor {ff / 10:.3f} ms  bwd_factor {bf / 10:.3f} ms  iter {sol['iter']}")

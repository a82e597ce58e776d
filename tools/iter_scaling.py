"""Developer tool (GPU box): device time of the iteration kernel as a function of the iteration count (max_iters cut),
per regime -- separates the per-launch cost (load, launch, barrier set-up) from the per-iteration cost.
    python tools/iter_scaling.py [dz ...]        (LQPB_ITER=stream forces the streaming kernel)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200 import _abi
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "128"))
dt = torch.float64 if os.environ.get("DTYPE") == "f64" else torch.float32
for dz in [int(a) for a in sys.argv[1:]] or [10, 100, 250]:
    data = [t.to(dev) for t in create_qp_data(dz, B, 2 * dz, seed=0, requires_grad=False, dtype=dt)[:6]]
    _abi.profile_enable(True)
    for mi, chk in ((1, None), (61, None), (61, 1)):
        ctl = box_qp_control(eps_rel=1e-5, eps_abs=1e-5, max_iters=mi)
        if chk is not None:
            ctl['check_solved'] = chk          # the key the solver reads (reference :139)
        acc = 0.0
        for rep in range(8):
            sol = torch_solve_box_qp(*data, ctl)
            if rep >= 3:
                acc += _abi.profile_get()["iterate_ms"]
        print(f"dz={dz} B={B} {os.environ.get('LQPB_ITER', 'auto')}: max_iters={mi:3d} check={chk} iter={sol['iter']:3d} iterate {acc / 5 * 1e3:8.1f} us",
              flush=True)

"""Time the unrolled mode (control['unroll'] = True) at the headline shape on one GPU:
forward (solve + recording pass) and backward (reverse sweep + tape products + scaling glue), CUDA events.
    python tools/unroll_bench.py [--dz 500] [--batch 128] [--dtype f32] [--steps 10]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP

ap = argparse.ArgumentParser()
ap.add_argument("--dz", type=int, default=500)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dt = torch.float32 if a.dtype == "f32" else torch.float64
dev = torch.device("cuda:0")
sets = [[t.to(dev) for t in create_qp_data(a.dz, a.batch, 2 * a.dz, seed=s, requires_grad=False, dtype=dt)[:6]] for s in range(3)]
g = torch.ones(a.batch, a.dz, 1, dtype=dt, device=dev)
res = {}
for unroll in (True, False):
    QP = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5, unroll=unroll))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for k in range(3 + a.steps):
        ins = [t.detach().requires_grad_(True) for t in sets[k % 3]]
        ev[0].record()
        x = QP.forward(*ins)
        ev[1].record()
        x.backward(g)
        ev[2].record()
        torch.cuda.synchronize()
        if k >= 3:
            tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
    res["unroll" if unroll else "fixed_point"] = {"forward_ms": tf / a.steps, "backward_ms": tb / a.steps,
                                                 "qp_per_s": a.batch * a.steps / ((tf + tb) * 1e-3)}
print(json.dumps({"dz": a.dz, "batch": a.batch, "dtype": a.dtype, **res}))

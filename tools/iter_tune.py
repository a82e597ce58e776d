"""Developer tool (GPU box): time the iteration kernel for several (warps, ring depth) plans.
Usage: python tools/iter_tune.py [f32|f64] [dz] [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lqp_py_b200 import _abi  # noqa: E402
from lqp_py_b200.control import box_qp_control  # noqa: E402
from lqp_py_b200.datasets import create_qp_data  # noqa: E402
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp  # noqa: E402


def main():
    dt = torch.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else torch.float64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    dev = torch.device("cuda:0")
    sets = [[t.to(dev) for t in create_qp_data(n, B, 2 * n, seed=s, requires_grad=False, dtype=dt)[:6]] for s in range(2)]
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    _abi.profile_enable(True)
    plans = [(0, 0), (16, 2), (14, 3), (12, 4), (12, 3), (12, 2), (8, 6), (8, 4), (8, 3), (6, 8), (4, 8)]
    for nw, d in plans:
        os.environ["LQPB_ITER_WARPS"] = str(nw) if nw else ""
        os.environ["LQPB_ITER_DEPTH"] = str(d) if d else ""
        try:
            ts, fs = [], []
            for rep in range(6):
                sol = torch_solve_box_qp(*sets[rep % 2], control)
                pr = _abi.profile_get()
                if rep >= 2:
                    ts.append(pr["iterate_ms"]); fs.append(pr["factor_ms"])
            it = sol["iter"]
            print(f"warps {nw:2d} depth {d}: iterate {min(ts):.3f} ms (median {sorted(ts)[len(ts)//2]:.3f}) "
                  f"= {min(ts) * 1e3 / (it + 1):.2f} us/iter, iter {it}, factor {min(fs):.3f} ms, scale {pr['scale_ms']:.3f}", flush=True)
        except Exception as ex:
            print(f"warps {nw} depth {d}: {ex}", flush=True)


if __name__ == "__main__":
    main()

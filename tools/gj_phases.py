"""Developer tool: per-phase cycle breakdown of the Gauss-Jordan / LDL kernel (block 0).
Build with LQPB_EXTRA_NVCC_FLAGS=-DLQPB_PHASE_TIMERS python -m lqp_py_b200.build, then run on the GPU."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lqp_py_b200 import _abi  # noqa: E402
from lqp_py_b200.control import box_qp_control  # noqa: E402
from lqp_py_b200.datasets import create_qp_data  # noqa: E402
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad  # noqa: E402

NAMES = ["prologue", "A pivot", "B panels", "LDL fwd", "C update", "D writeback", "epilogue", "c vector"]


def read(L, reset=True):
    buf = (ctypes.c_longlong * 16)()
    L.lqpb_debug_phase_cycles(buf, 1 if reset else 0)
    return list(buf)[:8]


def main():
    dt = torch.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else torch.float64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    L = _abi.lib()
    dev = torch.device("cuda:0")
    data = [t.to(dev) for t in create_qp_data(n, 128, 2 * n, seed=0, requires_grad=False, dtype=dt)[:6]]
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    g = torch.ones(128, n, 1, dtype=dt, device=dev)
    for rep in range(2):
        read(L)
        sol = torch_solve_box_qp(*data, control)
        fwd = read(L)
        torch_solve_box_qp_grad(g, sol["x"], sol["u"], sol["lams"], sol["nus"], data[0], data[2], data[4], data[5], sol["rho"])
        bwd = read(L)
    for label, cyc in (("forward inverse", fwd), ("backward LDL", bwd)):
        tot = sum(cyc)
        print(f"{label}: total {tot} cycles = {tot / 1.965e3:.0f} us @1.965GHz")
        for nm, c in zip(NAMES, cyc):
            print(f"   {nm:12s} {c:10d}  {100.0 * c / max(tot, 1):5.1f}%")


if __name__ == "__main__":
    main()

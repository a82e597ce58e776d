"""Developer tool (GPU box, build with LQPB_EXTRA_NVCC_FLAGS=-DLQPB_PHASE_TIMERS): timeline of the fused block-sweep kernel
(csrc/tcfused.cu) for CTA 0's first problem and the time its roles spent waiting, forward inverse and backward LDL^T.
Usage: python tools/tc_fused_phases.py [dz] [B]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200 import _abi
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
data = [t.to(dev) for t in create_qp_data(n, B, 2 * n, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
g = torch.ones(B, n, 1, device=dev)
L = _abi.lib()
buf = (C.c_ulonglong * 64)()
nb = (n + data[2].shape[1] + 127) // 128
def show(tag):
    L.lqpb_debug_fu_ns(buf, 1)
    t0 = buf[0]
    print(f"{tag}: CTA 0, problem 0 (us from the start of its sweep)")
    prev = t0
    for k in range(nb):
        row = []
        for j, nm in enumerate(("pivot", "PANEL", "TRAIL")):
            t = buf[1 + 3 * k + j]
            if t:
                row.append(f"{nm} {(t - prev) / 1e3:7.2f}")
                prev = t
        print(f"   step {k}: " + "  ".join(row))
    print(f"   total {(prev - t0) / 1e3:.2f} us; waits (sum over the CTA's warps of that role, us): staging on free stage "
          f"{buf[32] / 1e3:.1f}, MMA on full stage {buf[33] / 1e3:.1f}, MMA on drained accumulator {buf[34] / 1e3:.1f}, "
          f"epilogue on finished accumulator {buf[35] / 1e3:.1f}")
for rep in range(2):
    sol = torch_solve_box_qp(*data, control)
    torch.cuda.synchronize()
    if rep == 1:
        show("forward inverse")
    else:
        L.lqpb_debug_fu_ns(buf, 1)
    torch_solve_box_qp_grad(g, sol["x"], sol["u"], sol["lams"], sol["nus"], data[0], data[2], data[4], data[5], sol["rho"])
    torch.cuda.synchronize()
    if rep == 1:
        show("backward LDL^T")
    else:
        L.lqpb_debug_fu_ns(buf, 1)

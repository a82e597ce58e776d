"""Developer tool (GPU box, build with LQPB_EXTRA_NVCC_FLAGS=-DLQPB_PHASE_TIMERS): clock64 shares of the phases of a
tc_tile_kernel job (thread 0 of CTA 0), PANEL and TRAIL, over one forward + backward at dz=500, B=128."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200 import _abi
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad
n, B = 500, 128
dev = torch.device("cuda:0")
data = [t.to(dev) for t in create_qp_data(n, B, 2 * n, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
g = torch.ones(B, n, 1, device=dev)
L = _abi.lib()
buf = (C.c_longlong * 16)()
def run():
    sol = torch_solve_box_qp(*data, control)
    torch_solve_box_qp_grad(g, sol["x"], sol["u"], sol["lams"], sol["nus"], data[0], data[2], data[4], data[5], sol["rho"])
    torch.cuda.synchronize()
run(); run()
L.lqpb_debug_tc_cycles(buf, 1)
run()
L.lqpb_debug_tc_cycles(buf, 0)
names = ["C-tile TMA issue", "wait stage free", "split + st.shared", "fence + syncthreads", "MMA issue", "gload next + wait MMAs",
         "TMEM -> regs -> smem + sync", "smem -> global (+C) + sync"]
for mode, nm in ((0, "PANEL"), (1, "TRAIL")):
    tot = sum(buf[mode * 8 + k] for k in range(8)) or 1
    print(nm, "clock64 total", tot)
    for k in range(8):
        print(f"   {names[k]:30s} {buf[mode * 8 + k]:10d} {100 * buf[mode * 8 + k] / tot:5.1f}%")

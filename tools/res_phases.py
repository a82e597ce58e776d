"""Developer tool (GPU box, build with LQPB_EXTRA_NVCC_FLAGS=-DLQPB_PHASE_TIMERS): clock64 shares of the phases of the
resident iteration kernel (thread 0 of CTA 0).   python tools/res_phases.py dz [B]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200 import _abi
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp
dz = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
data = [t.to(dev) for t in create_qp_data(dz, B, 2 * dz, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
ctl = box_qp_control(eps_rel=1e-5, eps_abs=1e-5)
L = _abi.lib()
buf = (C.c_longlong * 16)()
for _ in range(3):
    sol = torch_solve_box_qp(*data, ctl)
L.lqpb_debug_res_cycles(buf, 1)
_abi.profile_enable(True)
sol = torch_solve_box_qp(*data, ctl)
ms = _abi.profile_get()["iterate_ms"]
L.lqpb_debug_res_cycles(buf, 0)
names = ["K pass", "sync", "nus dot", "elementwise", "sync", "nus out", "Q pass", "reduce+leader", "publish+barrier"]
tot = sum(buf[:9]) or 1
print(f"dz={dz} B={B} iter={sol['iter']} iterate {ms*1e3:.1f} us; clock64 total {tot}")
for k, nm in enumerate(names):
    print(f"  {nm:18s} {buf[k]:10d} {100*buf[k]/tot:5.1f}%  ~{ms*1e3*buf[k]/tot/(sol['iter']+1):.2f} us/iter")

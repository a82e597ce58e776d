"""Developer tool (GPU box): clock64 phase totals of tc_pivot_mma_kernel (build with LQPB_EXTRA_NVCC_FLAGS=-DLQPB_PHASE_TIMERS,
run with LQPB_TC_PIVOT=m).  python tools/piv_phases.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lqp_py_b200 import _abi
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp
dev = torch.device("cuda:0")
data = [t.to(dev) for t in create_qp_data(500, 128, 1000, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
L = _abi.lib()
buf = (C.c_longlong * 8)()
torch_solve_box_qp(*data, control)
L.lqpb_debug_piv_cycles(buf, 1)
reps = 5
for _ in range(reps):
    torch_solve_box_qp(*data, control)
L.lqpb_debug_piv_cycles(buf, 1)
names = ["loop top", "operands (+fence, sync)", "issue + side work", "sync + MMA wait", "row read-back (+sync)", "epilogue", "prologue: TMEM alloc + tile load (x16)", "step-0 setup (x16)"]
launches = reps * 4          # CTA 0 exists in every launch; slice 0 and slice 1 both have a blockIdx 0
tot = sum(buf)
for n, v in zip(names, buf):
    print(f"{n:28s} {v / (reps * 8) / 16:9.0f} cycles per step   {100 * v / tot:5.1f} %")
print("total per launch", tot / (reps * 8), "cycles")

span = (C.c_ulonglong * (64 * 3))()
L.lqpb_debug_piv_span(span)
t0 = min(span[3 * i] for i in range(64))
rows = sorted((span[3 * i] - t0, span[3 * i + 1] - t0, span[3 * i + 2]) for i in range(64))
print("per-CTA (start ns, end ns, sm) of the last pivot launch, sorted by start:")
print(" ".join(f"({a},{b},{c})" for a, b, c in rows[:6]), "...", " ".join(f"({a},{b},{c})" for a, b, c in rows[-6:]))
print("kernel span", max(r[1] for r in rows), "ns; mean CTA life", sum(r[1] - r[0] for r in rows) / 64, "ns; distinct SMs", len({r[2] for r in rows}))

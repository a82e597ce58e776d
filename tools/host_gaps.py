"""Where does the host spend the step?  (GPU box)  Forward call (ends with the one stream synchronisation of a solve),
backward call (asynchronous), and the same with the profiling hooks bench.py uses.   python tools/host_gaps.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200 import _abi
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP, _solve_device, _grad_device

dev = torch.device("cuda:0")
sets = [[t.to(dev) for t in create_qp_data(500, 128, 1000, seed=s, requires_grad=False, dtype=torch.float32)[:6]] for s in range(3)]
g = torch.ones(128, 500, 1, device=dev)
ctl = box_qp_control(eps_rel=1e-5, eps_abs=1e-5)
QP = SolveBoxQP(control=ctl)
for prof in (False, True):
    _abi.profile_enable(prof)
    tf = tb = tp = 0.0
    N = 40
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(N + 5):
        if k == 5:
            torch.cuda.synchronize(); ev0.record(); tf = tb = tp = 0.0
        ins = [t.detach().requires_grad_(True) for t in sets[k % 3]]
        t0 = time.perf_counter()
        x = QP.forward(*ins)
        t1 = time.perf_counter()
        if prof:
            _abi.profile_get()
        t2 = time.perf_counter()
        x.backward(g)
        t3 = time.perf_counter()
        tf += t1 - t0; tp += t2 - t1; tb += t3 - t2
    ev1.record(); torch.cuda.synchronize()
    print(f"profiling {prof}: step {ev0.elapsed_time(ev1) / N:.3f} ms | host: forward call {tf / N * 1e3:.3f} ms, profile_get {tp / N * 1e3:.3f} ms, "
          f"backward call {tb / N * 1e3:.3f} ms")
# raw C-level calls without autograd
tf = tb = 0.0
for k in range(45):
    if k == 5:
        torch.cuda.synchronize(); t_all = time.perf_counter(); tf = tb = 0.0
    d = sets[k % 3]
    t0 = time.perf_counter()
    sol = _solve_device(*d, ctl, host_keys=())
    t1 = time.perf_counter()
    _grad_device(g, sol["_x_dev"], sol["_u_dev"], sol["_lams_dev"], sol["_nus_dev"], d[0], d[2], d[4], d[5], sol["rho_dev"], (True,) * 6)
    t2 = time.perf_counter()
    tf += t1 - t0; tb += t2 - t1
torch.cuda.synchronize()
print(f"no autograd: step {(time.perf_counter() - t_all) / 40 * 1e3:.3f} ms | host: _solve_device {tf / 40 * 1e3:.3f} ms, _grad_device {tb / 40 * 1e3:.3f} ms")

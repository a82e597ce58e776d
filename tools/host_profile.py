"""Developer tool (GPU box): where the HOST time of a forward+backward step goes (cProfile over N steps) and how the
step splits into forward call / backward call / GPU-idle gaps.   python tools/host_profile.py [dz] [B] [steps]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
dz = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
N = int(sys.argv[3]) if len(sys.argv) > 3 else 300
dev = torch.device("cuda:0")
data = [t.to(dev) for t in create_qp_data(dz, B, 2 * dz, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
g = torch.ones(B, dz, 1, device=dev)
QP = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5))
def step():
    ins = [t.detach().requires_grad_(True) for t in data]
    x = QP.forward(*ins)
    x.backward(g)
for _ in range(20):
    step()
torch.cuda.synchronize()
tf = tb = tp = 0.0
t_all0 = time.perf_counter()
for _ in range(N):
    t0 = time.perf_counter()
    ins = [t.detach().requires_grad_(True) for t in data]
    t1 = time.perf_counter()
    x = QP.forward(*ins)
    t2 = time.perf_counter()
    x.backward(g)
    t3 = time.perf_counter()
    tp += t1 - t0; tf += t2 - t1; tb += t3 - t2
torch.cuda.synchronize()
wall = time.perf_counter() - t_all0
print(f"dz={dz} B={B}: wall {wall / N * 1e6:.0f} us/step; host: leaves {tp / N * 1e6:.0f}, forward call {tf / N * 1e6:.0f} (includes the "
      f"wait for the solve), backward call {tb / N * 1e6:.0f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)

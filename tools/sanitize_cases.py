"""One forward + backward per kernel family, small enough to run under compute-sanitizer (tools/gpu_sanitize.sh):
    python tools/sanitize_cases.py rows|packed|stream|split|split64|fused|adapt|unroll|f64|f64blk"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
case = sys.argv[1]
if case == "fused":                    # the fused persistent block-sweep kernel (the dispatch picks it for B >= 64 only)
    os.environ["LQPB_TC_FUSED"] = "1"
if case == "stream":                   # one CTA per problem (small batches otherwise take the cluster-split kernel)
    os.environ["LQPB_ITER_SPLIT"] = "0"
dev = torch.device("cuda:0")
cfg = {"rows": (64, 6, torch.float32, {}),            # Gauss-Jordan factorisation, dense-row iteration kernel (cluster)
       "packed": (200, 4, torch.float32, {}),          # tcgen05 block sweep, packed resident iteration kernel
       "stream": (320, 3, torch.float32, {}),          # tcgen05 block sweep (3 blocks), TMA-streamed iteration kernel
       "split": (320, 3, torch.float32, {}),           # the same, every problem split over a cluster of 4 CTAs (iterate_split.cu)
       "split64": (400, 2, torch.float64, {}),         # cluster-split iteration kernel in fp64
       "fused": (320, 3, torch.float32, {}),           # tcgen05 block sweep as ONE persistent warp-specialised kernel (tcfused.cu)
       "adapt": (60, 4, torch.float64, {"rho": 100.0}),  # adaptive-rho refactorisation (host relaunch)
       "unroll": (64, 3, torch.float64, {"unroll": True}),
       "f64": (100, 3, torch.float64, {}),               # Gauss-Jordan factorisation in fp64
       "f64blk": (200, 3, torch.float64, {})}[case]      # blocked sweep with DMMA tile products (f64block.cu)
n, B, dt, kw = cfg
data = [t.to(dev).requires_grad_(True) for t in create_qp_data(n, B, 2 * n, seed=0, requires_grad=False, dtype=dt)[:6]]
x = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5, **kw)).forward(*data)
x.backward(torch.randn(x.shape, generator=torch.Generator().manual_seed(0), dtype=dt).to(dev))
torch.cuda.synchronize()
print(case, "ok", float(x.abs().max()), float(data[1].grad.abs().max()))

"""Large-batch robustness on one GPU (index arithmetic beyond 2^31 elements): solve + backward at dz=500 for
B = 4096 and 8192 (B n^2 = 2.05e9 elements of Q), and compare the first and last 8 problems with a small-batch
solve of the same problems.   python tools/big_batch_check.py [B ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP

dev = torch.device("cuda:0")
n = 500
for B in [int(v) for v in sys.argv[1:]] or [4096, 8192]:
    g = torch.Generator(device=dev).manual_seed(B)
    L = torch.randn(B, 2 * n, n, generator=g, device=dev)
    Q = torch.matmul(L.transpose(1, 2), L) / (2 * n)
    del L
    p = torch.randn(B, n, 1, generator=g, device=dev)
    A = torch.ones(B, 1, n, device=dev); b = torch.ones(B, 1, 1, device=dev)
    lb = -(1 + torch.rand(B, n, 1, generator=g, device=dev)); ub = 1 + torch.rand(B, n, 1, generator=g, device=dev)
    ctl = box_qp_control(eps_rel=1e-5, eps_abs=1e-5)
    ins = [t.requires_grad_(True) for t in (Q, p, A, b, lb, ub)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x = SolveBoxQP(control=ctl).forward(*ins)
    x.backward(torch.ones_like(x))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    errs = []
    for sl in (slice(0, 8), slice(B - 8, B)):
        sub = [t.detach()[sl].clone().requires_grad_(True) for t in ins]
        xs = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5)).forward(*sub)
        xs.backward(torch.ones_like(xs))
        errs.append(float((x.detach()[sl] - xs.detach()).abs().max()))
        errs.append(float((ins[0].grad[sl] - sub[0].grad).abs().max() / sub[0].grad.abs().max()))
        errs.append(float((ins[1].grad[sl] - sub[1].grad).abs().max() / sub[1].grad.abs().max()))
    feas = float((torch.matmul(A.detach(), x.detach()) - b.detach()).abs().max())
    print(json.dumps({"B": B, "first_call_s": dt, "max|Ax-b|": feas, "finite": bool(torch.isfinite(x).all() and torch.isfinite(ins[0].grad).all()),
                      "x/dQ/dp vs small-batch solve (first 8, last 8)": errs, "mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    del Q, p, A, b, lb, ub, ins, x
    torch.cuda.empty_cache()

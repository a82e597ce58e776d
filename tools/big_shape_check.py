"""Developer tool (GPU box): shapes beyond the golden / baseline sets against the CPU oracle -- many block rows in fp64
(dz = 1000: 8 block rows of the DMMA sweep) and many equality rows inside the blocked sweeps (m = 20 at n = 400).
Usage: python tools/big_shape_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import box_qp_oracle as orc
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP

dev = torch.device("cuda:0")
def rel(a, r):
    return float((a.double().cpu() - r.double()).abs().max() / r.double().abs().max().clamp(min=1e-300))

def run(tag, data, dtype, tol):
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    g = torch.randn(data[1].shape, generator=torch.Generator().manual_seed(7), dtype=dtype)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        ref_sol, ref_g = orc.solve_and_grad(*data, control, g)
    finally:
        torch.set_default_dtype(prev)
    ins = [t.to(dev).requires_grad_(True) if t is not None else None for t in data]
    x = SolveBoxQP(control=box_qp_control(eps_abs=1e-5, eps_rel=1e-5)).forward(*ins)
    x.backward(g.to(dev))
    errs = {"x": rel(x.detach(), ref_sol["x"])}
    for name, t, r in zip(("dQ", "dp", "dA", "db", "dlb", "dub"), ins, ref_g):
        if t is not None and r is not None and t.grad is not None:
            errs[name] = rel(t.grad, r)
    worst = max(errs.values())
    print(f"{tag}: " + "  ".join(f"{k} {v:.2e}" for k, v in errs.items()) + f"   -> {'ok' if worst < tol else 'FAIL'} (bound {tol:g})", flush=True)
    return worst < tol

ok = True
ok &= run("exp1 dz=1000 B=4 f64 (8 block rows, DMMA sweep)", orc.make_exp1_data(1000, 4, seed=1, dtype=torch.float64), torch.float64, 1e-8)
ok &= run("exp1 dz=700 B=3 f64 (6 block rows)", orc.make_exp1_data(700, 3, seed=2, dtype=torch.float64), torch.float64, 1e-8)
ok &= run("hard n=400 m=20 f64", orc.make_hard_data(400, 0.15, (3, 4, 5), dtype=torch.float64), torch.float64, 1e-7)
ok &= run("hard n=400 m=20 f32 (informative: the hard generator is ill-conditioned for fp32)", orc.make_hard_data(400, 0.15, (3, 4, 5), dtype=torch.float32), torch.float32, 1e-2)
print("ALL OK" if ok else "FAILURES")

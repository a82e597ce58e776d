"""Developer tool (GPU box): how far is each factorisation path from the fp32 reference (oracle) -- the margin
against the 1e-5 bar.  Usage: python tools/parity_margin.py [n] [B] [seed]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import box_qp_oracle as orc  # noqa: E402
from lqp_py_b200.control import box_qp_control  # noqa: E402
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad  # noqa: E402


def rel(a, r):
    return float((a.double().cpu() - r.double()).abs().max() / r.double().abs().max().clamp(min=1e-300))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device("cuda:0")
    Q, p, A, b, lb, ub = orc.make_exp1_data(n, B, seed=seed, dtype=torch.float32)
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    g = torch.randn(p.shape, generator=torch.Generator().manual_seed(99))
    refs = {}
    for dt in (torch.float32, torch.float64):
        torch.set_default_dtype(dt)
        refs[dt] = orc.solve_and_grad(*(t.to(dt) for t in (Q, p, A, b, lb, ub)), control, g.to(dt))
    torch.set_default_dtype(torch.float32)
    ref, rg = refs[torch.float32]
    ref64, rg64 = refs[torch.float64]
    ins = [t.to(dev) for t in (Q, p, A, b, lb, ub)]
    names = ("x", "z", "u", "lams", "nus")
    gn = ("dQ", "dp", "dA", "db", "dlb", "dub")
    print(f"n={n} B={B} seed={seed}: oracle iter fp32 {ref['iter']} fp64 {ref64['iter']}")
    print("reference fp32 vs reference fp64: " + "  ".join(f"{k} {rel(ref[k], ref64[k]):.1e}" for k in names) + "  |  " +
          "  ".join(f"{k} {rel(a, r):.1e}" for k, a, r in zip(gn, rg, rg64)))
    for mode, env in (("gj", {"LQPB_FACTOR": "gj"}), ("tc acc1", {"LQPB_TC_ACC2": "0"}), ("tc acc2", {})):
        for k in ("LQPB_FACTOR", "LQPB_TC_ACC2"):
            os.environ.pop(k, None)
        os.environ.update(env)
        sol = torch_solve_box_qp(*ins, control)
        grads = torch_solve_box_qp_grad(g.to(dev), sol["x"], sol["u"], sol["lams"], sol["nus"], ins[0], ins[2], ins[4],
                                        ins[5], sol["rho"])
        axb = float((ins[2] @ sol["x"] - ins[3]).abs().max())
        for tag, R, RG in (("vs ref fp32", ref, rg), ("vs ref fp64", ref64, rg64)):
            print(f"{mode:8s} iter {sol['iter']} {tag}: " + "  ".join(f"{k} {rel(sol[k], R[k]):.1e}" for k in names) + "  |  " +
                  "  ".join(f"{k} {rel(a, r):.1e}" for k, a, r in zip(gn, grads, RG)) + f"  | |Ax-b| {axb:.1e}")


if __name__ == "__main__":
    main()

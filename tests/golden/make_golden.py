"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run once in the build container (needs /root/reference; never runs on the GPU box):

    python tests/golden/make_golden.py

Each ``<case>.npz`` holds the inputs, the control dict (as JSON), the seeded
upstream gradient and the outputs of the reference's own functions
``torch_solve_box_qp`` (lqp_py/solve_box_qp_admm_torch.py:108-333) and
``torch_solve_box_qp_grad`` (:349-432).  For the large cases the inputs are not
stored (they are regenerated from the seed by oracle.make_exp1_data and pinned by
a checksum) and dQ is stored as two seeded probes instead of the full tensor.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from lqp_py.control import box_qp_control                              # noqa: E402  (reference)
from lqp_py.solve_box_qp_admm_torch import (torch_solve_box_qp,        # noqa: E402  (reference)
                                            torch_solve_box_qp_grad)
from lqp_py.lu_layer import TorchLU                                    # noqa: E402  (reference)
from oracle import box_qp_oracle as orc                                # noqa: E402


def npy(t):
    return None if t is None else t.detach().cpu().numpy()


def run_case(name, data, control_kw, dtype, seed_g=1234, store_inputs=True, extra_control=None):
    torch.set_default_dtype(dtype)
    Q, p, A, b, lb, ub = data
    control = box_qp_control(**control_kw)
    if extra_control:
        control.update(extra_control)
    any_ineq = bool(torch.max(lb) > -float("inf")) or bool(torch.min(ub) < float("inf"))
    if not any_ineq:                      # SolveBoxQPLayer.forward :33-38
        control["rho"] = 0
    sol = torch_solve_box_qp(Q=Q, p=p, A=A, b=b, lb=lb, ub=ub, control=control)
    gen = torch.Generator().manual_seed(seed_g)
    dl_dz = torch.randn(p.shape, generator=gen, dtype=dtype)
    grads = torch_solve_box_qp_grad(dl_dz, x=sol["x"], u=sol["u"], lams=sol["lams"], nus=sol["nus"],
                                    Q=Q, A=A, lb=lb, ub=ub, rho=sol["rho"])
    dQ, dp, dA, db, dlb, dub, _ = grads
    rho = sol["rho"]
    out = dict(
        control=json.dumps({k: v for k, v in control.items()}),
        control_kw=json.dumps(control_kw), extra_control=json.dumps(extra_control or {}),
        dtype=str(dtype).replace("torch.", ""),
        dl_dz=npy(dl_dz), x=npy(sol["x"]), z=npy(sol["z"]), u=npy(sol["u"]), lams=npy(sol["lams"]),
        iter=np.int64(sol["iter"]),
        rho=(npy(rho) if torch.is_tensor(rho) else np.float64(rho)),
        rho_is_tensor=np.bool_(torch.is_tensor(rho)),
        dp=npy(dp), dlb=npy(dlb), dub=npy(dub),
        has_A=np.bool_(A is not None),
    )
    if A is not None:
        out.update(nus=npy(sol["nus"]), dA=npy(dA), db=npy(db))
    if store_inputs:
        out.update(Q=npy(Q), p=npy(p), lb=npy(lb), ub=npy(ub), dQ=npy(dQ))
        if A is not None:
            out.update(A=npy(A), b=npy(b))
    else:
        gen = torch.Generator().manual_seed(4321)
        w = torch.randn(Q.shape[0], Q.shape[1], 2, generator=gen, dtype=dtype)
        out.update(dQ_probe=npy(torch.matmul(dQ, w)), dQ_fro=npy(torch.linalg.matrix_norm(dQ)),
                   input_checksum=np.array([float(Q.double().sum()), float(p.double().sum()),
                                            float(lb.double().sum()), float(ub.double().sum())]))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: v for k, v in out.items() if v is not None})
    print(f"{name}: iter={sol['iter']} rho_tensor={torch.is_tensor(rho)}")
    torch.set_default_dtype(torch.float32)


def exp1(n, B, seed, dtype):
    return orc.make_exp1_data(n, B, seed=seed, dtype=dtype)


def main():
    f32, f64 = torch.float32, torch.float64
    tol = dict(eps_rel=1e-5, eps_abs=1e-5)
    # --- Experiment-1 style data (BASELINE configs, scaled down), default control = config 3
    run_case("exp1_n10_b4_f64", exp1(10, 4, 0, f64), tol, f64)
    run_case("exp1_n10_b4_f32", exp1(10, 4, 0, f32), tol, f32)
    run_case("exp1_n50_b4_f64", exp1(50, 4, 1, f64), tol, f64)
    run_case("exp1_n50_b4_f32", exp1(50, 4, 1, f32), tol, f32)
    run_case("exp1_n100_b3_f64", exp1(100, 3, 2, f64), tol, f64)
    run_case("exp1_n37_b5_f64", exp1(37, 5, 3, f64), tol, f64)          # odd n (unaligned rows)
    # --- control variants
    run_case("noscale_rho1_n50_f64", exp1(50, 4, 4, f64), dict(scale=False, rho=1.0, adaptive_rho=False, **tol), f64)
    run_case("scale_rho1_n50_f64", exp1(50, 4, 4, f64), dict(rho=1.0, **tol), f64)
    run_case("beta_given_n50_f64", exp1(50, 4, 5, f64), dict(beta=0.3, **tol), f64)
    run_case("loose_tol_n50_f64", exp1(50, 4, 6, f64), dict(), f64)      # factory default 1e-3
    run_case("maxiter_n50_f64", exp1(50, 4, 6, f64), dict(max_iters=25, **tol), f64)   # no convergence
    run_case("check_solved_key_n50_f64", exp1(50, 4, 6, f64), tol, f64, extra_control={"check_solved": 7})
    # --- adaptive rho refactorisations (rho=100 / 1e-3 force updates)
    run_case("adapt_rho100_n60_f64", exp1(60, 6, 0, f64), dict(rho=100.0, **tol), f64)
    run_case("adapt_rho1e-3_n60_f64", exp1(60, 6, 0, f64), dict(rho=1e-3, **tol), f64)
    run_case("adapt_rho100_n60_f32", exp1(60, 6, 0, f32), dict(rho=100.0, **tol), f32)
    # --- no equality constraints
    Q, p, A, b, lb, ub = exp1(40, 4, 7, f64)
    run_case("noeq_n40_f64", (Q, p, None, None, lb, ub), tol, f64)
    # --- one-sided / unbounded boxes
    Q, p, A, b, lb, ub = exp1(40, 4, 8, f64)
    run_case("only_ub_n40_f64", (Q, p, A, b, torch.full_like(lb, -float("inf")), ub), tol, f64)
    run_case("only_lb_n40_f64", (Q, p, A, b, lb, torch.full_like(ub, float("inf"))), tol, f64)
    run_case("unbounded_n40_f64", (Q, p, A, b, torch.full_like(lb, -float("inf")),
                                   torch.full_like(ub, float("inf"))), tol, f64)
    lb2 = lb.clone(); lb2[:, ::3, :] = -float("inf")
    ub2 = ub.clone(); ub2[:, 1::4, :] = float("inf")
    run_case("partial_inf_n40_f64", (Q, p, A, b, lb2, ub2), tol, f64)
    # --- zero column in Q (scaling guard :164-168)
    Q, p, A, b, lb, ub = exp1(30, 3, 9, f64)
    Q = Q.clone(); Q[:, 4, :] = 0; Q[:, :, 4] = 0
    run_case("zero_col_n30_f64", (Q, p, A, b, lb, ub), tol, f64)
    # --- several general equality rows ("hard" generator, m = round(sqrt(n)))
    torch.set_default_dtype(f64)
    run_case("hard_n36_f64", orc.make_hard_data(36, 0.5, [0, 1, 2, 3], f64), tol, f64)
    run_case("hard_n64_f64", orc.make_hard_data(64, 0.3, [4, 5], f64), tol, f64)
    # --- full-size headline configs: vectors + dQ probes only
    run_case("exp1_n500_b8_f64", exp1(500, 8, 0, f64), tol, f64, store_inputs=False)
    run_case("exp1_n500_b8_f32", exp1(500, 8, 0, f32), tol, f32, store_inputs=False)
    run_case("exp1_n250_b8_f64", exp1(250, 8, 1, f64), tol, f64, store_inputs=False)
    run_case("exp1_n1000_b2_f64", exp1(1000, 2, 0, f64), tol, f64, store_inputs=False)

    # --- lu_layer (lqp_py/lu_layer.py:5-58): forward + backward through autograd
    torch.set_default_dtype(f64)
    torch.manual_seed(11)
    Bq, Nq = 3, 24
    S = torch.randn(Bq, Nq, Nq)
    M = (S + S.transpose(1, 2)) / 2 + 0.1 * torch.eye(Nq)          # symmetric indefinite (KKT-like)
    rhs = torch.randn(Bq, Nq, 1)
    M.requires_grad_(True); rhs.requires_grad_(True)
    lu = TorchLU(A=M.detach())
    xs = lu(M, rhs)
    g = torch.randn(Bq, Nq, 1)
    xs.backward(g)
    np.savez_compressed(os.path.join(HERE, "lu_layer_n24_f64.npz"), M=npy(M), rhs=npy(rhs), x=npy(xs), g=npy(g),
                        dM=npy(M.grad), drhs=npy(rhs.grad))
    print("lu_layer_n24_f64 done")
    torch.set_default_dtype(f32)


if __name__ == "__main__":
    main()

"""Golden vectors of the unrolled mode (control['unroll'] = True), from the UNMODIFIED reference.

Run once in the build container (needs /root/reference; never runs on the GPU box):

    python tests/golden/make_golden_unroll.py

With ``unroll=True`` the reference's ``SolveBoxQP.forward`` (lqp_py/solve_box_qp_admm_torch.py:13-15) calls
``torch_solve_box_qp`` directly and lets autograd differentiate the scaling (:161-197), the rho selection
(:200-203), every ADMM iteration (:259-282, the linear solve through ``TorchLULayer``, lqp_py/lu_layer.py:18-58)
and the un-scaling (:316).  Each ``tests/golden/unroll/<case>.npz`` holds the inputs (or, for the large cases,
the generator seed pinned by a checksum), the control dict, the seeded upstream gradient, the returned ``x`` and
the six input gradients autograd produced.  dQ is NOT symmetric in this mode (dl_dA = dx x^T, lu_layer.py:53);
for the large cases it is stored as probes dQ w, dQ^T w and its Frobenius norm.
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
from lqp_py.solve_box_qp_admm_torch import SolveBoxQP                  # noqa: E402  (reference)
from oracle import box_qp_oracle as orc                                # noqa: E402

OUT = os.path.join(HERE, "unroll")


def npy(t):
    return None if t is None else t.detach().cpu().numpy()


def run_case(name, data, control_kw, dtype, seed_g=1234, store_inputs=True, gen_seed=None):
    torch.set_default_dtype(dtype)
    control = box_qp_control(unroll=True, **control_kw)
    leaves = [None if t is None else t.clone().requires_grad_(True) for t in data]
    x = SolveBoxQP(control=control).forward(*leaves)
    gen = torch.Generator().manual_seed(seed_g)
    dl_dz = torch.randn(x.shape, generator=gen, dtype=dtype)
    x.backward(dl_dz)
    Q, p, A, b, lb, ub = leaves
    out = dict(control=json.dumps(dict(control)), dtype=str(dtype).replace("torch.", ""), dl_dz=npy(dl_dz),
               x=npy(x), dp=npy(p.grad), dlb=npy(lb.grad), dub=npy(ub.grad), has_A=np.bool_(A is not None))
    if A is not None:
        out.update(dA=npy(A.grad), db=npy(b.grad))
    dQ = Q.grad
    if store_inputs:
        out.update(Q=npy(Q), p=npy(p), lb=npy(lb), ub=npy(ub), dQ=npy(dQ))
        if A is not None:
            out.update(A=npy(A), b=npy(b))
    else:
        gen = torch.Generator().manual_seed(4321)
        w = torch.randn(Q.shape[0], Q.shape[1], 2, generator=gen, dtype=dtype)
        out.update(dQ_probe=npy(torch.matmul(dQ, w)), dQT_probe=npy(torch.matmul(dQ.transpose(1, 2), w)),
                   dQ_fro=npy(torch.linalg.matrix_norm(dQ)), gen_seed=np.int64(gen_seed),
                   gen_shape=np.array([Q.shape[1], Q.shape[0]]),
                   input_checksum=np.array([float(Q.double().sum()), float(p.double().sum()),
                                            float(lb.double().sum()), float(ub.double().sum())]))
    for k, v in list(out.items()):
        if v is None:            # gradient of an input autograd never reached (e.g. lb without a finite lower bound)
            out.pop(k)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"unroll/{name}: |x| {float(x.abs().max()):.4f} keys {sorted(k for k in out if k.startswith('d'))}")
    torch.set_default_dtype(torch.float32)


def exp1(n, B, seed, dtype):
    return orc.make_exp1_data(n, B, seed=seed, dtype=dtype)


def main():
    os.makedirs(OUT, exist_ok=True)
    f32, f64 = torch.float32, torch.float64
    tol = dict(eps_rel=1e-5, eps_abs=1e-5)
    run_case("exp1_n10_b4_f64", exp1(10, 4, 0, f64), tol, f64)
    run_case("exp1_n10_b4_f32", exp1(10, 4, 0, f32), tol, f32)
    run_case("exp1_n50_b4_f64", exp1(50, 4, 1, f64), tol, f64)
    run_case("exp1_n50_b4_f32", exp1(50, 4, 1, f32), tol, f32)
    run_case("exp1_n100_b3_f64", exp1(100, 3, 2, f64), tol, f64)
    run_case("exp1_n37_b5_f64", exp1(37, 5, 3, f64), tol, f64)
    run_case("exp1_n150_b3_f64", exp1(150, 3, 4, f64), tol, f64)           # n + m > 128 (block path of the setup)
    run_case("exp1_n150_b3_f32", exp1(150, 3, 4, f32), tol, f32)
    # --- control variants
    run_case("noscale_rho1_n50_f64", exp1(50, 4, 4, f64), dict(scale=False, rho=1.0, adaptive_rho=False, **tol), f64)
    run_case("noscale_rhoauto_n50_f64", exp1(50, 4, 4, f64), dict(scale=False, **tol), f64)
    run_case("scale_rho1_n50_f64", exp1(50, 4, 4, f64), dict(rho=1.0, **tol), f64)
    run_case("beta_given_n50_f64", exp1(50, 4, 5, f64), dict(beta=0.3, **tol), f64)
    run_case("loose_tol_n50_f64", exp1(50, 4, 6, f64), dict(), f64)
    run_case("maxiter_n50_f64", exp1(50, 4, 6, f64), dict(max_iters=25, **tol), f64)
    # --- adaptive-rho refactorisation inside the unrolled graph (rho_new = rho * ratio stays differentiable)
    run_case("adapt_rho100_n60_f64", exp1(60, 6, 0, f64), dict(rho=100.0, **tol), f64)
    run_case("adapt_rho1e-3_n60_f64", exp1(60, 6, 0, f64), dict(rho=1e-3, **tol), f64)
    # --- no equality constraints
    Q, p, A, b, lb, ub = exp1(40, 4, 7, f64)
    run_case("noeq_n40_f64", (Q, p, None, None, lb, ub), tol, f64)
    # --- one-sided / unbounded / partly infinite boxes
    Q, p, A, b, lb, ub = exp1(40, 4, 8, f64)
    run_case("only_ub_n40_f64", (Q, p, A, b, torch.full_like(lb, -float("inf")), ub), tol, f64)
    run_case("only_lb_n40_f64", (Q, p, A, b, lb, torch.full_like(ub, float("inf"))), tol, f64)
    run_case("unbounded_n40_f64", (Q, p, A, b, torch.full_like(lb, -float("inf")),
                                   torch.full_like(ub, float("inf"))), tol, f64)
    lb2 = lb.clone(); lb2[:, ::3, :] = -float("inf")
    ub2 = ub.clone(); ub2[:, 1::4, :] = float("inf")
    run_case("partial_inf_n40_f64", (Q, p, A, b, lb2, ub2), tol, f64)
    # --- several general equality rows ("hard" generator, m = round(sqrt(n)))
    torch.set_default_dtype(f64)
    run_case("hard_n36_f64", orc.make_hard_data(36, 0.5, [0, 1, 2, 3], f64), tol, f64)
    # --- headline size: vectors + dQ probes only
    run_case("exp1_n500_b4_f64", exp1(500, 4, 0, f64), tol, f64, store_inputs=False, gen_seed=0)
    run_case("exp1_n500_b4_f32", exp1(500, 4, 0, f32), tol, f32, store_inputs=False, gen_seed=0)
    run_case("exp1_n250_b4_f64", exp1(250, 4, 1, f64), tol, f64, store_inputs=False, gen_seed=1)


if __name__ == "__main__":
    main()

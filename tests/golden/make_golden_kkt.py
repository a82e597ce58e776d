"""Golden vectors of the KKT backward (backward='kkt'), from the UNMODIFIED reference.

Run once in the build container (needs /root/reference; never runs on the GPU box):

    python tests/golden/make_golden_kkt.py

For every case below the inputs, the upstream gradient and the forward solution (x, lams, nus) are taken from
the main fixture ``tests/golden/<case>.npz`` (reference outputs); ``tests/golden/kkt/<case>.npz`` adds what the
reference's ``torch_solve_box_qp_grad_kkt`` (lqp_py/solve_box_qp_admm_torch.py:435-584) returns for them.  The
reference's dense KKT system contains -inf for a one-sided or partly infinite box and all its gradients are NaN
there: those cases are recorded with ``reference_nan = True`` and no vectors.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from lqp_py.solve_box_qp_admm_torch import torch_solve_box_qp_grad_kkt      # noqa: E402  (reference)
from tests._golden import Case                                              # noqa: E402

CASES = ["exp1_n10_b4_f64", "exp1_n10_b4_f32", "exp1_n50_b4_f64", "exp1_n50_b4_f32", "exp1_n100_b3_f64",
         "exp1_n37_b5_f64", "noeq_n40_f64", "hard_n36_f64", "hard_n64_f64", "unbounded_n40_f64", "zero_col_n30_f64",
         "loose_tol_n50_f64", "only_ub_n40_f64", "only_lb_n40_f64", "partial_inf_n40_f64",
         "exp1_n250_b8_f64", "exp1_n500_b8_f64", "exp1_n500_b8_f32", "exp1_n1000_b2_f64"]


def npy(t):
    return None if t is None else t.detach().cpu().numpy()


def main():
    os.makedirs(os.path.join(HERE, "kkt"), exist_ok=True)
    for name in CASES:
        case = Case(name)
        torch.set_default_dtype(case.dtype)          # the reference builds G, zeros(...) in the default dtype
        Q, p, A, b, lb, ub = case.inputs()
        grads = torch_solve_box_qp_grad_kkt(case.t("dl_dz"), x=case.t("x"), lams=case.t("lams"), nus=case.t("nus"),
                                            Q=Q, A=A, lb=lb, ub=ub)
        dQ, dp, dA, db, dlb, dub, _ = grads
        nan = bool(torch.isnan(dp).any())
        out = dict(reference_nan=np.bool_(nan), has_dlb=np.bool_(dlb is not None), has_dub=np.bool_(dub is not None))
        if not nan:
            out.update(dp=npy(dp))
            for k, v in (("dA", dA), ("db", db), ("dlb", dlb), ("dub", dub)):
                if v is not None:
                    out[k] = npy(v)
            if case.full_inputs:
                out["dQ"] = npy(dQ)
            else:
                gen = torch.Generator().manual_seed(4321)
                w = torch.randn(Q.shape[0], Q.shape[1], 2, generator=gen, dtype=case.dtype)
                out.update(dQ_probe=npy(torch.matmul(dQ, w)), dQ_fro=npy(torch.linalg.matrix_norm(dQ)))
        np.savez_compressed(os.path.join(HERE, "kkt", name + ".npz"), **out)
        print(f"kkt/{name}: reference_nan={nan} dlb={dlb is not None} dub={dub is not None}")
        torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    main()

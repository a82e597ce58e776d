"""Pin the oracle (oracle/box_qp_oracle.py) to the reference's own outputs
(tests/golden/*.npz, written by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import box_qp_oracle as orc
from tests._golden import (Case, case_names, compare, GOLDEN_DIR, kkt_case_names, compare_kkt,
                           kkt_reference_is_nan, kkt_reduced_fp64, rel_err, UnrollCase, unroll_case_names)
import os


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_reference(name):
    case = Case(name)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(case.dtype)      # the reference creates temporaries in the default dtype
    try:
        Q, p, A, b, lb, ub = case.inputs()
        sol = orc.solve(Q, p, A, b, lb, ub, case.control_dict())
        grads = orc.grad(case.t("dl_dz"), sol["x"], sol["u"], sol["lams"], sol["nus"], Q, A, lb, ub, sol["rho"])
    finally:
        torch.set_default_dtype(prev)
    # same LAPACK/BLAS calls in the same order -> agreement to round-off
    tol = {"default": 1e-12 if case.dtype == torch.float64 else 2e-5}
    compare(case, sol, grads, tol)


@pytest.mark.parametrize("name", kkt_case_names())
def test_oracle_kkt_backward_matches_reference(name):
    """oracle.grad_kkt against the reference's torch_solve_box_qp_grad_kkt (:435-584), both evaluated at the
    reference's forward solution; the reduced closed form the CUDA path solves is checked alongside."""
    case = Case(name)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(case.dtype)
    try:
        Q, p, A, b, lb, ub = case.inputs()
        grads = orc.grad_kkt(case.t("dl_dz"), case.t("x"), case.t("lams"), case.t("nus"), Q, A, lb, ub)
    finally:
        torch.set_default_dtype(prev)
    if kkt_reference_is_nan(name):           # one-sided / partly infinite box: the dense system holds -inf
        assert torch.isnan(grads[1]).any()
        red = kkt_reduced_fp64(case.t("dl_dz"), case.t("x"), case.t("lams"), case.t("nus"), Q, A, lb, ub)
        assert all(torch.isfinite(g).all() for g in red if g is not None)
        return
    compare_kkt(case, grads, {"default": 1e-12 if case.dtype == torch.float64 else 2e-5})
    red = kkt_reduced_fp64(case.t("dl_dz"), case.t("x"), case.t("lams"), case.t("nus"), Q, A, lb, ub)
    lim = 1e-9 if case.dtype == torch.float64 else 2e-5
    for g, r in zip(grads, red):
        if g is not None:
            assert rel_err(g.numpy(), r.numpy()) <= lim


@pytest.mark.parametrize("name", unroll_case_names())
def test_oracle_unrolled_matches_reference(name):
    """oracle.solve_unrolled (autograd through the restated loop) against the reference run with
    control['unroll'] = True: x and all six gradients, including the adaptive-rho cases."""
    case = UnrollCase(name)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(case.dtype)
    try:
        leaves = [None if t is None else t.clone().requires_grad_(True) for t in case.inputs()]
        x = orc.solve_unrolled(*leaves, dict(case.control))
        x.backward(torch.from_numpy(case.z["dl_dz"]))
    finally:
        torch.set_default_dtype(prev)
    grads = [None if t is None else t.grad for t in leaves]
    case.compare(x.detach(), grads, {"default": 1e-11 if case.dtype == torch.float64 else 2e-5})


def test_oracle_lu_layer():
    z = np.load(os.path.join(GOLDEN_DIR, "lu_layer_n24_f64.npz"))
    M, rhs, g = (torch.from_numpy(z[k]) for k in ("M", "rhs", "g"))
    x, LU, piv = orc.lu_forward(M, rhs)
    dM, drhs = orc.lu_backward(LU, piv, x, g)
    np.testing.assert_allclose(x.numpy(), z["x"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(dM.numpy(), z["dM"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(drhs.numpy(), z["drhs"], rtol=1e-11, atol=1e-12)


def test_oracle_settings_quirks():
    """factory keys that the solver never reads."""
    c = orc.default_control(eps_abs=1e-5, eps_rel=1e-5, check_solved=3, adaptive_rho_max_iter=50)
    st = orc.derive_settings(c, 500)
    assert st.check_every == 20 and st.adaptive_until == 1000 and st.adaptive_every == 100
    assert orc.derive_settings(c, 1000).check_every == 30 and orc.derive_settings(c, 1000).adaptive_every == 90
    assert orc.derive_settings(c, 10).check_every == 1
    c["check_solved"] = 7
    assert orc.derive_settings(c, 500).check_every == 7 and orc.derive_settings(c, 500).adaptive_every == 98

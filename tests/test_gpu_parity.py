"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the public Python API,
which calls the C ABI (include/lqpb.h) -- the CUDA library is the only compute path.

Tolerances (north_star): fp64 -- x*, duals and gradients within 1e-8 relative (max-norm) of the
reference, fp32 -- 1e-5, iteration counts equal (+-2 allowed; checks happen every check_solved
iterations so in practice equal).  Documented exceptions, all properties of the reference's own
formulas rather than of this implementation:
  * fp32 nus / db: the reference's own fp32-vs-fp64 gap is 1e-5 / 8e-5 -> 2e-4 here;
  * dlb / dub: inactive coordinates that were active earlier carry u ~ 1e-17 of either sign, which the
    reference turns into +-1e-8*dv noise in dlb or dub (kkt/(rho u) * relu(+-rho u)) -> 1e-7 in fp64;
  * after an adaptive-rho update the adapted rho itself (and u = lams / rho) amplifies round-off
    (2e-11 in fp64, percent-level in fp32) while x, z, lams, nus, gradients and iter agree.
"""
import os

import numpy as np
import pytest
import torch

from oracle import box_qp_oracle as orc
from tests._golden import (Case, case_names, compare, rel_err, GOLDEN_DIR, kkt_case_names, compare_kkt,
                           kkt_reference_is_nan, kkt_reduced_fp64, UnrollCase, unroll_case_names)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import __graft_entry__ as entry
    entry.build()
    return torch.device("cuda:0")


def _tols(case):
    f64 = case.dtype == torch.float64
    tol = {"default": 1e-8 if f64 else 1e-5, "iter": 0,
           "dlb": 1e-7 if f64 else 1e-5, "dub": 1e-7 if f64 else 1e-5}
    if not f64:
        tol.update(nus=2e-4, db=2e-4)
    skip = ()
    if case.name.startswith("adapt"):
        if f64:
            tol.update(rho=1e-8, u=1e-8)
        else:
            skip = ("rho", "u")
    return tol, skip


@pytest.mark.parametrize("name", case_names())
def test_golden_case(name, dev):
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad
    case = Case(name)
    ins = [None if t is None else t.to(dev) for t in case.inputs()]
    control = case.control_dict()
    sol = torch_solve_box_qp(*ins, control)
    assert sol["x"].is_cuda and sol["x"].shape == (ins[1].shape[0], ins[1].shape[1], 1)
    grads = torch_solve_box_qp_grad(case.t("dl_dz").to(dev), sol["x"], sol["u"], sol["lams"], sol["nus"],
                                    ins[0], ins[2], ins[4], ins[5], sol["rho"])
    assert len(grads) == 7 and grads[6] is None
    tol, skip = _tols(case)
    if skip and "rho" in skip:
        assert torch.is_tensor(sol["rho"]) == case.rho_is_tensor
    compare(case, {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in sol.items()},
            [None if g is None else g.cpu() for g in grads[:6]], tol, skip=skip)


@pytest.mark.parametrize("n,B,dtype,seed", [(500, 16, torch.float64, 3), (500, 16, torch.float32, 3),
                                            (250, 12, torch.float64, 2), (100, 40, torch.float32, 5),
                                            (10, 128, torch.float64, 0), (64, 200, torch.float64, 1),
                                            (1000, 4, torch.float32, 0)])
def test_against_oracle_fresh_seeds(n, B, dtype, seed, dev):
    """CUDA path vs the CPU oracle on inputs that are not in the fixtures (B=200 > 148 SMs exercises
    several problems per persistent CTA)."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad
    Q, p, A, b, lb, ub = orc.make_exp1_data(n, B, seed=seed, dtype=dtype)
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    g = torch.randn(p.shape, generator=torch.Generator().manual_seed(99), dtype=dtype)

    def run_oracle(dt):
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dt)
        try:
            return orc.solve_and_grad(*(t.to(dt) for t in (Q, p, A, b, lb, ub)), control, g.to(dt))
        finally:
            torch.set_default_dtype(prev)

    ref, rg = run_oracle(dtype)
    f64 = dtype == torch.float64
    # fp32: the reference's own rounding noise (its fp32 run vs its fp64 run on the same data) is the floor
    # no other fp32 implementation can beat; allow max(1e-5, 4 x that gap) per quantity.
    gap = {}
    if not f64:
        ref64, rg64 = run_oracle(torch.float64)
        if ref64["iter"] == ref["iter"]:
            for k in ("x", "z", "u", "lams", "nus"):
                gap[k] = rel_err(ref[k].numpy(), ref64[k].numpy())
            for k, a32, a64 in zip(("dQ", "dp", "dA", "db", "dlb", "dub"), rg, rg64):
                gap[k] = rel_err(a32.numpy(), a64.numpy())
    ins = [t.to(dev) for t in (Q, p, A, b, lb, ub)]
    sol = torch_solve_box_qp(*ins, control)
    grads = torch_solve_box_qp_grad(g.to(dev), sol["x"], sol["u"], sol["lams"], sol["nus"], ins[0], ins[2], ins[4],
                                    ins[5], sol["rho"])
    assert abs(sol["iter"] - ref["iter"]) <= 2
    base = 1e-8 if f64 else 1e-5
    lim = {"x": base, "z": base, "u": base, "lams": base, "nus": base}
    for k, t in lim.items():
        t = max(t, 4 * gap.get(k, 0.0))
        e = rel_err(sol[k].cpu().numpy(), ref[k].numpy())
        assert e <= t, f"{k}: {e:.2e} > {t:.1e}"
    assert rel_err(sol["x"].cpu().numpy(), ref["x"].numpy()) <= (1e-8 if f64 else 1e-5)     # north_star bar on x*
    glim = {"dQ": base, "dp": base, "dA": base, "db": base, "dlb": 1e-7 if f64 else 1e-5,
            "dub": 1e-7 if f64 else 1e-5}
    for (k, t), mine, theirs in zip(glim.items(), grads[:6], rg):
        t = max(t, 4 * gap.get(k, 0.0))
        e = rel_err(mine.cpu().numpy(), theirs.numpy())
        assert e <= t, f"{k}: {e:.2e} > {t:.1e}"


@pytest.mark.parametrize("name", kkt_case_names())
def test_kkt_backward_golden(name, dev):
    """backward='kkt' (reference :435-584) through lqpb_backward_kkt_*, evaluated at the reference's forward
    solution and compared with the reference's own KKT gradients (tests/golden/kkt).  fp64 1e-8; fp32 1e-5
    except nus-driven db (2e-4, as for the fixed-point mode).  One-sided / partly infinite boxes: the reference
    returns NaN (its dense system holds -inf); here they must be finite and equal the closed-form reduced
    adjoint evaluated in fp64."""
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp_grad_kkt
    case = Case(name)
    Q, p, A, b, lb, ub = case.inputs()
    to = lambda t: None if t is None else t.to(dev)
    grads = torch_solve_box_qp_grad_kkt(to(case.t("dl_dz")), to(case.t("x")), to(case.t("lams")), to(case.t("nus")),
                                        to(Q), to(A), to(lb), to(ub))
    assert len(grads) == 7 and grads[6] is None
    grads = [None if g is None else g.cpu() for g in grads[:6]]
    f64 = case.dtype == torch.float64
    if kkt_reference_is_nan(name):
        red = kkt_reduced_fp64(case.t("dl_dz"), case.t("x"), case.t("lams"), case.t("nus"), Q, A, lb, ub)
        has_lb, has_ub = bool(lb.max() > -float("inf")), bool(ub.min() < float("inf"))
        assert (grads[4] is not None) == has_lb and (grads[5] is not None) == has_ub
        for g, r, nm in zip(grads, red, ("dQ", "dp", "dA", "db", "dlb", "dub")):
            if g is not None:
                assert torch.isfinite(g).all(), nm
                assert rel_err(g.numpy(), r.numpy()) <= 1e-8, nm
        return
    tol = {"default": 1e-8 if f64 else 1e-5}
    if not f64:
        tol.update(db=2e-4, dA=2e-5)
    compare_kkt(case, grads, tol)


def test_kkt_backward_through_the_layer(dev):
    """control['backward'] = 'kkt' (reference :63-64): forward on the GPU, KKT gradients on the leaves; checked
    against the oracle pipeline (forward + grad_kkt on the CPU).  The KKT adjoint divides by slacks clamped at
    1e-8, so it amplifies the 1e-15 forward differences a little: 1e-6 here, 1e-8 in the golden test above."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    dtype = torch.float64
    Q, p, A, b, lb, ub = orc.make_exp1_data(48, 6, seed=11, dtype=dtype)
    control = box_qp_control(eps_abs=1e-6, eps_rel=1e-6, backward='kkt')
    g = torch.randn(p.shape, generator=torch.Generator().manual_seed(5), dtype=dtype)
    torch.set_default_dtype(dtype)
    try:
        ref = orc.solve(Q, p, A, b, lb, ub, dict(control))
        rg = orc.grad_kkt(g, ref["x"], ref["lams"], ref["nus"], Q, A, lb, ub)
    finally:
        torch.set_default_dtype(torch.float32)
    ins = [t.to(dev).requires_grad_(True) for t in (Q, p, A, b, lb, ub)]
    x = SolveBoxQP(control=control).forward(*ins)
    x.backward(g.to(dev))
    for t, r, name in zip(ins, rg, ("dQ", "dp", "dA", "db", "dlb", "dub")):
        assert t.grad is not None and t.grad.shape == t.shape
        assert rel_err(t.grad.cpu().numpy(), r.numpy()) <= 1e-6, name
    # no finite bound at all: dlb / dub are None like in the reference (:572-579)
    inf = float("inf")
    ins = [t.to(dev).requires_grad_(True) for t in (Q, p, A, b, torch.full_like(lb, -inf), torch.full_like(ub, inf))]
    x = SolveBoxQP(control=box_qp_control(backward='kkt')).forward(*ins)
    x.backward(g.to(dev))
    assert ins[4].grad is None and ins[5].grad is None and ins[0].grad is not None


def test_module_autograd_and_needs_input_grad(dev):
    """SolveBoxQP(...).forward + x.backward: gradients land on the leaves that asked for them
    (Experiment 2 only differentiates p, experiment_2.py:50,88)."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    dtype = torch.float64
    Q, p, A, b, lb, ub = orc.make_exp1_data(48, 6, seed=11, dtype=dtype)
    control = box_qp_control(eps_abs=1e-6, eps_rel=1e-6)
    g = torch.randn(p.shape, generator=torch.Generator().manual_seed(5), dtype=dtype)
    torch.set_default_dtype(dtype)
    try:
        ref, rg = orc.solve_and_grad(Q, p, A, b, lb, ub, control, g)
    finally:
        torch.set_default_dtype(torch.float32)
    ins = [t.to(dev).requires_grad_(True) for t in (Q, p, A, b, lb, ub)]
    x = SolveBoxQP(control=control).forward(*ins)
    assert x.requires_grad
    x.backward(g.to(dev))
    for t, r, name in zip(ins, rg, ("dQ", "dp", "dA", "db", "dlb", "dub")):
        assert t.grad is not None and t.grad.shape == t.shape
        assert rel_err(t.grad.cpu().numpy(), r.numpy()) <= 1e-7, name
    # only p requires grad
    ins2 = [t.to(dev) for t in (Q, p, A, b, lb, ub)]
    ins2[1].requires_grad_(True)
    x2 = SolveBoxQP(control=control)(*ins2)
    loss = (x2 * g.to(dev)).sum()
    loss.backward()
    assert rel_err(ins2[1].grad.cpu().numpy(), rg[1].numpy()) <= 1e-8
    assert all(t.grad is None for k, t in enumerate(ins2) if k != 1)


def test_cpu_tensors_drop_in(dev):
    """The reference's callers pass CPU tensors; the layer stages them to the GPU and returns CPU results."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    Q, p, A, b, lb, ub = orc.make_exp1_data(32, 5, seed=4, dtype=torch.float32)
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    ref = orc.solve(Q, p, A, b, lb, ub, control)
    Q.requires_grad_(True); p.requires_grad_(True)
    x = SolveBoxQP(control=control).forward(Q=Q, p=p, A=A, b=b, lb=lb, ub=ub)
    assert x.device.type == "cpu" and x.shape == (5, 32, 1)
    assert rel_err(x.detach().numpy(), ref["x"].numpy()) <= 1e-5
    x.backward(torch.ones(5, 32, 1))
    assert Q.grad.device.type == "cpu" and p.grad.shape == p.shape


@pytest.mark.parametrize("n,B,dtype,backward", [(40, 70, torch.float64, "fixed_point"), (200, 64, torch.float32, "fixed_point"),
                                                (40, 70, torch.float64, "kkt"), (500, 128, torch.float32, "fixed_point"),
                                                (200, 64, torch.float32, "kkt")])
def test_host_buffer_pipeline_equals_device_path(n, B, dtype, backward, dev):
    """CPU tensors go through lqpb_forward_host_* / lqpb_backward_host_* (Q uploaded and dQ returned in chunks that
    overlap the per-problem setup / adjoint chains).  Problems are independent and every kernel is deterministic
    per problem, so the results must be BIT-identical to the device-tensor path (one chunk = whole batch)."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    data = orc.make_exp1_data(n, B, seed=21, dtype=dtype)
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5, backward=backward)
    g = torch.randn(B, n, 1, generator=torch.Generator().manual_seed(9), dtype=dtype)
    host = [t.clone().pin_memory().requires_grad_(True) for t in data]
    xh = SolveBoxQP(control=control).forward(*host)
    assert xh.device.type == "cpu"
    xh.backward(g)
    devs = [t.to(dev).requires_grad_(True) for t in data]
    xd = SolveBoxQP(control=control).forward(*devs)
    xd.backward(g.to(dev))
    assert torch.equal(xh.detach(), xd.detach().cpu())
    for th, td, name in zip(host, devs, ("dQ", "dp", "dA", "db", "dlb", "dub")):
        assert th.grad is not None and th.grad.device.type == "cpu", name
        assert torch.equal(th.grad, td.grad.cpu()), name
    # pageable (not pinned) host tensors take the same path
    host2 = [t.clone().requires_grad_(k < 2) for k, t in enumerate(data)]
    x2 = SolveBoxQP(control=control).forward(*host2)
    x2.backward(g)
    assert torch.equal(x2.detach(), xh.detach()) and torch.equal(host2[0].grad, host[0].grad)
    assert host2[2].grad is None and host2[4].grad is None


def test_unbounded_batch_mutates_control_like_reference(dev):
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    Q, p, A, b, lb, ub = (t.to(dev) for t in orc.make_exp1_data(20, 3, seed=1, dtype=torch.float64))
    control = box_qp_control()
    inf = float("inf")
    x = SolveBoxQP(control)(Q, p, A, b, torch.full_like(lb, -inf), torch.full_like(ub, inf))
    assert control["rho"] == 0                                   # reference :37-38
    # equality-constrained optimum: A x = b exactly, KKT stationarity
    assert float((A @ x - b).abs().max()) < 1e-12


def test_full_size_properties(dev):
    """BASELINE config 3 at full size (dz=500, B=128, fp32): properties that need no oracle run."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp
    n, B = 500, 128
    Q, p, A, b, lb, ub = (t.to(dev) for t in orc.make_exp1_data(n, B, seed=0, dtype=torch.float32))
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    sol = torch_solve_box_qp(Q, p, A, b, lb, ub, control)
    assert sol["iter"] == 60                                     # all ten seeds stop at 60
    x, z = sol["x"], sol["z"]
    # x comes from the KKT solve: A x = b up to the round-off of the explicit fp32 operator K11 (a sum over 500
    # coordinates of ~1e-7 errors; measured 1e-5 with the FP32-pipe factorisation, 2.3e-5 with the 3xTF32 one)
    assert float((A @ x - b).abs().max()) < 5e-5
    # z~ is clamped exactly in the scaled space; un-scaling (z = D z~, lb~ = lb / D) costs one rounding
    assert float((z - lb).min()) >= -1e-6 and float((ub - z).min()) >= -1e-6
    assert float((x - z).abs().max()) < 1e-3                     # primal residual small at the stop
    lam = sol["lams"]
    assert float(lam.min()) >= 0 and float((lam[:, :n] * lam[:, n:]).abs().max()) == 0   # complementary split
    # determinism / batch-position independence: same problems in reverse order, same iteration count
    perm = torch.arange(B - 1, -1, -1, device=dev)
    sol2 = torch_solve_box_qp(Q[perm].contiguous(), p[perm].contiguous(), A[perm].contiguous(), b[perm].contiguous(),
                              lb[perm].contiguous(), ub[perm].contiguous(), control)
    assert sol2["iter"] == sol["iter"]
    assert torch.equal(sol2["x"][perm], sol["x"])
    # batch shards solved separately agree with the full batch when they stop at the same check
    half = torch_solve_box_qp(Q[:64].contiguous(), p[:64].contiguous(), A[:64].contiguous(), b[:64].contiguous(),
                              lb[:64].contiguous(), ub[:64].contiguous(), control)
    assert half["iter"] == sol["iter"] and torch.equal(half["x"], sol["x"][:64])


def test_lu_layer_golden(dev):
    from lqp_py_b200.lu_layer import TorchLU, lu_factor, lu_solve
    z = np.load(os.path.join(GOLDEN_DIR, "lu_layer_n24_f64.npz"))
    M = torch.from_numpy(z["M"]).to(dev).requires_grad_(True)
    rhs = torch.from_numpy(z["rhs"]).to(dev).requires_grad_(True)
    lu = TorchLU(A=M.detach())
    x = lu(M, rhs)
    x.backward(torch.from_numpy(z["g"]).to(dev))
    assert rel_err(x.detach().cpu().numpy(), z["x"]) <= 1e-10
    assert rel_err(M.grad.cpu().numpy(), z["dM"]) <= 1e-10
    assert rel_err(rhs.grad.cpu().numpy(), z["drhs"]) <= 1e-10
    # factors are LAPACK-compatible: torch's own lu_solve accepts them
    LU, P = lu_factor(M.detach())
    tl, tp = torch.linalg.lu_factor(M.detach().cpu())
    assert torch.equal(P.cpu(), tp)
    assert rel_err(LU.cpu().numpy(), tl.numpy()) <= 1e-10
    big = torch.randn(3, 130, 130, dtype=torch.float32, generator=torch.Generator().manual_seed(0))
    rb = torch.randn(3, 130, 2, dtype=torch.float32, generator=torch.Generator().manual_seed(1))
    LU, P = lu_factor(big.to(dev))
    xs = lu_solve(LU, P, rb.to(dev))
    ref = torch.linalg.solve(big.double(), rb.double())
    assert rel_err(xs.cpu().numpy(), ref.numpy()) <= 2e-3


def test_dtype_mismatch_raises(dev):
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp
    Q, p, A, b, lb, ub = (t.to(dev) for t in orc.make_exp1_data(8, 2, dtype=torch.float32))
    with pytest.raises(TypeError):
        torch_solve_box_qp(Q.double(), p, A, b, lb, ub, box_qp_control())
    with pytest.raises(TypeError):
        torch_solve_box_qp(Q.half(), p.half(), A.half(), b.half(), lb.half(), ub.half(), box_qp_control())


def test_verbose_prints_like_reference(dev, capsys):
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp
    Q, p, A, b, lb, ub = orc.make_exp1_data(30, 4, seed=2, dtype=torch.float64)
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5, verbose=True)
    torch.set_default_dtype(torch.float64)
    try:
        orc.solve(Q, p, A, b, lb, ub, control)
    finally:
        torch.set_default_dtype(torch.float32)
    ref_out = capsys.readouterr().out
    torch_solve_box_qp(*(t.to(dev) for t in (Q, p, A, b, lb, ub)), control)
    out = capsys.readouterr().out
    assert out.splitlines()[0::3] == ref_out.splitlines()[0::3]          # iteration lines
    ours = [float(l.split("=")[1]) for l in out.splitlines() if "error" in l]
    theirs = [float(l.split("=")[1]) for l in ref_out.splitlines() if "error" in l]
    assert len(ours) == len(theirs) and np.allclose(ours, theirs, rtol=0, atol=2e-9)


# ------------------------------------------------------------------------------------------ unrolled mode
@pytest.mark.parametrize("name", unroll_case_names())
@pytest.mark.parametrize("where", ["cuda", "cpu"])
def test_unrolled_golden_case(name, where, dev):
    """control['unroll'] = True (reference :13-15): x and the six input gradients of the recorded loop + reverse
    sweep kernels against the reference's autograd-through-the-loop run.  fp64 1e-8; fp32 1e-5 except where the
    reference's own fp32 noise is larger (its fp32-vs-fp64 gap on these cases is up to 3e-5 in dA)."""
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    case = UnrollCase(name)
    if where == "cpu" and not (name.startswith("exp1_n50") or name.startswith("adapt_rho100")):
        pytest.skip("CPU-tensor drop-in is exercised on two cases")
    target = dev if where == "cuda" else torch.device("cpu")
    leaves = [None if t is None else t.to(target).requires_grad_(True) for t in case.inputs()]
    control = dict(case.control)
    x = SolveBoxQP(control=control).forward(*leaves)
    assert torch.is_tensor(x) and x.device.type == target.type and x.requires_grad
    x.backward(torch.from_numpy(case.z["dl_dz"]).to(target))
    grads = [None if t is None else (None if t.grad is None else t.grad.cpu()) for t in leaves]
    f64 = case.dtype == torch.float64
    tol = {"default": 1e-8 if f64 else 1e-5}
    if not f64:
        tol.update(dA=1e-4, db=1e-4, dQ=3e-5, dQ_probe=3e-5, dQT_probe=3e-5)
    case.compare(x.detach().cpu(), grads, tol)


def test_unrolled_matches_oracle_fresh_seed_and_needs_input_grad(dev):
    """Fresh inputs (not in the fixtures), B > 148 so that persistent CTAs sweep several problems, only p and lb
    require a gradient (dQ~ / dA~ products are skipped)."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    dtype = torch.float64
    Q, p, A, b, lb, ub = orc.make_exp1_data(48, 160, seed=21, dtype=dtype)
    g = torch.randn(p.shape, generator=torch.Generator().manual_seed(5), dtype=dtype)
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5, unroll=True)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        leaves = [t.clone().requires_grad_(True) for t in (Q, p, A, b, lb, ub)]
        xr = orc.solve_unrolled(*leaves, dict(control))
        xr.backward(g)
    finally:
        torch.set_default_dtype(prev)
    ins = [t.to(dev) for t in (Q, p, A, b, lb, ub)]
    ins[1].requires_grad_(True)
    ins[4].requires_grad_(True)
    x = SolveBoxQP(control=control).forward(*ins)
    x.backward(g.to(dev))
    assert ins[0].grad is None and ins[2].grad is None
    assert rel_err(x.detach().cpu().numpy(), xr.detach().numpy()) <= 1e-8
    assert rel_err(ins[1].grad.cpu().numpy(), leaves[1].grad.numpy()) <= 1e-8
    assert rel_err(ins[4].grad.cpu().numpy(), leaves[4].grad.numpy()) <= 1e-8


def test_unrolled_full_size_fp32(dev):
    """dz=500, B=128 (the headline shape) in unrolled mode: x equals the implicit layer's x (same kernels; the final
    un-scaling D x~ is a torch multiply there, hence not bit for bit), the unrolled dp approaches the fixed-point dp (the ADMM map is contractive; the reference's own gap
    between the two modes is ~1e-3 relative at tol 1e-5), dQ is finite and NOT symmetric (lu_layer.py:53)."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    Q, p, A, b, lb, ub = orc.make_exp1_data(500, 128, seed=0, dtype=torch.float32)
    g = torch.randn(p.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float32).to(dev)
    outs = {}
    for unroll in (False, True):
        ins = [t.to(dev).requires_grad_(True) for t in (Q, p, A, b, lb, ub)]
        x = SolveBoxQP(control=box_qp_control(eps_abs=1e-5, eps_rel=1e-5, unroll=unroll)).forward(*ins)
        x.backward(g)
        outs[unroll] = (x.detach(), [t.grad for t in ins])
    assert float((outs[True][0] - outs[False][0]).abs().max()) <= 1e-6 * float(outs[False][0].abs().max())
    dp_u, dp_f = outs[True][1][1], outs[False][1][1]
    assert float((dp_u - dp_f).abs().max() / dp_f.abs().max()) < 2e-2
    dQ = outs[True][1][0]
    assert torch.isfinite(dQ).all()
    assert float((dQ - dQ.transpose(1, 2)).abs().max()) > 1e-3 * float(dQ.abs().max())


def test_boxqpth_stateful_wrapper(dev):
    """BoxQPTH (reference :70-105): solve() returns x and keeps the dict; update() swaps p; passing lb / ub to
    update() stores None exactly like the reference does (:99-102)."""
    from lqp_py_b200.solve_box_qp_admm_torch import BoxQPTH
    case = Case("exp1_n50_b4_f64")
    Q, p, A, b, lb, ub = [t.to(dev) for t in case.inputs()]
    holder = BoxQPTH(Q, p, A, b, lb, ub, case.control_dict())
    x = holder.solve()
    assert rel_err(x.cpu().numpy(), case.z["x"]) <= 1e-8 and holder.sol["iter"] == case.iter
    assert holder.update(p=2 * p) is None
    x2 = holder.solve()
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        ref = orc.solve(*[t.cpu() for t in (Q, 2 * p, A, b, lb, ub)], case.control_dict())
    finally:
        torch.set_default_dtype(prev)
    assert rel_err(x2.cpu().numpy(), ref["x"].numpy()) <= 1e-8
    holder.update(lb=lb, ub=ub)
    assert holder.lb is None and holder.ub is None


# ------------------------------------------------------------------------------------------ sizes at the seams
def _oracle_run(data, control, g, dtype):
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        return orc.solve_and_grad(*data, control, g)
    finally:
        torch.set_default_dtype(prev)


@pytest.mark.parametrize("n,B,dtype", [(1, 1, torch.float64), (2, 3, torch.float64), (5, 1, torch.float32),
                                       (127, 3, torch.float32), (128, 3, torch.float32), (129, 2, torch.float64),
                                       (255, 2, torch.float32), (257, 2, torch.float32), (384, 2, torch.float32),
                                       (640, 2, torch.float32), (33, 150, torch.float32)])
def test_sizes_at_the_seams(n, B, dtype, dev):
    """Single problems, n = 1, and the sizes either side of every internal switch: n + m = 128 / 129 (Gauss-Jordan
    kernel vs tensor-core block sweep), block counts 2..6 of the sweep, tile padding of the packed operator
    (n = 33, 255, 257), more problems than SMs."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad
    data = orc.make_exp1_data(n, B, seed=n, dtype=dtype)
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    g = torch.randn(B, n, 1, generator=torch.Generator().manual_seed(n), dtype=dtype)
    from tests.test_gpu_baseline_configs import _oracle, assert_borderline_and_rerun
    trace = []
    ref, rg = _oracle(data, control, g, dtype, trace=trace)
    ins = [t.to(dev) for t in data]
    sol = torch_solve_box_qp(*ins, control)
    grads = torch_solve_box_qp_grad(g.to(dev), sol["x"], sol["u"], sol["lams"], sol["nus"], ins[0], ins[2], ins[4],
                                    ins[5], sol["rho"])
    assert abs(sol["iter"] - ref["iter"]) <= 2
    f64 = dtype == torch.float64
    if sol["iter"] != ref["iter"]:
        # never a silent pass: the differing stop check must be borderline in the oracle's own arithmetic, and the
        # states are then compared at the SAME iteration (oracle re-run with max_iters cut to our exit)
        ref, rg = assert_borderline_and_rerun(data, control, g, dtype, ref["iter"], sol["iter"], trace)
    lim = 1e-8 if f64 else 2e-5
    for k in ("x", "z", "lams"):
        assert rel_err(sol[k].cpu().numpy(), ref[k].numpy()) <= lim, k
    for k, a, r in zip(("dQ", "dp"), grads[:2], rg[:2]):
        assert rel_err(a.cpu().numpy(), r.numpy()) <= (1e-7 if f64 else 1e-4), k


def test_equality_row_limits(dev):
    """m = 64 and m = 100 general equality rows against the oracle (the reference concatenates any m, :208-212; here
    the rows live in the padding of the factorisation, up to 256 of them); m = 257 is refused loudly."""
    from lqp_py_b200 import _abi
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp
    dtype = torch.float64
    n, B = 160, 3
    Q, p, _, _, lb, ub = orc.make_exp1_data(n, B, seed=5, dtype=dtype)
    gen = torch.Generator().manual_seed(8)
    control = box_qp_control(eps_abs=1e-6, eps_rel=1e-6)
    for m in (64, 100, 257):
        A = torch.randn(B, m, n, generator=gen, dtype=dtype)
        x0 = torch.rand(B, n, 1, generator=gen, dtype=dtype) - 0.5           # a point inside the box: feasible b
        b = A @ x0
        ins = [t.to(dev) for t in (Q, p, A, b, lb, ub)]
        if m == 257:
            with pytest.raises(_abi.LqpbError, match="256 equality rows"):
                torch_solve_box_qp(*ins, control)
            continue
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            ref = orc.solve(Q, p, A, b, lb, ub, control)
        finally:
            torch.set_default_dtype(prev)
        sol = torch_solve_box_qp(*ins, control)
        assert sol["iter"] == ref["iter"], (m, sol["iter"], ref["iter"])
        assert rel_err(sol["x"].cpu().numpy(), ref["x"].numpy()) <= 1e-8
        assert rel_err(sol["nus"].cpu().numpy(), ref["nus"].numpy()) <= 1e-7


def test_single_iteration_and_noncontiguous_inputs(dev):
    """max_iters = 1 (the loop exits on its first pass, iter = 0) and inputs that are views (transposed Q, strided
    vectors): the adapter makes them contiguous, results equal the contiguous call bit for bit."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp
    dtype = torch.float64
    Q, p, A, b, lb, ub = orc.make_exp1_data(40, 6, seed=12, dtype=dtype)
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5, max_iters=1)
    ref = _oracle_run((Q, p, A, b, lb, ub), control, torch.zeros_like(p), dtype)[0]
    ins = [t.to(dev) for t in (Q, p, A, b, lb, ub)]
    sol = torch_solve_box_qp(*ins, control)
    assert sol["iter"] == ref["iter"] == 0
    assert rel_err(sol["x"].cpu().numpy(), ref["x"].numpy()) <= 1e-8
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    base = torch_solve_box_qp(*ins, control)
    Qv = ins[0].transpose(1, 2)                                   # symmetric: same matrix, non-contiguous view
    wide = torch.stack((ins[1], ins[1] + 1), dim=3)               # (B, n, 1, 2): [..., 0] is a strided view of p
    pv = wide[..., 0]
    assert not Qv.is_contiguous() and not pv.is_contiguous()
    sol = torch_solve_box_qp(Qv, pv, ins[2], ins[3], ins[4], ins[5], control)
    assert sol["iter"] == base["iter"] and rel_err(sol["x"].cpu().numpy(), base["x"].cpu().numpy()) <= 1e-12


def test_experiment2_learning_loop_matches_oracle_layer(dev):
    """Experiment 2 (experiments/experiment_2.py:52-99; BASELINE config 5 scaled down): Linear(5, n) -> p_hat ->
    SolveBoxQP -> true cost -> backward -> SGD, 6 epochs.  The loop with the CUDA layer must trace the same loss
    curve and end at the same weights as the same loop with the layer built on the CPU oracle."""
    from lqp_py_b200 import sharding
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    from tests.test_sharding_cpu import _OracleLayer
    dtype = torch.float64
    n, B, nf = 60, 24, 5
    Q, _, A, b, lb, ub = orc.make_exp1_data(n, B, seed=2, dtype=dtype)
    gen = torch.Generator().manual_seed(4)
    feats = torch.randn(B, nf, generator=gen, dtype=dtype)
    p_true = (feats @ torch.randn(nf, n, generator=gen, dtype=dtype)).unsqueeze(2)
    control = box_qp_control(eps_abs=1e-8, eps_rel=1e-8)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        oracle_layer = lambda *a: _OracleLayer.apply(*a, dict(control))
        m_ref, h_ref = sharding.train_learn_p(oracle_layer, Q, p_true, A, b, lb, ub, feats, n_epochs=6, n_mini_batch=8,
                                              lr=5e-3, seed=1)
        qp = SolveBoxQP(control=dict(control))
        on = [t.to(dev) for t in (Q, p_true, A, b, lb, ub, feats)]
        m_gpu, h_gpu = sharding.train_learn_p(qp, *on, n_epochs=6, n_mini_batch=8, lr=5e-3, seed=1)
    finally:
        torch.set_default_dtype(prev)
    assert rel_err(np.array(h_gpu), np.array(h_ref)) <= 1e-7
    for a, r in zip(m_gpu.parameters(), m_ref.parameters()):
        assert rel_err(a.detach().cpu().numpy(), r.detach().numpy()) <= 1e-7


@pytest.mark.parametrize("backward", ["fixed_point", "kkt"])
def test_prepared_backward_equals_plain_backward(backward, dev):
    """The forward call queues the dl_dz-independent part of the backward (lqpb_forward_prep_*), .backward() then only
    substitutes and assembles (lqpb_backward_finish_*).  A second backward through the retained graph takes the
    plain, unprepared path: both must give bit-identical gradients.  Two graphs alive at once keep separate
    workspaces; a forward whose backward never runs leaves nothing behind."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    Q, p, A, b, lb, ub = orc.make_exp1_data(200, 24, seed=4, dtype=torch.float32)
    QP = SolveBoxQP(control=box_qp_control(eps_abs=1e-5, eps_rel=1e-5, backward=backward))
    g1 = torch.randn(p.shape, generator=torch.Generator().manual_seed(1)).to(dev)
    g2 = torch.randn(p.shape, generator=torch.Generator().manual_seed(2)).to(dev)
    ins_a = [t.to(dev).requires_grad_(True) for t in (Q, p, A, b, lb, ub)]
    ins_b = [t.to(dev).requires_grad_(True) for t in (Q, 2 * p, A, b, lb, ub)]
    xa = QP.forward(*ins_a)
    xb = QP.forward(*ins_b)                      # second graph before the first backward
    QP.forward(*[t.to(dev).requires_grad_(True) for t in (Q, 3 * p, A, b, lb, ub)])   # never differentiated
    xa.backward(g1, retain_graph=True)           # prepared path
    first = [t.grad.clone() for t in ins_a]
    for t in ins_a:
        t.grad = None
    xa.backward(g1)                              # plain path on the same graph
    for name, u, v in zip(("dQ", "dp", "dA", "db", "dlb", "dub"), first, [t.grad for t in ins_a]):
        assert torch.equal(u, v), name
    xb.backward(g2)
    ins_c = [t.to(dev).requires_grad_(True) for t in (Q, 2 * p, A, b, lb, ub)]
    xc = QP.forward(*ins_c)
    xc.backward(g2)
    for u, v in zip(ins_b, ins_c):
        assert torch.equal(u.grad, v.grad)


def test_unrolled_first_pass_tape_and_its_fallback_agree(dev, monkeypatch):
    """The unrolled solve records itself when it fits the first-pass tape; with a tape too short for the solve it
    falls back to solve + recording pass.  Same kernels, same arithmetic: x and every gradient bit-identical."""
    import lqp_py_b200.solve_box_qp_admm_torch as mod
    case = UnrollCase("exp1_n50_b4_f64")
    outs = []
    for cap in (256, 8):
        monkeypatch.setattr(mod, "_UNROLL_TAPE_CAP", cap)
        leaves = [None if t is None else t.to(dev).requires_grad_(True) for t in case.inputs()]
        x = mod.SolveBoxQP(control=dict(case.control)).forward(*leaves)
        x.backward(torch.from_numpy(case.z["dl_dz"]).to(dev))
        outs.append((x.detach(), [t.grad for t in leaves]))
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------ additions to the reference
def test_status_residuals_and_breakdown(dev):
    """What the reference drops (:235, :331): how the solve ended.  (a) a converged solve: status 1, every problem
    converged, residuals below their tolerances and equal to the oracle's last check; (b) max_iters cut short: status 2
    and the per-problem flags single out the problems that had not converged (checked against the oracle's own
    residuals at that iteration); (c) a NaN in p: status 4 at the first check instead of 10 000 silent iterations."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import (torch_solve_box_qp, STATUS_CONVERGED, STATUS_MAX_ITERS,
                                                     STATUS_BREAKDOWN)
    dtype = torch.float64
    Q, p, A, b, lb, ub = orc.make_exp1_data(60, 12, seed=3, dtype=dtype)
    ins = [t.to(dev) for t in (Q, p, A, b, lb, ub)]
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    sol = torch_solve_box_qp(*ins, control)
    assert sol["status"] == STATUS_CONVERGED and bool(sol["converged"].all())
    assert bool((sol["primal_residual"] < sol["primal_tolerance"]).all())
    assert bool((sol["dual_residual"] < sol["dual_tolerance"]).all())
    trace = []
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        ref = orc.solve(Q, p, A, b, lb, ub, dict(control), trace=trace)
        cut = box_qp_control(eps_abs=1e-5, eps_rel=1e-5, max_iters=21)        # checks at 0, 10, 20: stops unconverged
        tr2 = []
        ref2 = orc.solve(Q, p, A, b, lb, ub, dict(cut), trace=tr2)
    finally:
        torch.set_default_dtype(prev)
    assert sol["iter"] == ref["iter"]
    # the slowest problem's residual / tolerance ratio at the last check, as the oracle saw it
    worst = max(float((sol["primal_residual"] / sol["primal_tolerance"]).max()),
                float((sol["dual_residual"] / sol["dual_tolerance"]).max()))
    assert abs(worst - max(trace[-1][1], trace[-1][2])) <= 1e-6 * max(trace[-1][1], trace[-1][2])
    sol2 = torch_solve_box_qp(*ins, cut)
    assert sol2["status"] == STATUS_MAX_ITERS and sol2["iter"] == 20 == ref2["iter"]
    assert not bool(sol2["converged"].all())
    worst2 = max(float((sol2["primal_residual"] / sol2["primal_tolerance"]).max()),
                 float((sol2["dual_residual"] / sol2["dual_tolerance"]).max()))
    assert worst2 >= 1.0 and abs(worst2 - max(tr2[-1][1], tr2[-1][2])) <= 1e-6 * worst2
    bad = [t.clone() for t in ins]
    bad[1][3, 5, 0] = float("nan")
    sol3 = torch_solve_box_qp(*bad, control)
    assert sol3["status"] == STATUS_BREAKDOWN and sol3["iter"] <= 10
    with pytest.raises(ValueError, match="non-finite"):
        torch_solve_box_qp(*bad, box_qp_control(validate=True))
    skew = [t.clone() for t in ins]
    skew[0][0, 1, 2] += 0.5
    with pytest.raises(ValueError, match="not symmetric"):
        torch_solve_box_qp(*skew, box_qp_control(validate=True))


@pytest.mark.parametrize("n,dtype", [(40, torch.float64), (200, torch.float32), (500, torch.float32)])
def test_warm_start_saves_iterations(n, dtype, dev):
    """Warm start (an addition; reference :221-223 always starts from zero).  (a) restarting a converged solve from
    its own z, u stops at the first check with the same solution; (b) after a small change of p -- a learning step --
    the warm solve needs fewer iterations than the cold one and both agree to the solver tolerance; (c) BoxQPTH."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, BoxQPTH
    B = 16
    Q, p, A, b, lb, ub = [t.to(dev) for t in orc.make_exp1_data(n, B, seed=7, dtype=dtype)]
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    cold = torch_solve_box_qp(Q, p, A, b, lb, ub, control)
    again = torch_solve_box_qp(Q, p, A, b, lb, ub, control, z0=cold["z"], u0=cold["u"])
    assert again["status"] == 1 and again["iter"] <= 20 and again["iter"] < cold["iter"], (again["iter"], cold["iter"])
    scale = float(cold["x"].abs().max())
    assert float((again["x"] - cold["x"]).abs().max()) <= 5e-5 * scale
    p2 = p + 0.01 * torch.randn(p.shape, generator=torch.Generator().manual_seed(1), dtype=dtype).to(dev)
    cold2 = torch_solve_box_qp(Q, p2, A, b, lb, ub, control)
    warm2 = torch_solve_box_qp(Q, p2, A, b, lb, ub, control, z0=cold["z"], u0=cold["u"])
    assert warm2["status"] == 1 and warm2["iter"] < cold2["iter"], (warm2["iter"], cold2["iter"])
    assert float((warm2["x"] - cold2["x"]).abs().max()) <= 1e-3 * scale        # both within tol 1e-5 of the optimum
    with pytest.raises(ValueError):
        torch_solve_box_qp(Q, p2, A, b, lb, ub, control, z0=cold["z"])
    holder = BoxQPTH(Q, p, A, b, lb, ub, control)
    holder.solve()
    holder.update(p=p2)
    holder.solve(warm_start=True)
    assert holder.sol["iter"] == warm2["iter"] and torch.equal(holder.sol["x"], warm2["x"])


def test_two_devices_in_one_process():
    """The library keeps per-device state (function attributes, streams, events): a forward + backward on cuda:1 after
    one on cuda:0 in the same process, on the tensor-core path (fp32, n + m > 128).  Skips on a single-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    data = orc.make_exp1_data(200, 8, seed=2, dtype=torch.float32)
    g = torch.randn(8, 200, 1, generator=torch.Generator().manual_seed(3))
    outs = []
    for d in ("cuda:0", "cuda:1", "cuda:0"):
        ins = [t.to(d).requires_grad_(True) for t in data]
        x = SolveBoxQP(control=box_qp_control(eps_abs=1e-5, eps_rel=1e-5)).forward(*ins)
        x.backward(g.to(d))
        torch.cuda.synchronize(d)
        outs.append((x.detach().cpu(), ins[0].grad.cpu(), ins[1].grad.cpu()))
    for o in outs[1:]:
        for a, r in zip(o, outs[0]):
            assert torch.equal(a, r)


def test_experiment2_nccl_two_gpus():
    """The Experiment-2 learning loop data-parallel over NCCL (tools/exp2_nccl.py launched with torchrun on two GPUs):
    ranks end with identical weights and trace the single-process loss curve.  Skips on a single-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "exp2_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["world"] == 2 and out["ranks_hold_identical_weights"]
    assert out["loss_curve_rel_err_vs_single_process"] < 1e-6 and out["weights_rel_err_vs_single_process"] < 1e-6


def test_unrolled_adaptive_rho_iter_shorter_than_check(dev):
    """unroll=True with adaptive_rho_iter rounding below check_solved (reference :145-147 turns 4 into 1 at n = 30):
    every iteration may update rho from the residuals of a check made several updates ago (:237-250).  Round 1 refused
    this configuration; the stale check's state and the rho then in force are now kept across segments.  Compared with
    autograd through the oracle's loop.  The rho recursion is unstable by construction (the same stale ratio is applied
    at every iteration, rho runs into its clamp and amplifies round-off: the plain solve itself drifts from 1e-15 to
    1e-8 of the oracle between iterations 11 and 25), so the run is cut at 12 iterations: ten updates from the check
    at 0, the check at 10, and one update from that check with the rho then in force."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    dtype = torch.float64
    Q, p, A, b, lb, ub = orc.make_exp1_data(30, 4, seed=3, dtype=dtype)
    g = torch.randn(p.shape, generator=torch.Generator().manual_seed(5), dtype=dtype)
    control = box_qp_control(eps_abs=1e-6, eps_rel=1e-6, rho=100.0, adaptive_rho_iter=4, max_iters=12, unroll=True)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        leaves = [t.clone().requires_grad_(True) for t in (Q, p, A, b, lb, ub)]
        xr = orc.solve_unrolled(*leaves, dict(control))
        xr.backward(g)
        plain = orc.solve(Q, p, A, b, lb, ub, dict(control, unroll=False))
    finally:
        torch.set_default_dtype(prev)
    assert plain["factorisations"] > 3                     # several updates between two checks
    ins = [t.to(dev).requires_grad_(True) for t in (Q, p, A, b, lb, ub)]
    x = SolveBoxQP(control=control).forward(*ins)
    x.backward(g.to(dev))
    assert rel_err(x.detach().cpu().numpy(), xr.detach().numpy()) <= 1e-8
    for a, r, nm in zip(ins, leaves, ("dQ", "dp", "dA", "db", "dlb", "dub")):
        assert rel_err(a.grad.cpu().numpy(), r.grad.numpy()) <= 1e-6, nm


@pytest.mark.parametrize("n,dtype", [(40, torch.float64), (200, torch.float32)])
def test_per_problem_rho_tensor_in_control(n, dtype, dev):
    """The reference broadcasts a (B,1,1) tensor in control['rho'] (:200: only None triggers the automatic choice), e.g.
    the `rho` of an earlier solve fed back.  Feeding the automatic rho back must reproduce that solve bit for bit
    (same kernels, same per-problem rho), through the functional API and through the layer."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP, torch_solve_box_qp
    Q, p, A, b, lb, ub = [t.to(dev) for t in orc.make_exp1_data(n, 6, seed=9, dtype=dtype)]
    auto = torch_solve_box_qp(Q, p, A, b, lb, ub, box_qp_control(eps_abs=1e-5, eps_rel=1e-5))
    assert torch.is_tensor(auto["rho"]) and auto["rho"].shape == (6, 1, 1)
    again = torch_solve_box_qp(Q, p, A, b, lb, ub, box_qp_control(eps_abs=1e-5, eps_rel=1e-5, rho=auto["rho"]))
    assert again["iter"] == auto["iter"] and torch.equal(again["x"], auto["x"])
    assert torch.is_tensor(again["rho"]) and torch.equal(again["rho"], auto["rho"])
    ins = [t.clone().requires_grad_(True) for t in (Q, p, A, b, lb, ub)]
    x = SolveBoxQP(control=box_qp_control(eps_abs=1e-5, eps_rel=1e-5, rho=auto["rho"])).forward(*ins)
    x.backward(torch.ones_like(x))
    assert torch.equal(x.detach(), auto["x"]) and ins[1].grad is not None


def test_fused_block_sweep_equals_per_phase_kernels():
    """The fused persistent factorisation kernel (csrc/tcfused.cu) and the per-phase kernels (csrc/tcfactor.cu) run the same
    arithmetic in the same order: forward solution and all gradients must agree BIT FOR BIT.  The switch is read once per
    process, so each form runs in its own child (tools/tc_fused_ab.py); shapes: 3 and 4 block rows, with equality rows."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dz, B in ((300, 8), (500, 6)):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "tc_fused_ab.py"), str(dz), str(B)], capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "bit-identical" in r.stdout and '"finite": true' in r.stdout, r.stdout[-2000:]


def test_blocked_sweeps_beyond_the_baseline_shapes():
    """Shapes the golden / baseline sets do not reach, against the CPU oracle (tools/big_shape_check.py): 6 and 8 block rows
    of the fp64 DMMA sweep (dz = 700 / 1000: transposed PANEL jobs at every level) and 20 equality rows inside the blocked
    sweeps (the "hard" generator at n = 400), fp64 at 1e-8 / 1e-7."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "big_shape_check.py")], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ALL OK" in r.stdout and "FAIL" not in r.stdout.replace("FAILURES", ""), r.stdout[-2000:]


def test_prefetched_inputs_equal_the_host_path(dev):
    """SolveBoxQP.prefetch (upload of an announced batch on a copy stream) must not change a single bit: the later forward
    finds the device copies by the identity of the host tensors and takes the device-pointer path; a batch modified after
    the announcement is uploaded again."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP, prefetch_inputs, _PREFETCH
    n, B = 300, 8
    raw = [t.pin_memory() for t in orc.make_exp1_data(n, B, seed=5, dtype=torch.float32)]
    g = torch.randn(B, n, 1, generator=torch.Generator().manual_seed(1)).pin_memory()

    def run(announce):
        ins = [t.detach().requires_grad_(j < 2) for j, t in enumerate(raw)]
        if announce:
            assert prefetch_inputs(*raw)
        x = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5)).forward(*ins)
        x.backward(g)
        return x.detach().clone(), ins[0].grad.clone(), ins[1].grad.clone()
    base = run(False)
    pre = run(True)
    assert not _PREFETCH, "the announced batch was not consumed"
    for a_, b_ in zip(base, pre):
        assert a_.device.type == "cpu" and torch.equal(a_, b_)
    # stale announcement: the version counter of Q moves, the copy must not be used
    assert prefetch_inputs(*raw)
    raw[1].mul_(1.5)
    moved = run(False)
    fresh = [t.clone().pin_memory() for t in raw]
    ins = [t.detach().requires_grad_(j < 2) for j, t in enumerate(fresh)]
    x = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5)).forward(*ins)
    assert torch.equal(moved[0], x.detach())
    assert not torch.equal(moved[0], base[0])
    _PREFETCH.clear()


def test_solve_ahead_equals_the_host_path(dev):
    """SolveBoxQP.solve_ahead (upload + forward of announced batches on a worker thread and a second stream, two batches
    in flight) must not change a single bit of x, dQ, dp; a control changed after the announcement, or a batch modified
    since, must drop the speculated solution."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP, _PREFETCH
    n, B = 300, 8
    sets = [[t.pin_memory() for t in orc.make_exp1_data(n, B, seed=5 + k, dtype=torch.float32)] for k in range(4)]
    g = torch.randn(B, n, 1, generator=torch.Generator().manual_seed(1)).pin_memory()
    ctl = lambda **kw: box_qp_control(eps_rel=1e-5, eps_abs=1e-5, **kw)

    def step(QP, raw):
        ins = [t.detach().requires_grad_(j < 2) for j, t in enumerate(raw)]
        x = QP.forward(*ins)
        return x, ins

    def finish(x, ins):
        x.backward(g)
        return x.detach().clone(), ins[0].grad.clone(), ins[1].grad.clone()
    base = [finish(*step(SolveBoxQP(control=ctl()), raw)) for raw in sets]
    for backward in ("fixed_point", "kkt"):
        QP = SolveBoxQP(control=ctl(backward=backward))
        ref = base if backward == "fixed_point" else [finish(*step(SolveBoxQP(control=ctl(backward="kkt")), raw)) for raw in sets]
        # the loop of bench.py's e2e leg: two batches announced ahead, uploads / solves started by the backward
        assert QP.solve_ahead(*sets[0]) and QP.solve_ahead(*sets[1])
        out = []
        for k in range(len(sets)):
            x, ins = step(QP, sets[k])
            if k + 2 < len(sets):
                assert QP.solve_ahead(*sets[k + 2])
            out.append(finish(x, ins))
        assert not _PREFETCH, "an announced batch was not consumed"
        for r_, o_ in zip(ref, out):
            for a_, b_ in zip(r_, o_):
                assert a_.device.type == "cpu" and torch.equal(a_, b_)
    # another control at the consuming call: the speculated solution must not be used
    QP = SolveBoxQP(control=ctl())
    assert QP.solve_ahead(*sets[0])
    QP.control['eps_abs'] = 1e-3
    QP.control['eps_rel'] = 1e-3
    x, ins = step(QP, sets[0])
    loose = finish(x, ins)
    want = finish(*step(SolveBoxQP(control=box_qp_control(eps_rel=1e-3, eps_abs=1e-3)), sets[0]))
    assert torch.equal(loose[0], want[0]) and not torch.equal(loose[0], base[0][0])
    assert not _PREFETCH
    # a batch modified after the announcement is solved afresh
    QP = SolveBoxQP(control=ctl())
    assert QP.solve_ahead(*sets[1])
    sets[1][1].mul_(1.5)
    moved = finish(*step(QP, sets[1]))
    fresh = [t.clone().pin_memory() for t in sets[1]]
    want = finish(*step(SolveBoxQP(control=ctl()), fresh))
    assert torch.equal(moved[0], want[0]) and not torch.equal(moved[0], base[1][0])
    _PREFETCH.clear()
    # no gradient wanted: nothing of the backward is prepared, the forward result is the same
    QP = SolveBoxQP(control=ctl())
    assert QP.solve_ahead(*sets[2], requires_grad=False)
    with torch.no_grad():
        x = QP.forward(*sets[2])
    assert torch.equal(x, base[2][0])
    assert not _PREFETCH
    # the other factorisation paths: fp64 blocked sweep (n + m > 128) and the small-problem path (no prepared backward)
    for n2, dt in ((200, torch.float64), (40, torch.float64), (40, torch.float32)):
        sets2 = [[t.pin_memory() for t in orc.make_exp1_data(n2, B, seed=11 + k, dtype=dt)] for k in range(3)]
        g2 = torch.randn(B, n2, 1, generator=torch.Generator().manual_seed(2), dtype=dt).pin_memory()

        def run2(QP, raw, ahead):
            ins = [t.detach().requires_grad_(j < 2) for j, t in enumerate(raw)]
            x = QP.forward(*ins)
            if ahead is not None:
                assert QP.solve_ahead(*ahead)
            x.backward(g2)
            return x.detach().clone(), ins[0].grad.clone(), ins[1].grad.clone()
        ref2 = [run2(SolveBoxQP(control=ctl()), raw, None) for raw in sets2]
        QP = SolveBoxQP(control=ctl())
        assert QP.solve_ahead(*sets2[0]) and QP.solve_ahead(*sets2[1])
        out2 = [run2(QP, sets2[k], sets2[k + 2] if k + 2 < len(sets2) else None) for k in range(len(sets2))]
        assert not _PREFETCH
        for r_, o_ in zip(ref2, out2):
            for a_, b_ in zip(r_, o_):
                assert torch.equal(a_, b_), (n2, dt)


def test_pageable_inputs_are_staged_bit_identically(dev):
    """Ordinary (pageable) CPU tensors -- what the reference's callers hold -- are uploaded through page-locked staging
    memory (caller's thread for a plain call, worker thread for announced batches); x, dQ, dp must equal the page-locked
    run bit for bit in the plain, prefetched and solved-ahead forms."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200 import solve_box_qp_admm_torch as M
    n, B = 300, 16
    raw = [orc.make_exp1_data(n, B, seed=21 + k, dtype=torch.float32) for k in range(3)]
    assert M._needs_staging(raw[0][0]) and not M._needs_staging(raw[0][0].pin_memory())
    g = torch.randn(B, n, 1, generator=torch.Generator().manual_seed(3))
    ctl = lambda: box_qp_control(eps_rel=1e-5, eps_abs=1e-5)

    def run(QP, ts, announce=None):
        ins = [t.detach().requires_grad_(j < 2) for j, t in enumerate(ts)]
        x = QP.forward(*ins)
        if announce is not None:
            announce()
        x.backward(g)
        return x.detach().clone(), ins[0].grad.clone(), ins[1].grad.clone()
    want = [run(M.SolveBoxQP(control=ctl()), [t.pin_memory() for t in ts]) for ts in raw]
    plain = [run(M.SolveBoxQP(control=ctl()), ts) for ts in raw]
    QP = M.SolveBoxQP(control=ctl())
    pre = [run(QP, raw[k], (lambda k=k: QP.prefetch(*raw[k + 1])) if k + 1 < len(raw) else None) for k in range(len(raw))]
    assert not M._PREFETCH
    QP = M.SolveBoxQP(control=ctl())
    assert QP.solve_ahead(*raw[0]) and QP.solve_ahead(*raw[1])
    ahead = [run(QP, raw[k], (lambda k=k: QP.solve_ahead(*raw[k + 2])) if k + 2 < len(raw) else None) for k in range(len(raw))]
    assert not M._PREFETCH
    for got in (plain, pre, ahead):
        for w_, g_ in zip(want, got):
            for a_, b_ in zip(w_, g_):
                assert a_.device.type == "cpu" and torch.equal(a_, b_)


@pytest.mark.parametrize("split", ["0", None])
def test_adaptive_rho_refactorisation_in_the_streamed_regimes(dev, split, monkeypatch):
    """A badly chosen rho (100, as in the golden case adapt_rho100 at n = 60) at a STREAMED size: the iteration kernel exits
    with a refactorisation request, the host refactors and relaunches it mid-solve (i0 > 0) -- through the one-CTA-per-problem
    kernel (LQPB_ITER_SPLIT=0) and through the cluster-split kernel a small batch takes by default.  fp64 against the oracle:
    same number of factorisations, iteration count within 2, x / duals / rho at 1e-8 (1e-7 where rho amplifies round-off)."""
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import _solve_device
    if split is None:
        monkeypatch.delenv("LQPB_ITER_SPLIT", raising=False)
    else:
        monkeypatch.setenv("LQPB_ITER_SPLIT", split)
    n, B = 320, 3
    data = orc.make_exp1_data(n, B, seed=31, dtype=torch.float64)
    control = box_qp_control(eps_abs=1e-8, eps_rel=1e-8, rho=100.0, adaptive_rho_iter=40)
    ref = orc.solve(*data, dict(control))
    sol = _solve_device(*[t.to(dev) for t in data], dict(control))      # (torch_solve_box_qp's dict + n_factor)
    assert sol["n_factor"] == ref["factorisations"] >= 2, "the case must refactorise at least once"
    assert abs(sol["iter"] - ref["iter"]) <= 2
    assert torch.is_tensor(sol["rho"]) and rel_err(sol["rho"].cpu().numpy(), ref["rho"].numpy()) <= 1e-8
    for k, t in (("x", 1e-8), ("z", 1e-8), ("lams", 1e-7), ("nus", 1e-7), ("u", 1e-7)):
        e = rel_err(sol[k].cpu().numpy(), ref[k].numpy())
        assert e <= t, f"{k}: {e:.2e} > {t:.1e}"

"""CPU oracle for the batched ADMM box-QP path of ipo-lab/lqp_py.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lqp_py_b200/`` may import this
module; it is the checker for ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

What it is: an independent restatement (torch CPU tensors, batched LAPACK
through ``torch.linalg`` -- the same numerical backend the reference uses, so
the timing of this port is representative of the reference's CPU path) of

* the forward ADMM solver   reference ``lqp_py/solve_box_qp_admm_torch.py:108-333``
* the fixed-point backward  reference ``lqp_py/solve_box_qp_admm_torch.py:349-432``
* the KKT backward          reference ``lqp_py/solve_box_qp_admm_torch.py:435-584``
* the cached-factor LU op   reference ``lqp_py/lu_layer.py:5-58``
* the experiment generators reference ``experiments/utils.py:35-61,64-131``

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference``) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every fixture.

The code is organised as three small stages (``prepare`` -> ``iterate`` ->
``conclude``) rather than one long function, and keeps every quirk of the
reference that is observable in the outputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

INF = float("inf")
TINY = 1e-16  # reference :229 zero_clamp


# --------------------------------------------------------------------------
# settings
# --------------------------------------------------------------------------
@dataclass
class Settings:
    """Derived settings, exactly the keys the reference *reads*
    (``solve_box_qp_admm_torch.py:134-154``), not the keys its factory writes."""
    max_iters: int
    eps_abs: float
    eps_rel: float
    check_every: int
    rho: Optional[float]
    rho_min: float
    rho_max: float
    adaptive: bool
    adaptive_tol: float
    adaptive_every: int
    adaptive_until: int
    adaptive_floor: float
    scale: bool
    beta: Optional[float]
    verbose: bool


def derive_settings(control: dict, n_x: int) -> Settings:
    """reference :134-154.  Note the read keys 'check_solved' and
    'adaptive_max_iter' differ from the factory's 'check_terimnation' and
    'adaptive_rho_max_iter' (control.py:8,15) so factory values are ignored."""
    g = control.get
    check = g("check_solved", max(round((n_x ** 0.5) / 10) * 10, 1))
    every = g("adaptive_rho_iter", 100)
    every = max(round(every / check) * check, 1)
    return Settings(
        max_iters=g("max_iters", 10_000),
        eps_abs=max(g("eps_abs", 1e-3), 1e-12),
        eps_rel=max(g("eps_rel", 1e-3), 1e-12),
        check_every=check,
        rho=g("rho", None),
        rho_min=g("rho_min", 1e-6),
        rho_max=g("rho_max", 1e6),
        adaptive=g("adaptive_rho", False),
        adaptive_tol=g("adaptive_rho_tol", 5),
        adaptive_every=every,
        adaptive_until=g("adaptive_max_iter", 1000),
        adaptive_floor=g("adaptive_rho_threshold", 1e-5),
        scale=g("scale", False),
        beta=g("beta"),
        verbose=g("verbose", False),
    )


def default_control(**kw) -> dict:
    """What ``box_qp_control(**kw)`` produces (reference ``control.py:1-24``)."""
    c = dict(max_iters=10_000, eps_abs=1e-3, eps_rel=1e-3, check_terimnation=None,
             rho=None, rho_min=1e-6, rho_max=1e6, adaptive_rho=True,
             adaptive_rho_tol=10, adaptive_rho_iter=100, adaptive_rho_max_iter=1000,
             adaptive_rho_threshold=1e-5, verbose=False, scale=True, unroll=False,
             beta=None, backward="fixed_point")
    if "check_solved" in kw:  # factory stores it under the misspelt key
        c["check_terimnation"] = kw.pop("check_solved")
    c.update(kw)
    return c


# --------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------
def _guard_zero(norms: torch.Tensor) -> torch.Tensor:
    """reference :164-168 / :182-186 -- replace non-positive norms by
    max(mean of the row of norms, 1e-6)."""
    bad = norms <= 0.0
    if bool(bad.any()):
        floor = norms.mean(dim=1).clamp(min=1e-6).unsqueeze(1)
        norms = torch.where(bad, norms.clamp(min=floor), norms)
    return norms


def prepare(Q, p, A, b, lb, ub, st: Settings):
    """Scaling, rho selection, KKT factorisation (reference :124-223)."""
    B, n = Q.shape[0], p.shape[1]
    m = 0 if A is None else A.shape[1]
    dt = p.dtype
    p_inf = p.abs().amax(dim=1, keepdim=True)                       # :127 (unscaled p)
    has_lb = bool(lb.max() > -INF)                                  # :129
    has_ub = bool(ub.min() < INF)                                   # :130
    boxed = has_lb or has_ub
    rho = st.rho if boxed else 0                                    # :157-158

    if st.scale:
        colmax = _guard_zero(Q.abs().amax(dim=1))                   # :163-168
        D = (1.0 / colmax).sqrt()                                   # :170
        beta = st.beta
        if beta is None:                                            # :171-174
            qs = torch.quantile(D, torch.tensor([0.10, 0.90], dtype=D.dtype), dim=1)
            beta = (1 - qs[0] / qs[1]).unsqueeze(1)
        D = (1 - beta) * D + beta * D.mean(dim=1, keepdim=True)     # :175
        Q = D.unsqueeze(2) * Q * D.unsqueeze(1)                     # :176
        p = D.unsqueeze(2) * p                                      # :177
        E = 1.0
        if m:
            A = A * D.unsqueeze(1)                                  # :180
            rown = _guard_zero(A.abs().amax(dim=2))                 # :181-186
            E = (1.0 / rown).unsqueeze(2)
            A, b = E * A, E * b                                     # :189-190
        D = D.unsqueeze(2)
        if boxed:
            lb, ub = lb / D, ub / D                                 # :193-194
    else:
        D, E = 1.0, 1.0                                             # :196-197

    if rho is None:                                                 # :200-203
        rho = (torch.linalg.matrix_norm(Q, keepdim=True) / n ** 0.5).clamp(st.rho_min, st.rho_max)

    eye = torch.eye(n, dtype=dt).unsqueeze(0)
    K = Q + rho * eye                                               # :207
    if m:                                                           # :208-212
        K = torch.cat((torch.cat((K, A.transpose(1, 2)), 2),
                       torch.cat((A, torch.zeros(B, m, m, dtype=dt)), 2)), 1)
    LU, piv = torch.linalg.lu_factor(K)                             # :215
    return dict(B=B, n=n, m=m, dt=dt, Q=Q, p=p, A=A, b=b, lb=lb, ub=ub, D=D, E=E, rho=rho,
                K=K, LU=LU, piv=piv, eye=eye, p_inf=p_inf, has_lb=has_lb, has_ub=has_ub)


def iterate(w: dict, st: Settings):
    """The ADMM loop (reference :235-313).  All problems run in lock step and
    stop together (global ``all``), the adaptive-rho step uses the residuals of
    the previous check and does not rescale u."""
    B, n, m, dt = w["B"], w["n"], w["m"], w["dt"]
    Q, p, A, b, lb, ub, D = w["Q"], w["p"], w["A"], w["b"], w["lb"], w["ub"], w["D"]
    rho, K, LU, piv = w["rho"], w["K"], w["LU"], w["piv"]
    x = torch.zeros(B, n, 1, dtype=dt)
    z = torch.zeros_like(x)
    u = torch.zeros_like(x)
    tiny = torch.ones(1) * TINY                                     # :230 (default dtype)
    floor = torch.ones(1) * st.adaptive_floor                       # :150
    res_p = res_d = scale_p = scale_d = None
    wants_update = st.adaptive
    sol = None
    factorisations = 1
    i = 0
    for i in range(st.max_iters):
        if st.adaptive and i % st.adaptive_every == 0 and 0 < i < st.adaptive_until:   # :237
            if bool(torch.as_tensor(wants_update).any()):
                ratio = ((res_p / scale_p).clamp(min=TINY) / (res_d / scale_d).clamp(min=TINY)) ** 0.5
                if bool((ratio > st.adaptive_tol).any()) or bool((ratio < 1 / st.adaptive_tol).any()):
                    rho = torch.where(wants_update, rho * ratio, rho * torch.ones_like(ratio))
                    rho = rho.clamp(st.rho_min, st.rho_max)         # :248-250
                    K = K.clone()                                   # (no in-place write into a tensor autograd saved)
                    K[:, :n, :n] = Q + rho * w["eye"]               # :252
                    LU, piv = torch.linalg.lu_factor(K)             # :254
                    factorisations += 1
        rhs = -p + rho * (z - u)                                    # :259-262
        if m:
            rhs = torch.cat((rhs, b), 1)
        sol = torch.linalg.lu_solve(LU, piv, rhs)                   # :267
        x = sol[:, :n, :]
        z_old = z
        z = x + u                                                   # :272-276
        if w["has_lb"]:
            z = torch.maximum(z, lb)
        if w["has_ub"]:
            z = torch.minimum(z, ub)
        r = x - z                                                   # :279
        s = rho * (z - z_old)                                       # :280
        u = u + r                                                   # :282
        if i % st.check_every == 0:                                 # :285-313
            amax = lambda t: t.abs().amax(dim=1, keepdim=True)
            res_p, res_d = amax(D * r), amax(D * s)
            if st.verbose:
                print(f"iteration = {i}")
                print(f"|| primal_error|| = {res_p.max().item():.10f}")
                print(f"|| dual_error|| = {res_d.max().item():.10f}")
            scale_p = torch.maximum(torch.maximum(amax(D * x), amax(D * z)), tiny)
            scale_d = torch.maximum(torch.maximum(torch.maximum(
                amax(rho * D * u), amax(torch.matmul(Q, x) / D)), w["p_inf"]), tiny)
            tol_p = st.eps_abs + st.eps_rel * scale_p
            tol_d = st.eps_abs + st.eps_rel * scale_d
            done = (res_p < tol_p) & (res_d < tol_d)
            if w.get("trace") is not None:      # test aid: how far the slowest problem is from the stop threshold
                w["trace"].append((i, float((res_p / tol_p).max()), float((res_d / tol_d).max())))
            wants_update = (res_p > torch.maximum(tol_p, floor)) | (res_d > torch.maximum(tol_d, floor))
            if bool(done.all()):
                break
    w.update(rho=rho)
    return x, z, u, sol, i, factorisations


def conclude(w: dict, x, z, u, sol, i):
    """Undo the scaling and split the duals (reference :315-331)."""
    D, E, rho, n, m = w["D"], w["E"], w["rho"], w["n"], w["m"]
    x, z, u = D * x, D * z, u / D
    y = u * rho
    lams = torch.cat((torch.relu(-y), torch.relu(y)), 1)           # :320-323 (threshold(.,0,0) == relu)
    nus = sol[:, n:n + m, :] * E if m else None                    # :327
    return {"x": x, "z": z, "u": u, "lams": lams, "nus": nus, "rho": rho, "iter": i}


def solve(Q, p, A, b, lb, ub, control: dict, trace=None) -> dict:
    """Forward solve; same dict as the reference's ``torch_solve_box_qp`` (:331)
    plus ``'factorisations'`` (number of LU factorisations, 1 + rho updates).  ``trace`` (a list, test aid)
    receives ``(i, max_b primal/tol_primal, max_b dual/tol_dual)`` of every stop check (:301-309)."""
    with torch.no_grad():
        st = derive_settings(control, p.shape[1])
        w = prepare(Q, p, A, b, lb, ub, st)
        w["trace"] = trace
        x, z, u, sol, i, nfac = iterate(w, st)
        out = conclude(w, x, z, u, sol, i)
    out["factorisations"] = nfac
    return out


def solve_unrolled(Q, p, A, b, lb, ub, control: dict):
    """``unroll=True`` (reference :13-15, :216-217, :264-265, :328-329): the same three stages with autograd
    recording, returning only ``x``.  The reference differentiates each linear solve with ``TorchLULayer``
    (lu_layer.py:41-58: ``dx = M^-1 (-g)``, ``dl_dA = dx x^T``, ``dl_db = -dx``), which is the exact adjoint of
    ``M^-1 rhs`` for the symmetric KKT matrix, so letting torch differentiate ``lu_factor`` / ``lu_solve`` here
    gives the same gradients to round-off (pinned by tests/golden/unroll/*.npz)."""
    st = derive_settings(control, p.shape[1])
    w = prepare(Q, p, A, b, lb, ub, st)
    x, z, u, sol, i, nfac = iterate(w, st)
    return w["D"] * x                                               # :316, :328-329


# --------------------------------------------------------------------------
# backward (implicit differentiation of the ADMM fixed point)
# --------------------------------------------------------------------------
def grad(dl_dz, x, u, lams, nus, Q, A, lb, ub, rho):
    """reference :349-432.  Returns (dQ, dp, dA, db, dlb, dub)."""
    with torch.no_grad():
        B, n = Q.shape[0], Q.shape[1]
        m = 0 if A is None else A.shape[1]
        dt = x.dtype
        if rho is None:                                             # :356-357
            rho = 1.0
        t = x + u
        free = torch.ones(B, n, 1, dtype=dt)                        # :363-365
        free[t > ub] = 0
        free[t < lb] = 0
        g = dl_dz * free                                            # :368
        rho_col = rho.reshape(B, 1) if torch.is_tensor(rho) else rho
        top = free * Q                                              # :378 (row mask)
        idx = torch.arange(n)
        top[:, idx, idx] = top[:, idx, idx] + rho_col * (1 - free.squeeze(2))   # :380-383
        rhs = -g
        if m:                                                       # :384-389
            top = torch.cat((top, free * A.transpose(1, 2)), 2)
            top = torch.cat((top, torch.cat((A, torch.zeros(B, m, m, dtype=dt)), 2)), 1)
            rhs = torch.cat((rhs, torch.zeros(B, m, 1, dtype=dt)), 1)
        jdx = torch.arange(n + m)
        top[:, jdx, jdx] = top[:, jdx, jdx] + 1e-8                  # :392
        d = torch.linalg.solve(top, rhs)                            # :393
        dv = d[:, :n, :]
        half = torch.matmul(0.5 * dv, x.transpose(1, 2))            # :403-404
        dQ = half + half.transpose(1, 2)
        dA = db = None
        resid = -dl_dz - torch.matmul(Q, dv)                        # :417
        if m:
            dnu = d[:, n:, :]
            db = -dnu                                               # :411
            dA = torch.matmul(dnu, x.transpose(1, 2)) + torch.matmul(nus, dv.transpose(1, 2))  # :412
            resid = resid - torch.matmul(A.transpose(1, 2), dnu)    # :419
        den = rho * u                                               # :420-421
        den = torch.where(den == 0, torch.ones_like(den), den)
        dlam = resid / den
        dlb = dlam * lams[:, :n, :]                                 # :426
        dub = -dlam * lams[:, n:2 * n, :]                           # :427
    return dQ, dv, dA, db, dlb, dub


# --------------------------------------------------------------------------
# backward, KKT mode (implicit differentiation of the optimality conditions)
# --------------------------------------------------------------------------
def grad_kkt(dl_dz, x, lams, nus, Q, A, lb, ub):
    """reference :435-584 (``torch_solve_box_qp_grad_kkt`` and its helpers).  Returns
    (dQ, dp, dA, db, dlb, dub); dlb / dub are None without finite lower / upper bounds.
    Builds the same dense (n + 2n + m) system as the reference, so a one-sided box (infinite
    slacks) yields NaN exactly like the reference does."""
    with torch.no_grad():
        B, n = Q.shape[0], Q.shape[1]
        m = 0 if A is None else A.shape[1]
        has_lb = bool(lb.max() > -INF)                              # :439-441
        has_ub = bool(ub.min() < INF)
        boxed = has_lb or has_ub
        xt = x.transpose(1, 2)
        rows = [Q]
        k = 0
        if boxed:                                                   # :444-451
            k = 2 * n
            eye = torch.eye(n)
            G = torch.cat((-eye, eye)).unsqueeze(0) * torch.ones(B, 1, 1)
            slack = torch.clamp(torch.cat((-lb, ub), 1) - torch.matmul(G, x), 10 ** -8)
            lams = torch.clamp(lams, 10 ** -8)
            rows.append(G.transpose(1, 2) * lams.transpose(1, 2))   # :477 / :485
        if m:
            rows.append(A.transpose(1, 2))
        lhs = torch.cat(rows, 2)
        if boxed:                                                   # :478 / :486
            blk = [G, torch.diag_embed(-slack.squeeze(2))]
            if m:
                blk.append(torch.zeros(B, k, m))
            lhs = torch.cat((lhs, torch.cat(blk, 2)), 1)
        if m:                                                       # :482 / :487
            blk = [A]
            if boxed:
                blk.append(torch.zeros(B, m, k))
            blk.append(torch.zeros(B, m, m))
            lhs = torch.cat((lhs, torch.cat(blk, 2)), 1)
        rhs = torch.cat((-dl_dz, torch.zeros(B, k + m, 1)), 1)       # :500-504
        d = torch.linalg.solve(lhs, rhs)
        dx = d[:, :n, :]
        half = torch.matmul(0.5 * dx, xt)                           # :536-537
        dQ = half + half.transpose(1, 2)
        dA = db = dlb = dub = None
        if m:                                                       # :550-552
            dnu = d[:, n + k:, :]
            dA = torch.matmul(dnu, xt) + torch.matmul(nus, dx.transpose(1, 2))
            db = -dnu
        if boxed:
            dh = -lams * d[:, n:n + k, :]                           # :545
            if has_lb and has_ub:                                   # :572-579
                dlb, dub = -dh[:, :n, :], dh[:, n:, :]
            elif has_lb:
                dlb = -dh[:, :n, :]
            else:
                dub = dh[:, :n, :]
    return dQ, dx, dA, db, dlb, dub


def solve_and_grad(Q, p, A, b, lb, ub, control, dl_dz):
    """forward + fixed-point backward, the benchmarked pair
    (reference ``experiments/experiment_1.py:70-77``)."""
    ctl = dict(control)
    if not (bool(lb.max() > -INF) or bool(ub.min() < INF)):         # layer :33-38
        ctl["rho"] = 0
    sol = solve(Q, p, A, b, lb, ub, ctl)
    grads = grad(dl_dz, sol["x"], sol["u"], sol["lams"], sol["nus"], Q, A, lb, ub, sol["rho"])
    return sol, grads


# --------------------------------------------------------------------------
# lu_layer (reference lqp_py/lu_layer.py:5-58)
# --------------------------------------------------------------------------
def lu_forward(M, rhs, LU=None, piv=None):
    if LU is None or piv is None:
        LU, piv = torch.linalg.lu_factor(M)
    return torch.linalg.lu_solve(LU, piv, rhs), LU, piv


def lu_backward(LU, piv, x, dl_dx):
    """dl_dA = dx x^T, dl_db = -dx with dx = A^-1 (-dl_dx) (valid for symmetric A)."""
    dx = torch.linalg.lu_solve(LU, piv, -dl_dx)
    return torch.matmul(dx, x.transpose(1, 2)), -dx


# --------------------------------------------------------------------------
# generators (reference experiments/utils.py:35-61, 64-131) -- RNG call order matters
# --------------------------------------------------------------------------
def make_exp1_data(n_x, n_batch, n_samples=None, seed=0, dtype=None):
    """``create_qp_data`` without the G/h extras.  manual_seed -> randn(L) ->
    randn(p) -> rand(lb) -> rand(ub), all in the current default dtype."""
    prev = torch.get_default_dtype()
    if dtype is not None:
        torch.set_default_dtype(dtype)
    try:
        n_samples = 2 * n_x if n_samples is None else n_samples
        torch.manual_seed(seed)
        L = torch.randn(n_batch, n_samples, n_x)
        Q = torch.matmul(L.transpose(1, 2), L) / n_samples
        p = torch.randn(n_batch, n_x, 1)
        A = torch.ones(n_batch, 1, n_x)
        b = torch.ones(n_batch, 1, 1)
        lb = -(torch.rand(n_batch, n_x, 1) * (2 - 1) + 1)
        ub = torch.rand(n_batch, n_x, 1) * (2 - 1) + 1
    finally:
        torch.set_default_dtype(prev)
    return Q, p, A, b, lb, ub


def _sparse_rows(n_x, prob):
    m = round(n_x ** 0.5)
    A = np.zeros((m, n_x))
    for r in range(m):
        vals = np.random.normal(size=(1, n_x))
        keep = np.zeros(1)
        while keep.sum() == 0:
            keep = np.random.binomial(1, prob, size=(1, n_x))
        A[r, :] = vals * keep
    return A


def make_hard_data(n_x, prob, seeds, dtype=torch.float64):
    """``generate_hard_qp_torch``: sparse M^T M + 0.01 I, m = round(sqrt(n)) sparse
    equality rows, bounds around a feasible x0."""
    out = [[] for _ in range(6)]
    for s in seeds:
        np.random.seed(s)
        M = np.random.normal(size=(n_x, n_x))
        M = M * np.random.binomial(1, prob, size=(n_x, n_x))
        Q = M.T @ M + 1e-2 * np.eye(n_x)
        p = np.random.normal(size=(n_x, 1))
        x0 = np.random.normal(size=(n_x, 1))
        s_lb = -np.random.uniform(size=(n_x, 1))
        s_ub = np.random.uniform(size=(n_x, 1))
        A = _sparse_rows(n_x, prob)
        for k, v in enumerate((Q, p, A, A @ x0, x0 + s_lb, x0 + s_ub)):
            out[k].append(v)
    return tuple(torch.tensor(np.stack(v), dtype=dtype) for v in out)

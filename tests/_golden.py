"""Helpers shared by the CPU (oracle) and GPU (C-ABI) parity tests: load a golden
fixture written by tests/golden/make_golden.py and compare a solver's outputs with it."""
import glob
import json
import os

import numpy as np
import torch

from oracle import box_qp_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INF = float("inf")


def case_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(f).startswith("lu_layer"))


class Case:
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.z = z
        self.dtype = getattr(torch, str(z["dtype"]))
        self.control = json.loads(str(z["control"]))
        self.has_A = bool(z["has_A"])
        self.iter = int(z["iter"])
        self.rho_is_tensor = bool(z["rho_is_tensor"])
        self.full_inputs = "Q" in z.files

    def t(self, key):
        return torch.from_numpy(self.z[key]) if key in self.z.files else None

    def inputs(self):
        if self.full_inputs:
            return (self.t("Q"), self.t("p"), self.t("A"), self.t("b"), self.t("lb"), self.t("ub"))
        # large cases: regenerate from the seed in the file name and pin by checksum
        parts = self.name.split("_")          # exp1_n500_b8_f64
        n, B = int(parts[1][1:]), int(parts[2][1:])
        seed = {"exp1_n500_b8_f64": 0, "exp1_n500_b8_f32": 0, "exp1_n250_b8_f64": 1, "exp1_n1000_b2_f64": 0}[self.name]
        Q, p, A, b, lb, ub = orc.make_exp1_data(n, B, seed=seed, dtype=self.dtype)
        chk = np.array([float(Q.double().sum()), float(p.double().sum()),
                        float(lb.double().sum()), float(ub.double().sum())])
        np.testing.assert_allclose(chk, self.z["input_checksum"], rtol=1e-9 if self.dtype == torch.float64 else 1e-6)
        return Q, p, A, b, lb, ub

    def control_dict(self):
        return dict(self.control)


def rel_err(a, b):
    """max-norm relative error ||a-b||_inf / max(||b||_inf, tiny), the yardstick of
    north_star ('within 1e-8 relative')."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    return float(np.abs(a - b).max() / den) if b.size else 0.0


def compare(case: Case, sol: dict, grads, tol: dict, skip=()):
    """sol: dict with x,z,u,lams,nus,rho,iter (torch CPU tensors / numbers);
    grads: (dQ, dp, dA, db, dlb, dub).  tol maps key -> relative tolerance
    ('default' applies to keys not listed)."""
    z = case.z
    errs = {}

    def chk(key, val):
        if key in skip or key not in z.files:
            return
        e = rel_err(val.detach().cpu().numpy() if torch.is_tensor(val) else val, z[key])
        errs[key] = e
        lim = tol.get(key, tol["default"])
        assert e <= lim, f"{case.name}: {key} rel err {e:.3e} > {lim:.1e}"

    for k in ("x", "z", "u", "lams", "nus"):
        if sol.get(k) is not None:
            chk(k, sol[k])
    if "iter" not in skip:
        assert abs(int(sol["iter"]) - case.iter) <= tol.get("iter", 0), \
            f"{case.name}: iter {sol['iter']} vs reference {case.iter}"
    if "rho" not in skip:
        assert torch.is_tensor(sol["rho"]) == case.rho_is_tensor, f"{case.name}: rho type"
        chk("rho", sol["rho"] if torch.is_tensor(sol["rho"]) else np.float64(sol["rho"]))
    if grads is not None:
        dQ, dp, dA, db, dlb, dub = grads
        for k, v in (("dp", dp), ("dA", dA), ("db", db), ("dlb", dlb), ("dub", dub)):
            if v is not None:
                chk(k, v)
        if dQ is not None:
            if "dQ" in z.files:
                chk("dQ", dQ)
            else:
                gen = torch.Generator().manual_seed(4321)
                w = torch.randn(dQ.shape[0], dQ.shape[1], 2, generator=gen, dtype=case.dtype)
                chk("dQ_probe", torch.matmul(dQ.cpu(), w))
                chk("dQ_fro", torch.linalg.matrix_norm(dQ.cpu()))
    return errs


# ---------------------------------------------------------------------------------- KKT backward fixtures
KKT_DIR = os.path.join(GOLDEN_DIR, "kkt")


def kkt_case_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(KKT_DIR, "*.npz")))


def compare_kkt(case: Case, grads, tol: dict):
    """grads = (dQ, dp, dA, db, dlb, dub) of the KKT backward evaluated at the REFERENCE's forward solution
    (case.t('x'), 'lams', 'nus'); compared with tests/golden/kkt/<case>.npz (make_golden_kkt.py)."""
    z = np.load(os.path.join(KKT_DIR, case.name + ".npz"), allow_pickle=False)
    assert not bool(z["reference_nan"])
    dQ, dp, dA, db, dlb, dub = grads
    assert (dlb is not None) == bool(z["has_dlb"]) and (dub is not None) == bool(z["has_dub"]), case.name
    errs = {}

    def chk(key, val):
        e = rel_err(val.detach().cpu().numpy(), z[key])
        errs[key] = e
        lim = tol.get(key, tol["default"])
        assert e <= lim, f"kkt/{case.name}: {key} rel err {e:.3e} > {lim:.1e}"

    for k, v in (("dp", dp), ("dA", dA), ("db", db), ("dlb", dlb), ("dub", dub)):
        if v is not None:
            chk(k, v)
    if "dQ" in z.files:
        chk("dQ", dQ)
    else:
        gen = torch.Generator().manual_seed(4321)
        w = torch.randn(dQ.shape[0], dQ.shape[1], 2, generator=gen, dtype=case.dtype)
        chk("dQ_probe", torch.matmul(dQ.cpu(), w))
        chk("dQ_fro", torch.linalg.matrix_norm(dQ.cpu()))
    return errs


def kkt_reference_is_nan(name):
    return bool(np.load(os.path.join(KKT_DIR, name + ".npz"), allow_pickle=False)["reference_nan"])


def kkt_reduced_fp64(dl_dz, x, lams, nus, Q, A, lb, ub):
    """The KKT adjoint with the 2n inequality rows eliminated in closed form (what csrc/backward.cu solves),
    evaluated in fp64 on the CPU.  Equal to the reference's dense (3n+m) solve wherever that is finite; used
    as the yardstick for the one-sided boxes where the reference returns NaN."""
    dl_dz, x, lams, Q, lb, ub = (t.double().cpu() for t in (dl_dz, x, lams, Q, lb, ub))
    B, n = Q.shape[0], Q.shape[1]
    lam = lams.clamp(min=1e-8)
    s_lo, s_hi = (x - lb).clamp(min=1e-8), (ub - x).clamp(min=1e-8)
    H = Q + torch.diag_embed((lam[:, :n] / s_lo + lam[:, n:] / s_hi).squeeze(2))
    if A is not None:
        A, nus = A.double().cpu(), nus.double().cpu()
        m = A.shape[1]
        K = torch.cat((torch.cat((H, A.transpose(1, 2)), 2), torch.cat((A, torch.zeros(B, m, m, dtype=torch.float64)), 2)), 1)
        rhs = torch.cat((-dl_dz, torch.zeros(B, m, 1, dtype=torch.float64)), 1)
    else:
        K, rhs = H, -dl_dz
    d = torch.linalg.solve(K, rhs)
    dx = d[:, :n]
    half = 0.5 * dx @ x.transpose(1, 2)
    dA = db = None
    if A is not None:
        dnu = d[:, n:]
        dA, db = dnu @ x.transpose(1, 2) + nus @ dx.transpose(1, 2), -dnu
    return half + half.transpose(1, 2), dx, dA, db, -lam[:, :n] * dx / s_lo, -lam[:, n:] * dx / s_hi


# ---------------------------------------------------------------------------------- unrolled mode fixtures
UNROLL_DIR = os.path.join(GOLDEN_DIR, "unroll")


def unroll_case_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(UNROLL_DIR, "*.npz")))


class UnrollCase:
    """tests/golden/unroll/<name>.npz (make_golden_unroll.py): reference run with control['unroll'] = True."""

    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(UNROLL_DIR, name + ".npz"), allow_pickle=False)
        self.z = z
        self.dtype = getattr(torch, str(z["dtype"]))
        self.control = json.loads(str(z["control"]))
        self.adaptive_update = name.startswith("adapt")     # an adaptive-rho refactorisation happens inside the loop

    def inputs(self):
        z = self.z
        if "Q" in z.files:
            return tuple(torch.from_numpy(z[k]) if k in z.files else None for k in ("Q", "p", "A", "b", "lb", "ub"))
        n, B = (int(v) for v in z["gen_shape"])
        Q, p, A, b, lb, ub = orc.make_exp1_data(n, B, seed=int(z["gen_seed"]), dtype=self.dtype)
        chk = np.array([float(Q.double().sum()), float(p.double().sum()), float(lb.double().sum()), float(ub.double().sum())])
        # fp32 generation (randn -> bmm) rounds differently from one host CPU to the next: pin the stream, not the bits
        np.testing.assert_allclose(chk, z["input_checksum"], rtol=1e-9 if self.dtype == torch.float64 else 1e-6)
        return Q, p, A, b, lb, ub

    def compare(self, x, grads, tol):
        """x and grads = (dQ, dp, dA, db, dlb, dub) as CPU tensors (None where autograd produced no gradient).
        A gradient the reference left at None (lb without a finite lower bound, ...) must be None or zero here;
        entries where the reference itself is NaN (0 * inf in the scaling of an infinite bound) are not compared."""
        z = self.z
        errs = {}

        def chk(key, val, ref):
            ref = np.asarray(ref)
            val = val.detach().cpu().numpy()
            ok = np.isfinite(ref)
            if not ok.all():
                val, ref = np.where(ok, val, 0.0), np.where(ok, ref, 0.0)
            e = rel_err(val, ref)
            errs[key] = e
            lim = tol.get(key, tol["default"])
            assert e <= lim, f"unroll/{self.name}: {key} rel err {e:.3e} > {lim:.1e}"

        chk("x", x, z["x"])
        for k, g in zip(("dQ", "dp", "dA", "db", "dlb", "dub"), grads):
            if k == "dQ" and "dQ" not in z.files:
                gen = torch.Generator().manual_seed(4321)
                w = torch.randn(g.shape[0], g.shape[1], 2, generator=gen, dtype=self.dtype)
                chk("dQ_probe", torch.matmul(g.cpu(), w), z["dQ_probe"])
                chk("dQT_probe", torch.matmul(g.cpu().transpose(1, 2), w), z["dQT_probe"])
                chk("dQ_fro", torch.linalg.matrix_norm(g.cpu()), z["dQ_fro"])
            elif k in z.files:
                assert g is not None, f"unroll/{self.name}: {k} missing"
                chk(k, g, z[k])
            else:
                assert g is None or float(g.abs().max()) == 0.0, f"unroll/{self.name}: {k} should be None"
        return errs

"""Parity margins of the unrolled mode on the GPU box: max-norm relative error of x and of every input gradient
against each reference fixture in tests/golden/unroll/ (table to stdout).   python tools/unroll_parity.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP  # noqa: E402
from tests._golden import UnrollCase, unroll_case_names, rel_err  # noqa: E402

dev = torch.device("cuda:0")
print(f"{'case':28s} " + " ".join(f"{k:>9s}" for k in ("x", "dQ", "dp", "dA", "db", "dlb", "dub")))
for name in unroll_case_names():
    case = UnrollCase(name)
    leaves = [None if t is None else t.to(dev).requires_grad_(True) for t in case.inputs()]
    x = SolveBoxQP(control=dict(case.control)).forward(*leaves)
    x.backward(torch.from_numpy(case.z["dl_dz"]).to(dev))
    z = case.z
    cols = [rel_err(x.detach().cpu().numpy(), z["x"])]
    for k, t in zip(("dQ", "dp", "dA", "db", "dlb", "dub"), leaves):
        if k == "dQ" and "dQ" not in z.files:
            gen = torch.Generator().manual_seed(4321)
            w = torch.randn(t.shape[0], t.shape[1], 2, generator=gen, dtype=case.dtype)
            cols.append(max(rel_err(torch.matmul(t.grad.cpu(), w).numpy(), z["dQ_probe"]),
                            rel_err(torch.matmul(t.grad.cpu().transpose(1, 2), w).numpy(), z["dQT_probe"])))
        elif k in z.files and t is not None and t.grad is not None:
            ref = np.asarray(z[k])
            val = t.grad.cpu().numpy()
            ok = np.isfinite(ref)
            cols.append(rel_err(np.where(ok, val, 0.0), np.where(ok, ref, 0.0)))
        else:
            cols.append(float("nan"))
    print(f"{name:28s} " + " ".join(f"{c:9.1e}" for c in cols), flush=True)

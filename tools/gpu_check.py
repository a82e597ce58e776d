"""Developer tool (GPU box): run every golden case through the CUDA path and print an error table.
Unlike the pytest suite it never stops at the first failure.  Usage: python tools/gpu_check.py [name-substring]"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests._golden import Case, case_names, rel_err  # noqa: E402
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad  # noqa: E402


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    dev = torch.device("cuda:0")
    for name in case_names():
        if pat not in name:
            continue
        case = Case(name)
        try:
            torch.set_default_dtype(case.dtype)
            ins = [None if t is None else t.to(dev) for t in case.inputs()]
            t0 = time.time()
            sol = torch_solve_box_qp(*ins, case.control_dict())
            g = torch_solve_box_qp_grad(case.t("dl_dz").to(dev), sol["x"], sol["u"], sol["lams"], sol["nus"],
                                        ins[0], ins[2], ins[4], ins[5], sol["rho"])
            torch.cuda.synchronize()
            dt = time.time() - t0
            z = case.z
            parts = [f"iter {sol['iter']}/{case.iter}"]
            for k in ("x", "z", "u", "lams", "nus"):
                if k in z.files and sol[k] is not None:
                    parts.append(f"{k} {rel_err(sol[k].cpu().numpy(), z[k]):.1e}")
            rho = sol["rho"]
            parts.append(f"rho {rel_err(rho.cpu().numpy() if torch.is_tensor(rho) else np.float64(rho), z['rho']):.1e}")
            for k, v in zip(("dQ", "dp", "dA", "db", "dlb", "dub"), g[:6]):
                if k in z.files and v is not None:
                    parts.append(f"{k} {rel_err(v.cpu().numpy(), z[k]):.1e}")
            if "dQ_probe" in z.files:
                gen = torch.Generator().manual_seed(4321)
                w = torch.randn(g[0].shape[0], g[0].shape[1], 2, generator=gen, dtype=case.dtype)
                parts.append(f"dQp {rel_err(torch.matmul(g[0].cpu(), w).numpy(), z['dQ_probe']):.1e}")
            print(f"{name:28s} {dt*1e3:7.1f}ms  " + "  ".join(parts), flush=True)
        except Exception:
            print(f"{name:28s} EXCEPTION\n{traceback.format_exc()}", flush=True)
        finally:
            torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    main()

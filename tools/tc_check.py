"""Developer tool (GPU box): the tensor-core factorisation (csrc/tcfactor.cu) in isolation and end to end.
  1. lqpb_dev_tc_inverse_f32 against torch.linalg.inv in fp64 (SPD and KKT-shaped quasi-definite matrices)
  2. forward + backward at dz=500, B=128, fp32: tensor-core path vs the Gauss-Jordan path (LQPB_FACTOR=gj), timings
Usage: python tools/tc_check.py [N] [B]"""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lqp_py_b200 import _abi  # noqa: E402
from lqp_py_b200.control import box_qp_control  # noqa: E402
from lqp_py_b200.datasets import create_qp_data  # noqa: E402
from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp, torch_solve_box_qp_grad  # noqa: E402


def tc_inverse(A):
    L = _abi.lib()
    B, N, _ = A.shape
    out = torch.empty_like(A)
    wb = L.lqpb_dev_tc_inverse_work_bytes(B, N)
    work = torch.empty(wb, dtype=torch.uint8, device=A.device)
    rc = L.lqpb_dev_tc_inverse_f32(B, N, _abi.ptr(A), _abi.ptr(out), _abi.ptr(work),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _abi.check(rc, "dev_tc_inverse")
    torch.cuda.synchronize()
    return out


def rel(a, r):
    return float((a.double() - r.double()).abs().max() / r.double().abs().max())


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for N_ in sorted({128, 256, N}):
        Lm = torch.randn(B, 2 * N_, N_, dtype=torch.float64, device=dev)
        A = Lm.transpose(1, 2) @ Lm / (2 * N_) + 1.2 * torch.eye(N_, dtype=torch.float64, device=dev)
        ref = torch.linalg.inv(A)
        got = tc_inverse(A.float().contiguous())
        ref32 = torch.linalg.inv(A.float())
        print(f"SPD N={N_}: tc vs fp64 inverse {rel(got, ref):.3e}   (torch fp32 inverse vs fp64: {rel(ref32, ref):.3e})")
        # KKT-shaped: last row/col = ones constraint with zero diagonal, identity padding handled by caller here
        K = A.clone()
        K[:, -1, :] = 1.0
        K[:, :, -1] = 1.0
        K[:, -1, -1] = 0.0
        ref = torch.linalg.inv(K)
        got = tc_inverse(K.float().contiguous())
        print(f"KKT N={N_}: tc vs fp64 inverse {rel(got, ref):.3e}   (torch fp32: {rel(torch.linalg.inv(K.float()), ref):.3e})")
    # timing of the bare inverse
    A32 = A.float().contiguous().repeat(max(1, 128 // B), 1, 1)[:128].contiguous()
    for _ in range(3):
        tc_inverse(A32)
    t0 = time.perf_counter()
    for _ in range(10):
        tc_inverse(A32)
    print(f"dev inverse N={A32.shape[1]} B={A32.shape[0]}: {(time.perf_counter() - t0) * 100:.3f} ms per call (incl. pack/unpack)")

    # end to end, both paths
    data = [t.to(dev) for t in create_qp_data(500, 128, 1000, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
    control = box_qp_control(eps_abs=1e-5, eps_rel=1e-5)
    gen = torch.Generator().manual_seed(3)
    g = torch.randn(128, 500, 1, generator=gen).to(dev)
    res = {}
    _abi.profile_enable(True)
    for mode in ("gj", "tc"):
        if mode == "gj":
            os.environ["LQPB_FACTOR"] = "gj"
        else:
            os.environ.pop("LQPB_FACTOR", None)
        for rep in range(3):
            sol = torch_solve_box_qp(*data, control)
            prf = _abi.profile_get()
            grads = torch_solve_box_qp_grad(g, sol["x"], sol["u"], sol["lams"], sol["nus"], data[0], data[2], data[4],
                                            data[5], sol["rho"])
            torch.cuda.synchronize()
            prb = _abi.profile_get()
        res[mode] = (sol, grads)
        print(f"{mode}: iter {sol['iter']}  factor {prf['factor_ms']:.3f} ms ({prf['factor_launches']} launches)  "
              f"iterate {prf['iterate_ms']:.3f} ms  bwd_factor {prb['bwd_factor_ms']:.3f} ms  bwd_grad {prb['bwd_grad_ms']:.3f} ms")
    (s0, g0), (s1, g1) = res["gj"], res["tc"]
    for k in ("x", "z", "u", "lams", "nus"):
        print(f"  {k:5s} tc vs gj: {rel(s1[k], s0[k]):.3e}")
    for name, a, r in zip(("dQ", "dp", "dA", "db", "dlb", "dub"), g1, g0):
        print(f"  {name:5s} tc vs gj: {rel(a, r):.3e}")


if __name__ == "__main__":
    main()

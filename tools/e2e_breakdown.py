"""Developer tool (GPU box): where the end-to-end (host tensors in, host gradients out) step time goes.
Usage: python tools/e2e_breakdown.py [f32|f64] [dz] [B]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lqp_py_b200.control import box_qp_control  # noqa: E402
from lqp_py_b200.datasets import create_qp_data  # noqa: E402
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP  # noqa: E402


def t_ms(fn, reps=5):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


def main():
    dt = torch.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else torch.float64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    dev = torch.device("cuda:0")
    host = [t.pin_memory() for t in create_qp_data(n, B, 2 * n, seed=0, requires_grad=False, dtype=dt)[:6]]
    Qh = host[0]
    Qd = Qh.to(dev)
    out_pin = torch.empty_like(Qh).pin_memory()
    mb = Qh.numel() * Qh.element_size() / 1e6
    h2d = t_ms(lambda: Qd.copy_(Qh, non_blocking=True))
    d2h = t_ms(lambda: out_pin.copy_(Qd, non_blocking=True))
    print(f"raw pinned H2D {mb:.0f} MB: {h2d:.2f} ms = {mb / h2d:.1f} GB/s ; D2H {d2h:.2f} ms = {mb / d2h:.1f} GB/s")
    pag = torch.empty_like(Qh)
    print(f"pageable D2H: {t_ms(lambda: pag.copy_(Qd)):.2f} ms ; new pinned alloc+D2H: "
          f"{t_ms(lambda: torch.empty(Qh.shape, dtype=dt, pin_memory=True).copy_(Qd, non_blocking=True)):.2f} ms")
    control = box_qp_control(eps_rel=1e-5, eps_abs=1e-5)
    QP = SolveBoxQP(control=control)
    g_h = torch.ones(B, n, 1, dtype=dt).pin_memory()
    g_d = g_h.to(dev)
    devt = [t.to(dev) for t in host]

    def step(ins_src, g):
        ins = [t.detach().requires_grad_(k < 2) for k, t in enumerate(ins_src)]
        t0 = time.perf_counter()
        x = QP.forward(*ins)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        x.backward(g)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3

    for name, src, g in (("device tensors", devt, g_d), ("pinned host tensors", host, g_h)):
        for _ in range(3):
            step(src, g)
        f = b = 0.0
        R = 10
        for _ in range(R):
            a, c = step(src, g)
            f += a; b += c
        print(f"{name:22s}: forward {f / R:.2f} ms  backward {b / R:.2f} ms  total {(f + b) / R:.2f} ms")


if __name__ == "__main__":
    main()

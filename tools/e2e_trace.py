"""Host-side timeline of the end-to-end loop (pinned CPU tensors in, x / dQ / dp back on the host) in its three forms:
unannounced, next batch announced (SolveBoxQP.prefetch), two batches announced and solved ahead (SolveBoxQP.solve_ahead).
Prints, per form, the mean wall time of forward / announce / backward and of the whole step."""
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lqp_py_b200.control import box_qp_control                     # noqa: E402
from lqp_py_b200.datasets import create_qp_data                    # noqa: E402
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP         # noqa: E402
from lqp_py_b200 import solve_box_qp_admm_torch as M               # noqa: E402


def main():
    n, B, K = int(os.environ.get("DZ", 500)), int(os.environ.get("BATCH", 128)), int(os.environ.get("STEPS", 20))
    torch.cuda.set_device(0)
    sets = []
    for k in range(3):
        d = create_qp_data(n, B, 2 * n, seed=k, requires_grad=False, dtype=torch.float32)
        sets.append([t.pin_memory() for t in d[:6]])
    g = torch.ones(B, n, 1).pin_memory()
    QP = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5, verbose=False, reduce='max'))
    for mode in (0, 1, 2):
        if mode == 2:
            QP.solve_ahead(*sets[0])
            QP.solve_ahead(*sets[1])
        acc = [0.0, 0.0, 0.0]
        t_all = 0.0
        for k in range(K + 6):
            if k == 6:
                torch.cuda.synchronize()
                acc = [0.0, 0.0, 0.0]
                t_all = time.perf_counter()
            ins = [None if t is None else t.detach().requires_grad_(j < 2) for j, t in enumerate(sets[k % 3])]
            t0 = time.perf_counter()
            x = QP.forward(*ins)
            t1 = time.perf_counter()
            if mode == 1:
                QP.prefetch(*sets[(k + 1) % 3])
            elif mode == 2:
                QP.solve_ahead(*sets[(k + 2) % 3])
            t2 = time.perf_counter()
            x.backward(g)
            t3 = time.perf_counter()
            acc[0] += t1 - t0
            acc[1] += t2 - t1
            acc[2] += t3 - t2
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t_all
        for pf in list(M._PREFETCH.values()):
            if pf.get("future") is not None:
                pf["future"].result()
        M._PREFETCH.clear()
        print(f"mode {mode} chunks={os.environ.get('LQPB_HOST_CHUNKS', 'default')}: step {t_all / K * 1e3:.3f} ms  forward {acc[0] / K * 1e3:.3f}  "
              f"announce {acc[1] / K * 1e3:.3f}  backward {acc[2] / K * 1e3:.3f}  ({B / (t_all / K):.0f} QP/s)", flush=True)


if __name__ == "__main__":
    main()

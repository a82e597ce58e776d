"""Where an epoch of the Experiment-2 loop (lqp_py_b200/sharding.py:train_learn_p, mini-batch 32, dz = 500) goes: CUDA-event
times of its segments and the host wall time, averaged over the epochs after warm-up."""
import os
import sys
import time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lqp_py_b200.control import box_qp_control                     # noqa: E402
from lqp_py_b200.datasets import create_qp_data                    # noqa: E402
from lqp_py_b200.sharding import qp_cost                           # noqa: E402
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP         # noqa: E402
from lqp_py_b200 import _abi                                        # noqa: E402


def main():
    dz, nB, nf, mini = 500, 128, 5, 32
    dev = torch.device("cuda", 0)
    Q, _, A, b, lb, ub = [t.to(dev) for t in create_qp_data(dz, nB, 2 * dz, seed=0, requires_grad=False, dtype=torch.float32)[:6]]
    gen = torch.Generator().manual_seed(0)
    feats = torch.randn(nB, nf, generator=gen).to(dev)
    p_true = (feats @ torch.randn(nf, dz, generator=gen).to(dev)).unsqueeze(2)
    QP = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5))
    torch.manual_seed(0)
    model = torch.nn.Linear(nf, dz).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=5e-4)
    rng = np.random.RandomState(0)
    names = ["gather+linear", "qp forward", "loss", "backward", "step+float(loss)"]
    acc = np.zeros(len(names))
    wall = 0.0
    iters = []
    E, W = 60, 10
    for epoch in range(E + W):
        idx = torch.as_tensor(rng.randint(low=0, high=nB, size=mini), dtype=torch.long, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        t0 = time.perf_counter()
        ev[0].record()
        opt.zero_grad()
        p_hat = model(feats[idx]).unsqueeze(2)
        Qi = Q[idx]
        args = (Qi, p_hat, A[idx], b[idx], lb[idx], ub[idx])
        ev[1].record()
        z = QP(*args)
        ev[2].record()
        loss = qp_cost(z, Qi, p_true[idx])
        ev[3].record()
        loss.backward()
        ev[4].record()
        opt.step()
        lv = float(loss)
        ev[5].record()
        torch.cuda.synchronize()
        if epoch >= W:
            wall += time.perf_counter() - t0
            acc += [ev[k].elapsed_time(ev[k + 1]) for k in range(len(names))]
    print(f"epoch wall {wall / E * 1e3:.3f} ms; device segments (ms): " +
          ", ".join(f"{n} {a / E:.3f}" for n, a in zip(names, acc)) + f"; sum {acc.sum() / E:.3f}")


if __name__ == "__main__":
    main()

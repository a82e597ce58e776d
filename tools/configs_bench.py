"""Forward+backward timing of every BASELINE.json config on one GPU (CUDA events, device-resident inputs):
Experiment 1 at dz = 10 / 100 / 250 / 500 / 1000 with batch 128 (fixed-point and KKT backward), and Experiment 2
(learning p, dz=500, mini-batch 32, Linear(5, 500), SGD) as ms per epoch.  One JSON line per config.
    python tools/configs_bench.py [--dtype f32] [--steps 10]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lqp_py_b200 import sharding
from lqp_py_b200.control import box_qp_control
from lqp_py_b200.datasets import create_qp_data
from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="f32")
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dt = torch.float32 if a.dtype == "f32" else torch.float64
dev = torch.device("cuda:0")
torch.set_default_dtype(dt)
B = 128
for dz in (10, 100, 250, 500, 1000):
    sets = [[t.to(dev) for t in create_qp_data(dz, B, 2 * dz, seed=s, requires_grad=False, dtype=dt)[:6]] for s in range(3)]
    g = torch.ones(B, dz, 1, dtype=dt, device=dev)
    for backward in ("fixed_point", "kkt"):
        QP = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5, backward=backward))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        it = None
        for k in range(3 + a.steps):
            ins = [t.detach().requires_grad_(True) for t in sets[k % 3]]
            ev[0].record()
            x = QP.forward(*ins)
            ev[1].record()
            x.backward(g)
            ev[2].record()
            torch.cuda.synchronize()
            if k >= 3:
                tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
        print(json.dumps({"config": f"Experiment 1 dz={dz} B={B} tol=1e-5 backward={backward}", "dtype": a.dtype,
                          "forward_ms": tf / a.steps, "backward_ms": tb / a.steps,
                          "qp_per_s": B * a.steps / ((tf + tb) * 1e-3)}), flush=True)
# Experiment 2: learning p (experiments/experiment_2.py:52-99)
dz, nB, nf = 500, 128, 5
Q, _, A, b, lb, ub = [t.to(dev) for t in create_qp_data(dz, nB, 2 * dz, seed=0, requires_grad=False, dtype=dt)[:6]]
gen = torch.Generator().manual_seed(0)
feats = torch.randn(nB, nf, generator=gen, dtype=dt).to(dev)
p_true = (feats @ torch.randn(nf, dz, generator=gen, dtype=dt).to(dev)).unsqueeze(2)
QP = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5))
sharding.train_learn_p(QP, Q, p_true, A, b, lb, ub, feats, n_epochs=5, n_mini_batch=32, lr=5e-4, seed=0)
torch.cuda.synchronize()
t0 = time.perf_counter()
model, hist = sharding.train_learn_p(QP, Q, p_true, A, b, lb, ub, feats, n_epochs=100, n_mini_batch=32, lr=5e-4, seed=0)
torch.cuda.synchronize()
dtm = time.perf_counter() - t0
print(json.dumps({"config": "Experiment 2 learning p dz=500 mini-batch 32, 100 epochs, Linear(5,500), SGD lr 5e-4",
                  "dtype": a.dtype, "ms_per_epoch": dtm * 10, "qp_per_s": 32 * 100 / dtm, "loss_first": hist[0],
                  "loss_last": hist[-1]}), flush=True)

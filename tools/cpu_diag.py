"""Diagnose the CPU baseline on a GPU box: run the oracle port in fresh subprocesses with different thread
settings and report time / finiteness / MKL complaints.  python tools/cpu_diag.py"""
import os, subprocess, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, os, json
sys.path.insert(0, %r)
import torch
T = int(sys.argv[1]); B = int(sys.argv[2]); dt = torch.float32 if sys.argv[3] == "f32" else torch.float64
cuda_first = sys.argv[4] == "1"
if cuda_first and torch.cuda.is_available():
    torch.zeros(1, device="cuda")
if T > 0:
    torch.set_num_threads(T)
torch.set_default_dtype(dt)
from oracle import box_qp_oracle as orc
from lqp_py_b200.datasets import create_qp_data
Q, p, A, b, lb, ub, _, _ = create_qp_data(500, B, 1000, seed=0, requires_grad=False, dtype=dt)
control = orc.default_control(eps_abs=1e-5, eps_rel=1e-5)
g = torch.ones(B, 500, 1, dtype=dt)
t0 = time.perf_counter()
sol, grads = orc.solve_and_grad(Q, p, A, b, lb, ub, control, g)
t1 = time.perf_counter()
print(json.dumps(dict(threads=torch.get_num_threads(), B=B, dtype=sys.argv[3], cuda_first=cuda_first, s=t1 - t0, iter=sol["iter"],
                      finite=bool(torch.isfinite(sol["x"]).all()), qfinite=bool(torch.isfinite(Q).all()))))
''' % ROOT
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
print(subprocess.run("lscpu | head -20", shell=True, capture_output=True, text=True).stdout, flush=True)
for T, B, dt, cf in [(0, 8, "f32", 0), (16, 8, "f32", 0), (16, 8, "f32", 1), (1, 8, "f32", 0), (16, 32, "f32", 0), (16, 128, "f32", 0), (16, 128, "f64", 0), (0, 128, "f32", 1)]:
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, str(T), str(B), dt, str(cf)], capture_output=True, text=True, timeout=60)
        err = r.stderr.count("oneMKL ERROR")
        print(T, B, dt, cf, "->", r.stdout.strip()[-300:], "| rc", r.returncode, "mkl_errors", err, "| wall %.1f" % (time.time() - t0), flush=True)
        if err:
            print("   first stderr:", r.stderr.strip().splitlines()[:2], flush=True)
    except subprocess.TimeoutExpired as e:
        se = (e.stderr or b"")
        se = se.decode() if isinstance(se, bytes) else se
        print(T, B, dt, cf, "-> TIMEOUT 60 s, mkl_errors", se.count("oneMKL ERROR"), flush=True)

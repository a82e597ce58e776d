#!/usr/bin/env python
"""Headline benchmark: QPs/sec forward+backward of the ADMM box-QP layer, Experiment 1 of the reference
(experiments/experiment_1.py: dz=500, batch 128, tol 1e-5, default control = scale, rho=None, adaptive_rho),
and the HBM roofline of the ADMM iteration kernel.

    python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port of the reference path

A "step" = one forward solve + one fixed-point backward of a whole batch (B problems per GPU, fresh
synthetic data each step, rotating over datasets larger than L2).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "QPs/sec fwd+bwd at dz=500,B=128,tol=1e-5"
UNIT = "QP/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dz", type=int, default=500)
    ap.add_argument("--batch", type=int, default=128, help="problems per GPU (weak scaling)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="f32 = the reference experiments' dtype")
    ap.add_argument("--datasets", type=int, default=3, help="distinct seeded input sets rotated over the steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"Experiment 1 box QP dz={a.dz}, batch {a.batch} per GPU, tol 1e-5, scale=True, rho=None, "
            f"adaptive_rho=True, ADMM fixed-point forward+backward, {a.dtype}")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampler (B200_PROFILING.md clocks line) running during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def log(msg):
    """progress marker on stderr (stdout carries only the JSON line)"""
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def usable_cpus():
    """Host threads this process can really use: the affinity mask, capped by the cgroup CPU quota (a box can
    show 200 CPUs and grant 16; sizing the BLAS pool by os.cpu_count() there oversubscribes by 10x)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
    except Exception:
        try:                                        # cgroup v1
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0 and period > 0:
                n = min(n, max(1, int(quota / period + 0.5)))
        except Exception:
            pass
    return max(1, n)


def gen_data(a, seed, dtype):
    from lqp_py_b200.datasets import create_qp_data
    Q, p, A, b, lb, ub, _, _ = create_qp_data(a.dz, a.batch, 2 * a.dz, seed=seed, requires_grad=False, dtype=dtype)
    return [Q, p, A, b, lb, ub]


# ------------------------------------------------------------------------------------------------
def run_reference(a):
    """CPU arm: the oracle port of the reference's torch path (oracle/box_qp_oracle.py -- same batched
    LAPACK calls through torch.linalg as the reference), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import box_qp_oracle as orc
    # torchrun exports OMP_NUM_THREADS=1; this arm is the CPU path with every host thread it can use
    torch.set_num_threads(usable_cpus())
    dtype = torch.float32 if a.dtype == "f32" else torch.float64
    K = a.steps if a.steps is not None else 3
    W = a.warmup if a.warmup is not None else 1
    torch.set_default_dtype(dtype)
    data = gen_data(a, 0, dtype)
    control = orc.default_control(eps_abs=1e-5, eps_rel=1e-5)
    g = torch.ones(a.batch, a.dz, 1, dtype=dtype)            # experiment_1.py:75
    log(f"reference arm: {torch.get_num_threads()} threads, {W} warm-up + {K} steps of {a.batch} problems")
    for _ in range(W):
        orc.solve_and_grad(*data, control, g)
    t0 = time.perf_counter()
    iters = None
    for k in range(K):
        sol, _ = orc.solve_and_grad(*data, control, g)
        iters = sol["iter"]
        log(f"reference arm: step {k + 1}/{K} at {time.perf_counter() - t0:.1f} s")
    dt = time.perf_counter() - t0
    val = a.batch * K / dt
    cores = torch.get_num_threads()
    sample = f"{K} steps of the full batch ({a.batch} problems, dz={a.dz}), {W} warm-up, ADMM iter={iters}"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 0, "launched_as_gpus": a.gpus,
           "steps": K, "warmup": W,
           "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": a.dtype, "data": "synthetic", "config": {"workload": workload_name(a)},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                            "host_cpus": os.cpu_count()},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def cpu_baseline(a, dtype):
    from oracle import box_qp_oracle as orc
    torch.set_num_threads(usable_cpus())
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        data = gen_data(a, 0, dtype)
        control = orc.default_control(eps_abs=1e-5, eps_rel=1e-5)
        g = torch.ones(a.batch, a.dz, 1, dtype=dtype)
        # bounded sample: time 8 problems first, then as much of the batch as fits in ~10 s per pass
        nb0 = min(8, a.batch)
        small = [t[:nb0] for t in data]
        orc.solve_and_grad(*small, control, g[:nb0])           # warm up LAPACK / thread pool
        t0 = time.perf_counter()
        orc.solve_and_grad(*small, control, g[:nb0])
        per_problem = (time.perf_counter() - t0) / nb0
        nb = int(max(nb0, min(a.batch, 10.0 / max(per_problem, 1e-9))))
        log(f"cpu baseline: {torch.get_num_threads()} threads, {per_problem * 1e3:.1f} ms per problem on a batch of "
            f"{nb0} -> sample of {nb} problems per pass")
        sample = [t[:nb] for t in data]
        t0 = time.perf_counter()
        reps = 0
        iters = None
        while reps < 1 or (time.perf_counter() - t0 < 8.0 and reps < 6):
            sol, _ = orc.solve_and_grad(*sample, control, g[:nb])
            iters = sol["iter"]
            reps += 1
        dt = time.perf_counter() - t0
    finally:
        torch.set_default_dtype(prev)
    return {"value": nb * reps / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": f"{reps} forward+backward passes of {nb} of the {a.batch} problems (dz={a.dz}, {a.dtype}) "
                      f"with the oracle port (torch CPU, batched LAPACK), ADMM iter={iters}, {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch.distributed as dist
    from lqp_py_b200 import _abi
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if a.dtype == "f32" else torch.float64
    K = a.steps if a.steps is not None else 20
    W = a.warmup if a.warmup is not None else 3
    W = max(W, 3)
    n, B, s = a.dz, a.batch, (4 if a.dtype == "f32" else 8)

    # independent problems: every rank owns its own shard of the global batch, no collective in the solve
    host_sets = [gen_data(a, 1000 * rank + k, dtype) for k in range(a.datasets)]
    dev_sets = [[t.to(dev) for t in d] for d in host_sets]
    g_dev = torch.ones(B, n, 1, dtype=dtype, device=dev)                  # experiment_1.py:75
    control = box_qp_control(eps_rel=1e-5, eps_abs=1e-5, verbose=False, reduce='max')   # experiment_1.py:22
    QP = SolveBoxQP(control=control)
    _abi.profile_enable(True)

    prof_acc = {}
    launches = [0]
    iters_seen = []

    def add_prof(keys):
        pr = _abi.profile_get()
        for k in keys:
            prof_acc[k] = prof_acc.get(k, 0.0) + pr[k]
        return pr

    FWD = ("scale_ms", "factor_ms", "iterate_ms", "finalize_ms")
    BWD = ("bwd_factor_ms", "bwd_solve_ms", "bwd_grad_ms")

    have_bwd = [False]

    def step(k, record):
        ins = [t.detach().requires_grad_(True) for t in dev_sets[k % len(dev_sets)]]
        x = QP.forward(*ins)                       # syncs once at the end of the solve (reads `iter`)
        if record:
            # forward phases of this step; the previous step's backward events are complete too
            pr = add_prof(FWD + (BWD if have_bwd[0] else ()))
            launches[0] += pr["kernel_launches"]
        x.backward(g_dev)                          # asynchronous
        if record:
            launches[0] += 4
            have_bwd[0] = True
        return ins

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    log(f"rank {rank}: data on {dev}, {W} warm-up + {K} timed steps")
    for k in range(W):
        step(k, False)
    sync_all()
    log("warm-up done")
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        step(W + k, True)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    add_prof(BWD)            # the last step's backward
    log(f"timed region done: {ms / K:.3f} ms per step")
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * K / (ms * 1e-3)

    # ---- roofline of the iteration kernel (dominant): algorithmic bytes / CUDA-event time of the launches
    sol = None
    from lqp_py_b200.solve_box_qp_admm_torch import torch_solve_box_qp
    sol = torch_solve_box_qp(*dev_sets[0], control)
    it = sol["iter"]
    passes = it + 1
    check = max(round((n ** 0.5) / 10) * 10, 1)
    checks = it // check + 1
    N = n + 1
    bytes_iter = B * s * (N * N + 7 * n)
    bytes_launch = passes * bytes_iter + checks * B * s * n * n   # + Q~ x~ at the checks
    it_ms = prof_acc["iterate_ms"] / K
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = bytes_launch / (it_ms * 1e-3) / 1e9
    roofline = {"kernel": "iterate_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                "traffic": None, "bytes_per_launch": bytes_launch, "ms_per_launch": it_ms,
                "admm_passes": passes, "checks": checks, "us_per_admm_iteration": it_ms * 1e3 / passes}
    tr = os.path.join(ROOT, "profiles", "iterate_traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(f"{a.dtype}_dz{n}_B{B}")
        except Exception:
            pass
    phases = {k: prof_acc[k] / K for k in FWD + BWD}

    # ---- e2e: the same step through the public module API with HOST (pinned) tensors
    e2e = None
    if not a.no_e2e:
        pin_sets = [[t.pin_memory() for t in d] for d in host_sets]
        g_host = torch.ones(B, n, 1, dtype=dtype).pin_memory()
        def step_host(k):
            # leaves as experiments/utils.py:41-50 creates them: Q and p require grad, A, b, lb, ub do not
            ins = [t.detach().requires_grad_(j < 2) for j, t in enumerate(pin_sets[k % len(pin_sets)])]
            x = QP.forward(*ins)
            x.backward(g_host)
            return x, ins
        # warm-up until torch's caching pinned-host allocator holds every staging block a step needs
        # (a fresh cudaHostAlloc of a 128 MB gradient block costs tens of ms and is not steady state)
        for k in range(max(W, 5)):
            x, ins = step_host(k)
        sync_all()
        Ke = max(3, min(K, 10))
        t0 = time.perf_counter()
        for k in range(Ke):
            x, ins = step_host(5 + k)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = sum(t.numel() for t in pin_sets[0]) * s + g_host.numel() * s
        d2h = (x.numel() + sum(t.grad.numel() for t in ins if t.grad is not None)) * s
        log(f"e2e done: {dt / Ke * 1e3:.3f} ms per step")
        e2e = {"value": B * world * Ke / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": Ke, "ms_per_step": dt / Ke * 1e3,
               "how": "SolveBoxQP.forward + x.backward on pinned CPU tensors (Q, p require grad as in "
                      "experiments/utils.py:41-50; x, dQ, dp come back to the host); copies in the timed region"}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a, dtype)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": a.dtype, "data": "synthetic",
               "config": {"workload": workload_name(a), "global_batch": B * world, "dz": n, "n_eq": 1,
                          "admm_iter": it, "parallelism": f"batch-sharded x{world}, no collective in the solve",
                          "l2": f"{a.datasets} rotating input sets + workspace = {a.datasets * B * n * n * s / 1e6:.0f} MB "
                                f"+ {2 * B * n * n * s / 1e6:.0f} MB per step > 126 MB L2"},
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches[0],
               "phases_ms": phases, "clocks": clocks}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)

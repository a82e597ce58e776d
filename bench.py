#!/usr/bin/env python
"""Headline benchmark: QPs/sec forward+backward of the ADMM box-QP layer, Experiment 1 of the reference
(experiments/experiment_1.py: dz=500, batch 128, tol 1e-5, default control = scale, rho=None, adaptive_rho),
and the roofline of the ADMM iteration kernel.

    python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the UNMODIFIED reference (oracle/_ref) on the host cores

A "step" = one forward solve + one fixed-point backward of a whole batch (B problems per GPU, fresh synthetic data
each step, rotating over datasets larger than L2).  Prints ONE JSON line (rank 0).  Besides the contract's keys the
line carries measured bandwidth / tensor ceilings (`peaks`), a second roofline object from an HBM-bound configuration
(`roofline_hbm`), the other BASELINE.json configs (`configs`), the batch sweep of config 5 (`sweep`) and the
Experiment-2 learning loop with its NCCL all-reduce (`exp2`).
"""
import argparse
import hashlib
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
L2_BYTES = 126e6
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-child", "reference-cuda"])
    ap.add_argument("--threads", type=int, default=1, help="(reference-child) torch / BLAS threads")
    ap.add_argument("--nprob", type=int, default=8, help="(reference-child) problems per pass")
    ap.add_argument("--kind", default="reference", choices=["reference", "port"], help="(reference-child) which CPU code")
    ap.add_argument("--dz", type=int, default=500)
    ap.add_argument("--batch", type=int, default=128, help="problems per GPU (weak scaling)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="f32 = the reference experiments' dtype")
    ap.add_argument("--datasets", type=int, default=3, help="distinct seeded input sets rotated over the steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip peaks / roofline_hbm / configs / sweep / exp2")
    return ap.parse_args()


def workload_name(a):
    return (f"Experiment 1 box QP dz={a.dz}, batch {a.batch} per GPU, tol 1e-5, scale=True, rho=None, "
            f"adaptive_rho=True, ADMM fixed-point forward+backward, {a.dtype}")


def config_dict(a, world, it):
    """The `config` object -- built by ONE function for both arms, so that the driver's same_config check compares
    like with like."""
    s = 4 if a.dtype == "f32" else 8
    n, B = a.dz, a.batch
    return {"workload": workload_name(a), "global_batch": B * world, "dz": n, "n_eq": 1, "admm_iter": it,
            "parallelism": f"batch-sharded x{world}, no collective in the solve",
            "l2": f"{a.datasets} rotating input sets + workspace = {a.datasets * B * n * n * s / 1e6:.0f} MB "
                  f"+ {2 * B * n * n * s / 1e6:.0f} MB per step > 126 MB L2"}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region.  Primary source: NVML polled every
    4 ms from a thread of this process (the main thread sits in ctypes calls that release the GIL); fallback: the
    nvidia-smi query line of B200_PROFILING.md on a 25 ms loop."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []        # (host time, sm MHz, max MHz, set of reasons)
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.source = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except Exception:
                    phys = self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.source = "nvml"
            self.th = threading.Thread(target=self._poll_nvml, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi"
            self.th = threading.Thread(target=self._read_smi, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nvml
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((time.perf_counter(), mhz, self.max_mhz, {n for n, b in names if bits & b}))
            except Exception:
                pass
            time.sleep(0.004)

    def _read_smi(self):
        for ln in self.proc.stdout:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.samples.append((time.perf_counter(), sm, mx, {n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")}))

    @property
    def alive(self):
        return self.nvml is not None or self.proc is not None

    def count(self, t_lo):
        return sum(1 for s in self.samples if s[0] >= t_lo)

    def stop(self, t_lo=0.0, t_hi=float("inf"), t_timed_hi=None):
        """Statistics over the samples taken in [t_lo, t_hi] (host clock): the timed region plus, when that region is
        too short to hold 8 samples, the identical steps run right after it."""
        if not self.alive:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML, nvidia-smi) available"]}
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sel = [s for s in self.samples if t_lo <= s[0] <= t_hi]
        sm = sorted(s[1] for s in sel)
        reasons = set()
        for s in sel:
            reasons |= s[3]
        in_timed = sum(1 for s in sel if t_timed_hi is not None and s[0] <= t_timed_hi)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((s[2] for s in sel), default=None),
                "samples": len(sm), "samples_in_timed_region": in_timed, "reasons": sorted(reasons), "source": self.source}


def log(msg):
    """progress marker on stderr (stdout carries only the JSON line); rank 0 only under torchrun"""
    if int(os.environ.get("RANK", "0") or 0) != 0:
        return
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


def gen_data(dz, batch, seed, dtype):
    """experiments/utils.py:41-61 (bit-identical restatement in lqp_py_b200/datasets.py), on the CPU."""
    from lqp_py_b200.datasets import create_qp_data
    Q, p, A, b, lb, ub, _, _ = create_qp_data(dz, batch, 2 * dz, seed=seed, requires_grad=False, dtype=dtype)
    return [Q, p, A, b, lb, ub]


def gen_data_device(dz, batch, seed, dtype, dev):
    """The same recipe drawn on the device (large batches of the sweep: the CPU generator would take minutes)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    L = torch.randn(batch, 2 * dz, dz, generator=g, dtype=dtype, device=dev)
    Q = torch.matmul(L.transpose(1, 2), L) / (2 * dz)
    del L
    p = torch.randn(batch, dz, 1, generator=g, dtype=dtype, device=dev)
    A = torch.ones(batch, 1, dz, dtype=dtype, device=dev)
    b = torch.ones(batch, 1, 1, dtype=dtype, device=dev)
    lb = -(torch.rand(batch, dz, 1, generator=g, dtype=dtype, device=dev) + 1)
    ub = torch.rand(batch, dz, 1, generator=g, dtype=dtype, device=dev) + 1
    return [Q, p, A, b, lb, ub]


def have_reference():
    return os.path.exists(os.path.join(REF_DIR, "lqp_py", "solve_box_qp_admm_torch.py"))


# ------------------------------------------------------------------------------------------------
def reference_child(a):
    """One bounded run of the reference's CPU path in THIS process (spawned by cpu_arm with the thread count in the
    environment): W warm-up + K timed forward+backward passes over the first --nprob problems of the workload.
    --kind reference: the UNMODIFIED reference package from oracle/_ref (SolveBoxQP + autograd backward, exactly
    experiments/experiment_1.py:22-23,70-77); --kind port: the oracle's restatement of it."""
    dtype = torch.float32 if a.dtype == "f32" else torch.float64
    # the pool size comes from OMP_NUM_THREADS (set by the parent).  torch.set_num_threads(k > 1) is NOT used: with
    # this torch / oneMKL build it makes the threaded sgetrf fail ("Parameter 6 was incorrect on entry to SLASWP")
    # and spin for minutes, on every host tried.
    torch.set_default_dtype(dtype)
    data = [t[:a.nprob].contiguous() for t in gen_data(a.dz, a.batch, 0, dtype)]
    g = torch.ones(a.nprob, a.dz, 1, dtype=dtype)            # experiment_1.py:75
    if a.kind == "reference":
        sys.path.insert(0, REF_DIR)
        from lqp_py.control import box_qp_control as ref_control
        from lqp_py.solve_box_qp_admm_torch import SolveBoxQP as RefSolveBoxQP
        from lqp_py.solve_box_qp_admm_torch import torch_solve_box_qp as ref_solve
        control = ref_control(eps_rel=1e-5, eps_abs=1e-5, verbose=False, reduce='max')     # experiment_1.py:22
        QP = RefSolveBoxQP(control=control)

        def one():
            ins = [t.detach().clone().requires_grad_(j < 2) for j, t in enumerate(data)]   # experiments/utils.py:41-50
            x = QP.forward(Q=ins[0], p=ins[1], A=ins[2], b=ins[3], lb=ins[4], ub=ins[5])
            x.backward(g)
            return x.detach(), [ins[0].grad, ins[1].grad]
        it = int(ref_solve(*data, control)["iter"])          # untimed: the iteration count this sample needs
    else:
        from oracle import box_qp_oracle as orc
        control = orc.default_control(eps_abs=1e-5, eps_rel=1e-5)
        state = {}

        def one():
            sol, grads = orc.solve_and_grad(*data, control, g)
            state["iter"] = sol["iter"]
            return sol["x"], [t for t in grads if t is not None]
        one()
        it = int(state["iter"])
    for _ in range(a.warmup or 0):
        one()
    t0 = time.perf_counter()
    x = grads = None
    for _ in range(a.steps or 1):
        x, grads = one()
    dt = time.perf_counter() - t0
    ok = bool(torch.isfinite(x).all()) and all(bool(torch.isfinite(t).all()) for t in grads if t is not None)
    print(json.dumps({"seconds": dt, "iter": it, "finite": ok, "threads": torch.get_num_threads(), "kind": a.kind}),
          flush=True)


def _spawn_child(a, kind, threads, nprob, steps, warmup, timeout):
    """Run reference_child in a fresh interpreter with a hard wall-clock limit.  Returns the child's dict or None
    (time-out, crash, non-finite results or LAPACK complaints on stderr)."""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads)                     # torchrun exports OMP_NUM_THREADS=1
    env.pop("MKL_NUM_THREADS", None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-child", "--kind", kind, "--threads", str(threads),
           "--nprob", str(nprob), "--steps", str(steps), "--warmup", str(warmup), "--dz", str(a.dz),
           "--batch", str(a.batch), "--dtype", a.dtype]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    except subprocess.TimeoutExpired:
        log(f"cpu arm: {threads} threads x {nprob} problems did not finish in {timeout} s")
        return None
    bad = r.stderr.count("MKL ERROR")
    try:
        out = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        out = None
    if r.returncode != 0 or out is None or not out.get("finite") or bad:
        log(f"cpu arm ({kind}): {threads} threads x {nprob} problems failed (rc {r.returncode}, {bad} LAPACK errors): "
            f"{r.stderr[-300:]}")
        return None
    return out


def cpu_arm(a, steps, warmup, budget_s):
    """The reference's CPU path on the host cores, bounded: every run happens in a subprocess with a time limit, the
    thread count is the fastest of {all usable CPUs, half, 1} on a 16-problem probe (batched getrs does not scale with
    threads and an over-subscribed BLAS pool crawls), and a step covers as many problems of the batch as fit the time
    budget.  The code timed is the UNMODIFIED reference package (oracle/_ref, pip-installed from /root/reference by
    __graft_entry__.build(): kind "reference"); if it is absent or fails here, the oracle port (kind "port")."""
    ncpu = usable_cpus()
    cands = sorted({ncpu, max(1, ncpu // 2), 1}, reverse=True)
    nb0 = min(16, a.batch)
    kinds = (["reference"] if have_reference() else []) + ["port"]
    for kind in kinds:
        best = None
        for t in cands:
            out = _spawn_child(a, kind, t, nb0, 1, 1, 90)
            if out is None:
                continue
            log(f"cpu arm probe ({kind}): {t} threads, {nb0} problems: {out['seconds']:.2f} s")
            if best is None or out["seconds"] < best[1]:
                best = (t, out["seconds"])
        if best is None:
            continue
        threads, probe_s = best
        per_problem = probe_s / nb0
        # problems per step so that the whole K + W run fits the budget (never fewer than 4, never more than the batch)
        nb = int(max(min(4, a.batch), min(a.batch, budget_s / max(per_problem * (steps + warmup), 1e-9))))
        out = None
        for attempt in range(3):
            out = _spawn_child(a, kind, threads, nb, steps, warmup, 60 + 4 * budget_s)
            if out is not None:
                break
            nb = max(min(4, a.batch), nb // 4)        # shrink the sample (and, last resort, the thread pool) and retry
            if attempt == 1:
                threads = 1
        if out is None:
            continue
        val = nb * steps / out["seconds"]
        what = ("the unmodified reference package (oracle/_ref: lqp_py.SolveBoxQP.forward + x.backward, torch CPU)"
                if kind == "reference" else "the oracle port (torch CPU, batched LAPACK)")
        sample = (f"{steps} forward+backward passes ({warmup} warm-up) over {nb} of the {a.batch} problems (dz={a.dz}, "
                  f"{a.dtype}) with {what}, ADMM iter={out['iter']}, "
                  f"{out['seconds']:.1f} s, {threads} threads chosen from {cands} by a {nb0}-problem probe")
        return {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                "host_cpus": os.cpu_count(), "usable_cpus": ncpu, "seconds": out["seconds"], "problems_per_step": nb,
                "iter": out["iter"]}
    return None


def run_reference(a):
    """CPU arm: the reference's own torch path on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K = a.steps if a.steps is not None else 3
    W = a.warmup if a.warmup is not None else 1
    cpu = cpu_arm(a, K, W, budget_s=60.0)      # the K + W passes share ~60 s of CPU work: a step is a bounded sample
    if cpu is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither the reference package nor the oracle port "
                          "completed on this host (time-out or LAPACK failure in every configuration)"}), flush=True)
        return
    val = cpu["value"]
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 0, "launched_as_gpus": a.gpus,
           "steps": K, "warmup": W,
           "ms_per_step": cpu["seconds"] / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": a.dtype, "data": "synthetic", "config": config_dict(a, a.gpus, cpu["iter"]),
           "cpu_baseline": cpu,
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_reference_cuda(a, quiet=False):
    """Secondary baseline: the UNMODIFIED reference package (oracle/_ref) with its tensors on cuda:0, i.e.
    torch's library path on the same B200 (batched cuSOLVER / cuBLAS LU factor + 2 triangular solves per ADMM iteration,
    torch.linalg.solve in the backward) -- `torch.set_default_device` makes the reference's own factory calls
    (solve_box_qp_admm_torch.py:206-223) allocate on the GPU; nothing of this repo's library is on that path.  Rank 0 only;
    prints one JSON line with "impl": "reference-cuda" (returns it when quiet)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return None
    if not (have_reference() and torch.cuda.is_available()):
        out = {"impl": "reference-cuda", "unavailable": "needs oracle/_ref (built by __graft_entry__.build() where "
               "/root/reference exists) and a CUDA device"}
        if not quiet:
            print(json.dumps(out), flush=True)
        return out
    K = a.steps if a.steps is not None else 5
    W = a.warmup if a.warmup is not None else 3
    dtype = torch.float32 if a.dtype == "f32" else torch.float64
    data_cpu = gen_data(a.dz, a.batch, 0, dtype)
    prev_dev, prev_dt = torch.get_default_device(), torch.get_default_dtype()
    try:
        torch.set_default_device("cuda:0")
        torch.set_default_dtype(dtype)
        sys.path.insert(0, REF_DIR)
        from lqp_py.control import box_qp_control as ref_control
        from lqp_py.solve_box_qp_admm_torch import torch_solve_box_qp as ref_solve
        control = ref_control(eps_rel=1e-5, eps_abs=1e-5, verbose=False, reduce='max')     # experiment_1.py:22
        data = [t.to("cuda:0") for t in data_cpu]
        g = torch.ones(a.batch, a.dz, 1, dtype=dtype)                                      # experiment_1.py:75

        from lqp_py.solve_box_qp_admm_torch import torch_solve_box_qp_grad as ref_grad

        def one():
            # forward = SolveBoxQPLayer.forward's call (:39), backward = SolveBoxQPLayer.backward's call (:63), both made
            # from THIS thread: torch's default device is thread-local, and the reference's backward allocates with bare
            # factory calls (:363), so inside autograd's own thread it would land on the CPU
            sol = ref_solve(*data, control)
            grads = ref_grad(g, x=sol["x"], u=sol["u"], lams=sol["lams"], nus=sol["nus"], Q=data[0], A=data[2], lb=data[4],
                             ub=data[5], rho=sol["rho"])
            return sol["x"], grads[0]
        it = int(ref_solve(*data, control)["iter"])
        for _ in range(W):
            one()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(K):
            x, dQ = one()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / K
        ok = bool(torch.isfinite(x).all()) and bool(torch.isfinite(dQ).all())
    finally:
        torch.set_default_device(prev_dev)
        torch.set_default_dtype(prev_dt)
    val = a.batch / (ms * 1e-3)
    out = {"impl": "reference-cuda", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype,
           "data": "synthetic", "config": config_dict(a, 1, it), "finite": ok,
           "how": "unmodified reference package (oracle/_ref) under torch.set_default_device('cuda:0'): torch's library "
                  "kernels (batched LU factor / solve) on the same GPU, torch_solve_box_qp + torch_solve_box_qp_grad called as the layer calls them; CUDA events around K forward+backward passes"}
    if not quiet:
        print(json.dumps(out), flush=True)
    return out


def cpu_baseline(a):
    return cpu_arm(a, steps=2, warmup=1, budget_s=15.0)


# ------------------------------------------------------------------------------------------------
FWD = ("scale_ms", "factor_ms", "iterate_ms", "finalize_ms")
BWD = ("bwd_factor_ms", "bwd_solve_ms", "bwd_grad_ms")


class Ctx:
    """Process-wide state of the B200 arm."""
    def __init__(self, a):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != a.gpus and self.world == 1 and a.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def sync_all(self):
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v):
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def time_layer(cx, dev_sets, B, n, dtype, K, W, backward="fixed_point", sampler=None, prof_every=4):
    """W warm-up + K timed forward+backward steps of SolveBoxQP on device-resident inputs (rotating over dev_sets),
    bracketed by barrier + synchronize, CUDA events on the launch stream, max over ranks.  Returns a dict with the step
    time, per-phase device times (read back from the library's events on every `prof_every`-th step only), the kernel
    launch count and the ADMM iteration count."""
    from lqp_py_b200 import _abi
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP, torch_solve_box_qp
    dev = cx.dev
    g_dev = torch.ones(B, n, 1, dtype=dtype, device=dev)                  # experiment_1.py:75
    control = box_qp_control(eps_rel=1e-5, eps_abs=1e-5, verbose=False, reduce='max', backward=backward)   # experiment_1.py:22
    QP = SolveBoxQP(control=control)
    _abi.profile_enable(True)
    acc, n_prof = {}, [0, 0]

    def add_prof(keys):
        pr = _abi.profile_get()
        for k in keys:
            acc[k] = acc.get(k, 0.0) + pr[k]
        return pr

    def step(k, read_fwd):
        ins = [t.detach().requires_grad_(True) for t in dev_sets[k % len(dev_sets)]]
        x = QP.forward(*ins)                       # the library waits for the solve alone (it reports `iter`)
        if read_fwd:
            add_prof(FWD)
            n_prof[0] += 1
        x.backward(g_dev)                          # asynchronous
        return ins

    for k in range(W):
        step(k, False)
    cx.sync_all()
    # kernel launches of one forward / one backward call (the library counts them; constant for a given shape and
    # iteration count): read once here, multiplied by the timed steps below
    ins0 = [t.detach().requires_grad_(True) for t in dev_sets[0]]
    x0 = QP.forward(*ins0)
    fwd_l = _abi.profile_get()["kernel_launches"]
    x0.backward(g_dev)
    torch.cuda.synchronize(dev)
    bwd_l = _abi.profile_get()["kernel_launches"]
    del ins0, x0
    cx.sync_all()
    t_host0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # the forward phases are read back (a blocking event query) on a few steps only: on the small configurations the
    # layer's forward is asynchronous and every read-back would serialise host and device for that step
    prof_every = max(prof_every, K // 8)
    for k in range(K):
        step(W + k, k % prof_every == 0)
    e1.record()
    cx.sync_all()
    ms = e0.elapsed_time(e1)
    t_host1 = time.perf_counter()
    add_prof(BWD)            # the last step's backward
    n_prof[1] += 1
    clocks = None
    if sampler is not None:
        # keep the GPU under the SAME load (the same steps, untimed) until at least 8 clock samples exist
        k, t_end = 0, time.perf_counter() + 3.0
        while sampler.alive and sampler.count(t_host0) < 8 and time.perf_counter() < t_end:
            step(W + K + k, False)
            k += 1
        torch.cuda.synchronize(dev)
        try:
            clocks = sampler.stop(t_host0, time.perf_counter(), t_host1)
        except Exception as exc:
            clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"clock sampler failed: {exc!r}"]}
        clocks["extra_load_steps"] = k
    for _ in range(3):       # a few more backward-phase samples, outside the timed region
        step(0, False)
        torch.cuda.synchronize(dev)
        add_prof(BWD)
        n_prof[1] += 1
    ms = cx.max_over_ranks(ms)
    sol = torch_solve_box_qp(*dev_sets[0], control)
    _abi.profile_enable(False)
    phases = {k: acc.get(k, 0.0) / max(n_prof[0] if k in FWD else n_prof[1], 1) for k in FWD + BWD}
    phases["steps_sampled"] = {"forward": n_prof[0], "backward": n_prof[1]}
    return {"ms_per_step": ms / K, "value": B * cx.world * K / (ms * 1e-3), "iter": int(sol["iter"]), "phases_ms": phases,
            "launches_per_step": fwd_l + bwd_l, "clocks": clocks, "steps": K}


def source_digest():
    """Digest of the sources that define the iteration kernel: an ncu traffic figure is only quoted for the build it
    was captured from."""
    h = hashlib.sha256()
    for f in ("iterate.cu", "itergeom.cuh", "layout.cuh", "common.cuh"):
        with open(os.path.join(ROOT, "lqp_py_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


REGIMES = {0: "stream", 1: "packed-resident", 2: "dense-rows", 3: "fused dense-rows (one-launch forward)",
           4: "stream, every problem split over a cluster of 2 / 4 CTAs (small batches)"}


def iterate_regime(n, B, dtype_name):
    """Which iteration kernel this shape takes on this device (lqpb_iterate_regime_*)."""
    try:
        import ctypes as C
        from lqp_py_b200 import _abi
        from lqp_py_b200.control import box_qp_control
        from lqp_py_b200.solve_box_qp_admm_torch import _derive_config
        cfg = _derive_config(box_qp_control(eps_rel=1e-5, eps_abs=1e-5), n)
        return int(getattr(_abi.lib(), f"lqpb_iterate_regime_{dtype_name}")(C.byref(cfg), B, n, 1))
    except Exception:
        return None


def iterate_roofline(n, B, s, it, it_ms, peaks, hbm_peak, dtype_name):
    """Roofline object of one iterate_kernel launch.  `achieved` / `frac` follow SURVEY 8(d): ALGORITHMIC bytes (one
    pass over the full N x N operator per ADMM iteration + vectors, + n^2 per stop check) / CUDA-event time.  The kernel
    streams the PACKED lower triangle (symmetry halves the traffic), so `frac` can exceed 1; `achieved_moved` /
    `frac_moved` count the bytes the kernel really requests and are the physical fraction of the bound's ceiling.  The
    bound is "l2" when the packed operator set of the batch fits the 126 MB L2 (it is then re-read from L2 every
    iteration and the ceiling is the measured L2 streaming-read rate), otherwise "hbm"."""
    passes = it + 1
    check = max(round((n ** 0.5) / 10) * 10, 1)
    checks = it // check + 1
    N = n + 1
    nt = (n + 31) // 32
    tc = 32 if s == 4 else 16
    r = 32 // tc
    nbc = nt * r
    # tiles of the packed lower triangle (layout.cuh Pack<T>): block column Jc holds block rows Jc / R .. nt - 1
    ntiles = sum(nt - jc // r for jc in range(nbc))
    packed = ntiles * 4096
    alg_iter = B * s * (N * N + 7 * n)
    alg_launch = passes * alg_iter + checks * B * s * n * n          # + Q~ x~ at the checks
    moved_iter = B * (packed + 7 * n * s)
    moved_launch = passes * moved_iter + checks * B * packed
    resident = B * packed <= 0.75 * L2_BYTES
    bound = "l2" if resident else "hbm"
    l2_peak = peaks.get("l2_read_gbs") if peaks else None
    peak = (l2_peak if resident else hbm_peak) or hbm_peak
    no_l2_peak = resident and not l2_peak        # --no-extras: the L2 ceiling was not measured in this run
    sec = it_ms * 1e-3
    achieved = alg_launch / sec / 1e9
    moved = moved_launch / sec / 1e9
    regime = iterate_regime(n, B, dtype_name)
    if regime in (1, 2, 3):
        # the operators live in shared memory for the whole solve: HBM / L2 are touched once, the loop is bound by
        # instruction latency and synchronisation -- no bandwidth ceiling applies, none is quoted
        kern = {1: "iterate_res_kernel", 2: "iterate_row_kernel<FUSED=false>", 3: "iterate_row_kernel<FUSED=true>"}[regime]
        note = ("operators resident in shared memory: latency / synchronisation bound, no bandwidth ceiling applies; "
                "`achieved` is the SURVEY 8(d) algorithmic figure for comparison only")
        if regime == 3:
            note += " (the launch also contains scaling, factorisation and finalisation)"
        return {"kernel": kern, "regime": REGIMES[regime], "bound": "smem-resident (latency)", "achieved": achieved,
                "peak": None, "unit": "GB/s", "frac": None, "traffic": None, "bytes_per_launch": alg_launch,
                "ms_per_launch": it_ms, "admm_passes": passes, "checks": checks,
                "us_per_admm_iteration": it_ms * 1e3 / passes, "note": note}
    out = {"kernel": "iterate_split_kernel" if regime == 4 else "iterate_kernel", "regime": REGIMES.get(regime),
           "bound": bound, "achieved": achieved,
           "peak": None if no_l2_peak else peak, "unit": "GB/s",
           "frac": None if no_l2_peak else achieved / peak, "achieved_moved": moved,
           "frac_moved": None if no_l2_peak else moved / peak,
           "peak_source": ("measured in this run (lqpb_dev_stream_read on an L2-resident buffer)" if resident and l2_peak
                           else "MEASURED_PEAKS.json hbm_gbs" if hbm_peak_measured() else "fallback 6650 GB/s"),
           "frac_vs_hbm_copy_peak": achieved / hbm_peak, "hbm_peak": hbm_peak,
           "traffic": None, "bytes_per_launch": alg_launch, "bytes_moved_per_launch": moved_launch,
           "operator_set_bytes": B * packed, "ms_per_launch": it_ms, "admm_passes": passes, "checks": checks,
           "us_per_admm_iteration": it_ms * 1e3 / passes,
           "note": "frac = SURVEY 8(d) algorithmic bytes (full N^2 operator) / time / peak; the kernel moves the packed "
                   "lower triangle (about half), frac_moved is the physical fraction of the ceiling"}
    tr = os.path.join(ROOT, "profiles", "iterate_traffic.json")
    try:
        tj = json.load(open(tr))
        key = f"{dtype_name}_dz{n}_B{B}"
        if tj.get("digest") == source_digest():
            out["traffic"] = tj.get("entries", {}).get(key)
            out["traffic_source"] = tj.get("source")
        else:
            out["traffic_source"] = ("stale: profiles/iterate_traffic.json was captured from another build of the "
                                     "iteration kernel; not quoted")
    except Exception:
        pass
    return out


def hbm_peak_measured():
    try:
        return "hbm_gbs" in json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return False


def measure_peaks(cx):
    """Ceilings measured in THIS run on this GPU: streaming read of an L2-resident buffer (48 MB, 40 passes) and of a
    2 GB buffer (HBM), both through lqpb_dev_stream_read; cuBLAS TF32 and FP64 GEMM rates (the factorisation's
    ceilings: MEASURED_PEAKS.json only records bf16)."""
    import ctypes as C
    from lqp_py_b200 import _abi
    L = _abi.lib()
    dev = cx.dev
    out = {}
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream

    def stream_read(nbytes, reps):
        buf = torch.empty(nbytes // 4, dtype=torch.int32, device=dev).random_(0, 1 << 20)
        for _ in range(2):
            _abi.check(L.lqpb_dev_stream_read(_abi.ptr(buf), nbytes, reps, _abi.ptr(sink), C.c_void_p(st)), "stream_read")
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _abi.check(L.lqpb_dev_stream_read(_abi.ptr(buf), nbytes, reps, _abi.ptr(sink), C.c_void_p(st)), "stream_read")
            e1.record()
            torch.cuda.synchronize(dev)
            best = max(best, nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        return best

    out["l2_read_gbs"] = stream_read(48 << 20, 40)
    out["hbm_read_gbs"] = stream_read(2 << 30, 2)

    def gemm(dtype, n, tf32):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            a = torch.randn(n, n, dtype=dtype, device=dev)
            b = torch.randn(n, n, dtype=dtype, device=dev)
            for _ in range(2):
                torch.matmul(a, b)
            best = 0.0
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize(dev)
                best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
            return best
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    out["tf32_tflops"] = gemm(torch.float32, 8192, True)
    out["fp64_tflops"] = gemm(torch.float64, 4096, False)
    out["how"] = ("l2_read / hbm_read: lqpb_dev_stream_read (16-byte ld.global.cg, 4 in flight per thread, 592 CTAs) over "
                  "48 MB x 40 passes / 2 GB x 2 passes, best of 5, CUDA events; tf32 / fp64: torch.matmul (cuBLAS) "
                  "8192^3 with allow_tf32 / 4096^3 fp64, best of 5")
    return out


def factor_roofline(n, B, phases, peaks):
    """Tensor-pipe roofline of the fp32 factorisation (tcgen05 kind::tf32, 3 products per tile product: hi*hi + lo*hi +
    hi*lo): algorithmic flops of the block sweeps of one forward + one backward / their CUDA-event time, against the
    cuBLAS TF32 rate measured in this run."""
    N = n + 1
    if N <= 128 or not peaks or not peaks.get("tf32_tflops"):
        return None
    nb = (N + 127) // 128
    tile = 2.0 * 128 ** 3
    fwd_prod = nb * ((nb - 1) + (nb - 1) * nb / 2)                       # PANEL + TRAIL tile products, inverse mode
    bwd_prod = sum((nb - 1 - k) + (nb - 1 - k) * (nb - k) / 2 for k in range(nb))   # LDL^T mode
    piv = 2 * nb * 2.0 * 128 ** 3                                        # pivot-block inverses (FP32 pipe), fwd + bwd
    flops_tc = 3 * tile * B * (fwd_prod + bwd_prod)
    sec = (phases["factor_ms"] + phases["bwd_factor_ms"]) * 1e-3
    ach = flops_tc / sec / 1e12
    return {"kernel": "tc_tile_kernel<PANEL|TRAIL> + tc_pivot8_kernel (forward inverse + backward LDL^T)", "bound": "tensor",
            "achieved": ach, "peak": peaks["tf32_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tf32_tflops"],
            "peak_source": "measured in this run (cuBLAS TF32 GEMM 8192^3)", "flops_per_step_tensor": flops_tc,
            "flops_per_step_fp32_pipe_pivots": piv * B, "ms_per_step": sec * 1e3,
            "note": "tensor flops count the 3 TF32 products of every fp32-accurate tile product; the time includes the "
                    "FP32-pipe pivot-block inverses and the assembly / extract kernels"}


# ------------------------------------------------------------------------------------------------
def run_e2e(cx, a, host_sets, dtype, K, W):
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    dev, world = cx.dev, cx.world
    n, B, s = a.dz, a.batch, (4 if a.dtype == "f32" else 8)
    QP = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5, verbose=False, reduce='max'))
    pin_sets = [[t.pin_memory() for t in d] for d in host_sets]
    g_host = torch.ones(B, n, 1, dtype=dtype).pin_memory()

    def step_host(k, announce):
        # leaves as experiments/utils.py:41-50 creates them: Q and p require grad, A, b, lb, ub do not
        ins = [t.detach().requires_grad_(j < 2) for j, t in enumerate(pin_sets[k % len(pin_sets)])]
        x = QP.forward(*ins)
        if announce == 1:
            # a data loader that knows its next batch: the upload of step k + 1 (copy stream, H2D) runs while step k's
            # backward computes and streams dQ down (D2H) -- PCIe is full duplex.  Every step's H2D and D2H are inside
            # the timed region either way.
            QP.prefetch(*pin_sets[(k + 1) % len(pin_sets)])
        elif announce == 2:
            # ... and one that knows two: batch k + 2 is announced for upload AND solve (SolveBoxQP.solve_ahead); this
            # step's backward starts its copy, a worker thread queues its forward behind the copy on a second stream.
            # Steady state: H2D engine on batch k + 2, SMs on batch k + 1, D2H engine on batch k's gradients.
            QP.solve_ahead(*pin_sets[(k + 2) % len(pin_sets)])
        x.backward(g_host)
        return x, ins

    def timed(announce):
        # warm-up until torch's caching pinned-host allocator holds every staging block a step needs
        # (a fresh cudaHostAlloc of a 128 MB gradient block costs tens of ms and is not steady state)
        if announce == 2:
            QP.solve_ahead(*pin_sets[0])
            QP.solve_ahead(*pin_sets[1 % len(pin_sets)])
        for k in range(max(W, 5)):
            x, ins = step_host(k, announce)
        cx.sync_all()
        # timed on the device like `value`: events on the compute stream bracket the Ke steps (every step ends with the
        # copy stream joined back into it and the host buffers valid), max over ranks below
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        h0.record()
        for k in range(Ke):
            x, ins = step_host(max(W, 5) + k, announce)
        h1.record()
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        if announce == 2:
            # the two batches announced beyond the last step were uploaded and solved inside the timed region for nothing:
            # wait for them and drop them
            from lqp_py_b200 import solve_box_qp_admm_torch as _m
            for pf in list(_m._PREFETCH.values()):
                if pf.get("future") is not None:
                    pf["future"].result()
            _m._PREFETCH.clear()
            torch.cuda.synchronize(dev)
        if world > 1:
            cx.dist.barrier()
        return cx.max_over_ranks(h0.elapsed_time(h1) * 1e-3), wall, x, ins
    Ke = max(3, min(K, 20))
    dt_serial, _, x, ins = timed(0)
    dt, wall, x, ins = timed(1)
    # what a caller of the reference actually holds: ordinary (pageable) CPU tensors -- same loop, no announcement
    # (the library uploads them through page-locked staging memory, solve_box_qp_admm_torch._copy_up)
    page_sets = [[t.clone() for t in d] for d in host_sets[:2]]
    g_page = g_host.clone()

    def step_pageable(k):
        ins = [t.detach().requires_grad_(j < 2) for j, t in enumerate(page_sets[k % len(page_sets)])]
        x = QP.forward(*ins)
        x.backward(g_page)
    for k in range(4):
        step_pageable(k)
    cx.sync_all()
    tp = time.perf_counter()
    for k in range(max(3, Ke // 2)):
        step_pageable(k)
    torch.cuda.synchronize(dev)
    dt_page = (time.perf_counter() - tp) / max(3, Ke // 2)
    if world > 1:
        cx.dist.barrier()
    dt_page = cx.max_over_ranks(dt_page)
    del page_sets
    while len(pin_sets) < 3:       # batches k, k + 1, k + 2 are in flight at once: three distinct host buffers
        pin_sets.append([t.clone().pin_memory() for t in pin_sets[0]])
    dt_ahead, wall_ahead, x, ins = timed(2)
    h2d = sum(t.numel() for t in pin_sets[0]) * s + g_host.numel() * s
    grads = [t.grad for t in ins if t.grad is not None]
    d2h = (x.numel() + sum(t.numel() for t in grads)) * s
    # copy ceiling: the same bytes over the same pinned buffers with NO kernels (upload of the inputs, then download
    # of x and the gradients), timed the same way -- what PCIe / the host memory system allow for this step
    dev_in = [torch.empty_like(t, device=dev) for t in pin_sets[0]]
    dev_g = torch.empty_like(g_host, device=dev)
    dev_out = [torch.empty_like(t, device=dev) for t in [x] + grads]
    host_out = [torch.empty_like(t).pin_memory() for t in [x] + grads]

    def copy_step(k):
        for d, h in zip(dev_in, pin_sets[k % len(pin_sets)]):
            d.copy_(h, non_blocking=True)
        dev_g.copy_(g_host, non_blocking=True)
        for h, d in zip(host_out, dev_out):
            h.copy_(d, non_blocking=True)
    for k in range(3):
        copy_step(k)
    cx.sync_all()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for k in range(Ke):
        copy_step(k)
    c1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        cx.dist.barrier()
    ct = cx.max_over_ranks(c0.elapsed_time(c1) * 1e-3)
    # ... and the full-duplex ceiling: uploads on one stream, downloads on another
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def copy_duplex(k):
        with torch.cuda.stream(up):
            for d, h in zip(dev_in, pin_sets[k % len(pin_sets)]):
                d.copy_(h, non_blocking=True)
            dev_g.copy_(g_host, non_blocking=True)
        with torch.cuda.stream(down):
            for h, d in zip(host_out, dev_out):
                h.copy_(d, non_blocking=True)
    cx.sync_all()
    t1 = time.perf_counter()
    for k in range(Ke):
        copy_duplex(k)
    torch.cuda.synchronize(dev)
    cd = time.perf_counter() - t1
    if world > 1:
        cx.dist.barrier()
    cd = cx.max_over_ranks(cd)
    log(f"e2e done: {dt_ahead / Ke * 1e3:.3f} ms per step with two batches announced and solved ahead, "
        f"{dt / Ke * 1e3:.3f} ms with the next batch announced, {dt_serial / Ke * 1e3:.3f} ms without "
        f"(copy ceilings: {ct / Ke * 1e3:.3f} ms serial, {cd / Ke * 1e3:.3f} ms duplex)")
    # the headline is the loop a user would run: with the next batch announced where that pays (it does until the host
    # memory system is the limit -- at 8 ranks per host both forms sit on the same ceiling), else without; both are reported
    best, mode, wall = min((dt_ahead, "two batches announced and solved ahead (SolveBoxQP.solve_ahead)", wall_ahead),
                           (dt, "next batch announced (SolveBoxQP.prefetch)", wall),
                           (dt_serial, "unannounced", wall), key=lambda t: t[0])
    return {"value": B * world * Ke / best, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": Ke, "ms_per_step": best / Ke * 1e3, "host_wall_ms_per_step": wall / Ke * 1e3,
            "mode": mode,
            "with_solve_ahead": {"value": B * world * Ke / dt_ahead, "unit": UNIT, "ms_per_step": dt_ahead / Ke * 1e3,
                                 "host_wall_ms_per_step": wall_ahead / Ke * 1e3,
                                 "how": "SolveBoxQP.solve_ahead(batch k + 2) between forward(k) and backward(k): upload and "
                                        "forward solve of later batches overlap the gradient download of this one"},
            "with_prefetch": {"value": B * world * Ke / dt, "unit": UNIT, "ms_per_step": dt / Ke * 1e3},
            "pageable_inputs": {"value": B * world / dt_page, "unit": UNIT, "ms_per_step": dt_page * 1e3,
                                "how": "the unannounced loop on ordinary (pageable) CPU tensors, host wall clock: the "
                                       "library stages them through page-locked memory chunk by chunk "
                                       "(cudaMemcpyAsync straight from pageable memory: 16.5 ms per step)"},
            "copy_ceiling": {"ms_per_step": ct / Ke * 1e3, "value": B * world * Ke / ct, "unit": UNIT,
                             "gbs_per_gpu": (h2d + d2h) / (ct / Ke) / 1e9,
                             "how": "the step's H2D + D2H copies alone (same pinned buffers, same stream order, no "
                                    "kernels), max over ranks"},
            "copy_ceiling_duplex": {"ms_per_step": cd / Ke * 1e3, "value": B * world * Ke / cd, "unit": UNIT,
                                    "how": "the same copies with uploads and downloads on two streams (PCIe is full "
                                           "duplex); host wall clock around Ke steps, max over ranks"},
            "frac_of_copy_ceiling": (cd / Ke) / (best / Ke),
            "without_prefetch": {"value": B * world * Ke / dt_serial, "unit": UNIT, "ms_per_step": dt_serial / Ke * 1e3,
                                 "frac_of_serial_copy_ceiling": (ct / Ke) / (dt_serial / Ke),
                                 "how": "the same loop without SolveBoxQP.prefetch: upload, solve, backward and download "
                                        "of a step strictly one after the other"},
            "how": "SolveBoxQP.forward + x.backward on pinned CPU tensors (Q, p require grad as in "
                   "experiments/utils.py:41-50; x, dQ, dp come back to the host); measured three times -- with "
                   "batch k + 2 announced for upload and solve (SolveBoxQP.solve_ahead), with the next step's inputs "
                   "announced for upload only (SolveBoxQP.prefetch: overlaps this step's gradient download) and without "
                   "any announcement -- `value` is the fastest (`mode`); every step's H2D and D2H copies and every "
                   "step's kernels are inside the timed region in all three"}


def run_configs(cx, a, peaks, hbm_peak):
    """The other BASELINE.json configs on one GPU (device-resident, forward + backward, B = 128): Experiment 1 at
    dz = 10 / 100 / 250 / 1000 in fp32 (fixed-point and KKT backward) and dz = 500 in fp64 (the 1e-8 parity mode)."""
    out = []
    B = 128
    plan = [(10, "f32", 60), (100, "f32", 60), (250, "f32", 60), (1000, "f32", 12), (500, "f64", 20)]
    for dz, dn, K in plan:
        if dz == a.dz and dn == a.dtype and B == a.batch:
            continue
        dt = torch.float32 if dn == "f32" else torch.float64
        s = 4 if dn == "f32" else 8
        try:
            if dz <= 500:
                sets = [[t.to(cx.dev) for t in gen_data(dz, B, 100 + k, dt)] for k in range(3)]
            else:              # the CPU generator needs ~10 s per dz=1000 set: drawn on the device instead
                sets = [gen_data_device(dz, B, 100 + k, dt, cx.dev) for k in range(2)]
            for backward in ("fixed_point", "kkt"):
                r = time_layer(cx, sets, B, dz, dt, K, 5, backward=backward)
                rec = {"config": f"Experiment 1 dz={dz}, batch {B}, tol 1e-5, {dn}, backward={backward}", "dz": dz,
                       "dtype": dn, "backward": backward, "value": r["value"], "unit": UNIT,
                       "ms_per_step": r["ms_per_step"], "steps": K, "admm_iter": r["iter"], "phases_ms": r["phases_ms"],
                       "gpu_launches_per_step": r["launches_per_step"]}
                if backward == "fixed_point":
                    rec["roofline"] = iterate_roofline(dz, B, s, r["iter"], r["phases_ms"]["iterate_ms"], peaks, hbm_peak, dn)
                out.append(rec)
                log(f"config dz={dz} {dn} {backward}: {r['value']:.0f} QP/s ({r['ms_per_step']:.3f} ms/step, iter {r['iter']})")
            del sets
            torch.cuda.empty_cache()
        except Exception as exc:
            out.append({"config": f"Experiment 1 dz={dz} {dn}", "error": repr(exc)[:300]})
    return out


def run_sweep(cx, a, dtype):
    """BASELINE config 5, second half: batch-sharded sweep, B per GPU in {256, 512, 1024} at dz=500 on the N GPUs of this
    run (B = 128 per GPU is the headline `value`).  Inputs are drawn on the device."""
    out = []
    for Bg in (256, 512, 1024):
        try:
            sets = [gen_data_device(a.dz, Bg, 7000 + 10 * cx.rank + k, dtype, cx.dev) for k in range(2)]
            r = time_layer(cx, sets, Bg, a.dz, dtype, 6, 3)
            out.append({"batch_per_gpu": Bg, "global_batch": Bg * cx.world, "n_gpus": cx.world, "value": r["value"],
                        "unit": UNIT, "ms_per_step": r["ms_per_step"], "admm_iter": r["iter"],
                        "iterate_ms": r["phases_ms"]["iterate_ms"]})
            log(f"sweep B/GPU={Bg}: {r['value']:.0f} QP/s")
            del sets
            torch.cuda.empty_cache()
        except Exception as exc:
            out.append({"batch_per_gpu": Bg, "error": repr(exc)[:300]})
    return out


def run_exp2(cx, a, dtype):
    """BASELINE config 5, first half: Experiment 2 (experiments/experiment_2.py:52-99) -- learn p through the layer,
    dz=500, 128 stored QPs, mini-batch 32 PER GPU (drawn with replacement), Linear(5, 500), SGD lr 5e-4, 100 epochs;
    with N > 1 the ranks solve disjoint shards of each global mini-batch and the Linear gradients are summed by ONE
    NCCL all-reduce per epoch (lqp_py_b200/sharding.py)."""
    from lqp_py_b200 import sharding
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    dz, nB, nf, dev = 500, 128, 5, cx.dev
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        Q, _, A, b, lb, ub = [t.to(dev) for t in gen_data(dz, nB, 0, dtype)]
        gen = torch.Generator().manual_seed(0)
        feats = torch.randn(nB, nf, generator=gen, dtype=dtype).to(dev)
        p_true = (feats @ torch.randn(nf, dz, generator=gen, dtype=dtype).to(dev)).unsqueeze(2)
        QP = SolveBoxQP(control=box_qp_control(eps_rel=1e-5, eps_abs=1e-5))
        mini = 32 * cx.world
        sharding.train_learn_p(QP, Q, p_true, A, b, lb, ub, feats, n_epochs=5, n_mini_batch=mini, lr=5e-4, seed=0)
        cx.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model, hist = sharding.train_learn_p(QP, Q, p_true, A, b, lb, ub, feats, n_epochs=100, n_mini_batch=mini, lr=5e-4,
                                             seed=0)
        e1.record()
        cx.sync_all()
        ms = cx.max_over_ranks(e0.elapsed_time(e1))
    finally:
        torch.set_default_dtype(prev)
    return {"config": "Experiment 2 learning p, dz=500, mini-batch 32 per GPU, 100 epochs, Linear(5,500), SGD lr 5e-4, tol 1e-5",
            "n_gpus": cx.world, "global_mini_batch": mini, "epochs": 100, "ms_per_epoch": ms / 100,
            "value": mini * 100 / (ms * 1e-3), "unit": UNIT, "collective": "one NCCL all_reduce of the 3000 Linear(5,500) "
            "gradient elements per epoch" if cx.world > 1 else "none (single GPU)",
            "loss_first": hist[0], "loss_last": hist[-1]}


# ------------------------------------------------------------------------------------------------
def run_b200(a):
    cx = Ctx(a)
    dev, world, rank = cx.dev, cx.world, cx.rank
    dtype = torch.float32 if a.dtype == "f32" else torch.float64
    K = a.steps if a.steps is not None else 200
    W = a.warmup if a.warmup is not None else 5
    W = max(W, 3)
    n, B, s = a.dz, a.batch, (4 if a.dtype == "f32" else 8)

    # independent problems: every rank owns its own shard of the global batch, no collective in the solve
    host_sets = [gen_data(n, B, 1000 * rank + k, dtype) for k in range(a.datasets)]
    dev_sets = [[t.to(dev) for t in d] for d in host_sets]
    log(f"rank {rank}: data on {dev}, {W} warm-up + {K} timed steps")
    sampler = ClockSampler(cx.local) if rank == 0 else None
    if sampler:
        try:
            sampler.start()      # started before the warm-up: an nvidia-smi loop needs ~0.2 s to deliver its first line
        except Exception as exc:                     # the clocks object is evidence, never a reason to lose the run
            log(f"clock sampler failed to start: {exc!r}")
            sampler = None
    main = time_layer(cx, dev_sets, B, n, dtype, K, W, sampler=sampler)
    log(f"timed region done: {main['ms_per_step']:.3f} ms per step")
    it = main["iter"]
    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks_file.get("hbm_gbs", 6650.0)

    peaks = None
    if not a.no_extras:
        try:
            peaks = measure_peaks(cx)
            log(f"peaks: L2 read {peaks['l2_read_gbs']:.0f} GB/s, HBM read {peaks['hbm_read_gbs']:.0f} GB/s, "
                f"TF32 {peaks['tf32_tflops']:.0f} TF/s, FP64 {peaks['fp64_tflops']:.1f} TF/s")
        except Exception as exc:
            peaks = {"error": repr(exc)[:300]}
    roofline = iterate_roofline(n, B, s, it, main["phases_ms"]["iterate_ms"], peaks if peaks and "error" not in peaks else None,
                                hbm_peak, a.dtype)
    roofline_factor = factor_roofline(n, B, main["phases_ms"], peaks) if a.dtype == "f32" else None

    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(cx, a, host_sets, dtype, K, W)

    roofline_hbm = configs = sweep = exp2 = None
    if not a.no_extras:
        # ---- a second roofline object from a configuration that is HBM-bound on one GPU: the fp64 (1e-8 parity) mode
        #      at the same dz / B -- its packed operator set (142 MB) no longer fits the L2
        try:
            if a.dtype == "f32":
                sets64 = [[t.to(dev) for t in gen_data(n, B, 500 + k, torch.float64)] for k in range(2)]
                r64 = time_layer(cx, sets64, B, n, torch.float64, 10, 3)
                roofline_hbm = iterate_roofline(n, B, 8, r64["iter"], r64["phases_ms"]["iterate_ms"], peaks, hbm_peak, "f64")
                roofline_hbm["config"] = f"Experiment 1 dz={n}, batch {B} per GPU, tol 1e-5, f64 (the 1e-8 parity mode)"
                roofline_hbm["value"] = r64["value"]
                roofline_hbm["ms_per_step"] = r64["ms_per_step"]
                roofline_hbm["phases_ms"] = r64["phases_ms"]
                del sets64
                torch.cuda.empty_cache()
                log(f"fp64 config: {r64['value']:.0f} QP/s, iterate {roofline_hbm['achieved_moved']:.0f} GB/s moved")
        except Exception as exc:
            roofline_hbm = {"error": repr(exc)[:300]}
        if world == 1:
            configs = run_configs(cx, a, peaks, hbm_peak)
        del dev_sets
        torch.cuda.empty_cache()
        try:
            sweep = run_sweep(cx, a, dtype)
        except Exception as exc:
            sweep = [{"error": repr(exc)[:300]}]
        try:
            exp2 = run_exp2(cx, a, dtype)
            log(f"exp2: {exp2['ms_per_epoch']:.3f} ms per epoch")
        except Exception as exc:
            exp2 = {"error": repr(exc)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a)
    # the unmodified reference through torch's CUDA library path on this GPU (secondary baseline; its own process: it
    # changes torch's default device)
    library = None
    if rank == 0 and world == 1 and not a.no_extras and have_reference():
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-cuda", "--steps", "3", "--warmup", "2",
                   "--dz", str(a.dz), "--batch", str(a.batch), "--dtype", a.dtype]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            d = json.loads(line[-1]) if line else {"unavailable": r.stderr[-300:]}
            library = {k: d[k] for k in ("value", "unit", "ms_per_step", "steps", "finite", "how", "unavailable") if k in d}
            if "value" in library:
                log(f"reference on cuda (torch library path): {library['value']:.0f} QP/s")
        except Exception as exc:
            library = {"unavailable": repr(exc)[:300]}

    if rank == 0:
        out = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": a.dtype, "data": "synthetic", "config": config_dict(a, world, it),
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": main["launches_per_step"] * K,
               "phases_ms": main["phases_ms"], "clocks": main["clocks"], "peaks": peaks, "roofline_hbm": roofline_hbm,
               "roofline_factor": roofline_factor, "configs": configs, "sweep": sweep, "exp2": exp2,
               "library_baseline": library}
        print(json.dumps(out), flush=True)
    if world > 1:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference-child":
        reference_child(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-cuda":
        run_reference_cuda(args)
    else:
        run_b200(args)
